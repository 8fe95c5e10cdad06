// main.cpp — the `metheor` binary: argv -> mthh_main (cli.cpp).
#include "../../include/metheor_host.h"

int main(int argc, char** argv) { return mthh_main(argc, argv); }
