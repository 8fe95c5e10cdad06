#!/bin/bash
# build_variant.sh NAME "-DING_RPT=2 -DING_MINB=8 ..."  ->  metheor_b200/csrc/variant_NAME.so (kernel-tuning experiments;
# selected at run time with METHEOR_B200_LIB=...; never shipped as the default library)
set -e
cd "$(dirname "$0")/../metheor_b200/csrc"
NAME=$1; DEFS=$2
mkdir -p build/var_$NAME
for f in engine k_ingest k_sites k_pdr k_mhl k_mhl_site k_quartet k_fdrp k_fdrp_tile k_pairs k_expand tag; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off $DEFS -c $f.cu -o build/var_$NAME/$f.o -Xptxas -v 2> build/var_$NAME/$f.log &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variant_$NAME.so build/var_$NAME/*.o -lcudart
grep -h -A2 "k_ingestE\|k_pdr_scatterE" build/var_$NAME/k_ingest.log build/var_$NAME/k_pdr.log | grep "Used"
