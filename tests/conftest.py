import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # The built libraries normally travel with the tree; if a checkout arrives without them (they are git-ignored),
    # build once — nvcc cross-compiles sm_100a without a GPU, so this works on the CPU box as well.
    need = [os.path.join(ROOT, "metheor_b200", "csrc", "libmetheor_b200.so"),
            os.path.join(ROOT, "metheor_b200", "host", "libmetheor_host.so"),
            os.path.join(ROOT, "metheor_b200", "bin", "metheor"),
            os.path.join(ROOT, "oracle", "_build", "liboracle.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(HERE, "golden", "fixtures.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def pins():
    with open(os.path.join(HERE, "golden", "reference_pins.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def fixture_bams(golden, tmp_path_factory):
    """Rebuild the reference's fixture BAMs (tests/test{1..6}.bam) + the chr19 1000-read set from fixtures.json."""
    import bamio
    d = tmp_path_factory.mktemp("bams")
    out = {}
    for name, fx in golden.items():
        path = str(d / f"{name}.bam")
        bamio.write_bam(path, [tuple(r) for r in fx["refs"]], fx["reads"], header_text=fx.get("header_text"))
        out[name] = path
    return out
