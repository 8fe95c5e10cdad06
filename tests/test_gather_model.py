"""The gather formulation the CUDA kernels implement (one warp per CpG site walking a read window, with segments)
is equivalent to the reference's streaming flush/overwrite semantics — checked on CPU against the oracle."""
import numpy as np
import pytest

import gather_model
from metheor_b200 import batch as B
from metheor_b200 import synth
from oracle_lib import Oracle


def _data(seed, del_frac, nocall, length=40_000, cov=25.0):
    sites = synth.make_sites(seed, length, mean_gap=30.0)
    b = synth.make_reads(seed + 1, sites, length, cov, read_len=100, del_frac=del_frac, del_max=80, nocall=nocall, lowq=0.1)
    return B.to_oracle_soa([b])


@pytest.mark.parametrize("seed,del_frac,nocall", [(3, 0.0, 0.01), (4, 0.3, 0.02), (5, 0.5, 0.05)])
def test_pdr_gather_equals_streaming(seed, del_frac, nocall):
    soa = _data(seed, del_frac, nocall)
    for min_depth, min_cpgs, min_qual in ((10, 4, 10), (1, 1, 0), (5, 2, 10)):
        want = Oracle.from_soa(**soa).pdr(min_depth, min_cpgs, min_qual)
        got = gather_model.pdr(soa, min_depth, min_cpgs, min_qual)
        assert [g[0] for g in got] == list(want["pos"])
        assert [g[1] for g in got] == list(want["n_conc"])
        assert [g[2] for g in got] == list(want["n_disc"])


@pytest.mark.parametrize("seed,del_frac,nocall", [(6, 0.0, 0.01), (7, 0.3, 0.05)])
def test_strict_flush_site_sets(seed, del_frac, nocall):
    soa = _data(seed, del_frac, nocall)
    o = Oracle.from_soa(**soa)
    # MHL: every read with >=1 CpG triggers; contributors need mapq and min_cpgs
    got, _ = gather_model.site_sets_strict(soa, 10, lambda r: r["mapq"] >= 10 and len(r["pos"]) >= 4, lambda r: True)
    assert sorted(got) == list(o.mhl(10, 4, 10)["pos"])
    # FDRP: triggers == contributors == reads passing mapq with >=1 CpG (reads are <= 201 bp so no window drop)
    ok = lambda r: r["mapq"] >= 10
    got, _ = gather_model.site_sets_strict(soa, 10, ok, ok)
    assert sorted(got) == list(o.fdrp(min_depth=10, max_depth=4096, min_overlap=35)["pos"])
