"""The LPMD pair count as k_ingest formulates it (packed key, two nearest anchors inline, walk from the third:
tests/ingest_model.py) equals the reference's anchor-list walk (readutil.rs:166-224, oracle) — checked on CPU."""
import random

import pytest

import ingest_model
from metheor_b200 import batch as B
from metheor_b200 import synth
from oracle_lib import Oracle

WINDOWS = [(2, 16), (0, 3), (5, 5), (17, 16), (1, 10 ** 9), (3, 150), (4, 6), (2, 2), (1, 1), (16, 17)]


def _brute(rel, meth, dmin, dmax):
    c = d = 0
    for i in range(len(rel)):
        for z in range(i):
            if dmin <= rel[i] - rel[z] <= dmax:
                if meth[i] == meth[z]:
                    c += 1
                else:
                    d += 1
    return c, d


def test_packed_key_window_is_exact_on_random_reads():
    rng = random.Random(5)
    for _ in range(3000):
        n = rng.randint(0, 40)
        step = rng.choice((1, 2, 3, 8, 40, 2000))
        rel, r = [], rng.randint(0, 5)
        for _k in range(n):
            rel.append(r)
            r += rng.randint(1, step)
        if rel and rel[-1] > 65535:
            continue
        meth = [rng.randint(0, 1) for _k in range(n)]
        dmin = rng.choice((0, 1, 2, 3, 5, 16, 17, 100))
        dmax = rng.choice((0, 1, 2, 3, 5, 16, 17, 100, 65535, 10 ** 9))
        assert ingest_model.read_pairs(rel, meth, dmin, dmax) == _brute(rel, meth, dmin, dmax), (rel, meth, dmin, dmax)


@pytest.mark.parametrize("seed,mean_gap,kw", [(41, 30.0, {}), (42, 5.0, {}),
                                             (43, 30.0, dict(read_len=100, del_frac=0.5, del_max=60, nocall=0.05, lowq=0.15))])
def test_model_equals_oracle(seed, mean_gap, kw):
    length = 12_000
    sites = synth.make_sites(seed, length, mean_gap=mean_gap)
    soa = B.to_oracle_soa([synth.make_reads(seed + 100, sites, length, 12.0, **kw)])
    o = Oracle.from_soa(**soa)
    for dmin, dmax in WINDOWS:
        w = o.lpmd(min_distance=dmin, max_distance=dmax, min_qual=10)
        assert ingest_model.lpmd(soa, dmin, dmax, 10) == (w["n_conc"], w["n_disc"]), (dmin, dmax)
