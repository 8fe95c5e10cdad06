"""GPU parity of LPMD's distance window (lpmd.rs:175-199, readutil.rs:166-224) at its edges: k_ingest counts the two
nearest anchors of a call from a packed key (4 * query index + methylation bit) and only walks for the third and later
ones, so the window arithmetic is exercised here with empty, one-wide, zero-based, read-long and absurdly large windows,
on sparse reads, on CpG islands (many anchors per call) and on reads with deletions (query index != reference offset)."""
import numpy as np
import pytest

import parity
from metheor_b200 import synth

pytestmark = pytest.mark.gpu

WINDOWS = [(2, 16), (0, 3), (5, 5), (17, 16), (1, 10 ** 9), (3, 150), (4, 6), (2, 2)]


def _sparse():
    sites = synth.make_sites(31, 120_000)
    return synth.make_reads(32, sites, 120_000, 25.0), 120_000


def _islands():
    sites = synth.make_sites(33, 40_000, mean_gap=5.0)
    return synth.make_reads(34, sites, 40_000, 20.0), 40_000


def _deletions():
    sites = synth.make_sites(35, 120_000)
    return synth.make_reads(36, sites, 120_000, 25.0, read_len=140, del_frac=0.5, del_max=60, nocall=0.05, lowq=0.15), 120_000


@pytest.mark.parametrize("make", [_sparse, _islands, _deletions])
def test_lpmd_distance_windows(make):
    b, length = make()
    seen = set()
    for dmin, dmax in WINDOWS:
        res, _ = parity.check_all([b], [length], ("lpmd",), lpmd=dict(min_distance=dmin, max_distance=dmax))
        g = res["lpmd"]
        seen.add((g["n_conc"], g["n_disc"]))
        if dmin > dmax:
            assert g["n_conc"] == 0 and g["n_disc"] == 0 and np.isnan(g["lpmd"])
    assert len(seen) > 3  # the windows really select different pair sets


def test_lpmd_window_with_pdr_and_compact_batches():
    """Same windows through the compact wire format (query indices implied for plain `<len>M` reads) next to PDR."""
    b, length = _islands()
    for dmin, dmax in ((1, 40), (6, 9)):
        parity.check_all([b], [length], ("pdr", "lpmd"), compact="dense", lpmd=dict(min_distance=dmin, max_distance=dmax))
