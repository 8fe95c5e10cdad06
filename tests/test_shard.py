"""Multi-GPU host logic on CPU: world_size-2 gloo processes shard one read set by position bins with a
halo (and by contigs), compute their shard with the oracle standing in for the GPU engine, all-reduce LPMD's counters
and gather the owned rows; rank 0 checks the merged result against one oracle pass over everything."""
import os
import socket

import numpy as np
import pytest

from metheor_b200 import batch as B
from metheor_b200 import shard, synth
from oracle_lib import Oracle

LENS = [140_000, 60_000]


def _data():
    out = []
    for tid, L in enumerate(LENS):
        sites = synth.make_sites(300 + tid, L)
        b = synth.make_reads(310 + tid, sites, L, 20.0, tid=tid, read_len=140, del_frac=0.3, del_max=50, nocall=0.03, lowq=0.1)
        out.append(b)
    return out


def _oracle_rows(batches):
    o = Oracle.from_soa(**B.to_oracle_soa(batches))
    pdr = o.pdr(5, 2, 10)
    res = dict(pdr=dict(tid=pdr["tid"], pos=pdr["pos"], value=pdr["pdr"], n_conc=pdr["n_conc"], n_disc=pdr["n_disc"]),
               mhl=o.mhl(5, 2, 10), fdrp=o.fdrp(min_depth=5, min_overlap=20), qfdrp=o.qfdrp(min_depth=5, min_overlap=20))
    q = o.quartets(5, 10)
    res["pm"] = dict(tid=q["tid"], pos=q["p1"], p2=q["p2"], p3=q["p3"], p4=q["p4"], value=q["pm"])
    return res


def _oracle_lpmd(batches):
    if not batches:
        return np.zeros(4, np.int64)
    w = Oracle.from_soa(**B.to_oracle_soa(batches)).lpmd(2, 16, 10)
    return np.array([w["n_read"], w["n_valid_read"], w["n_conc"], w["n_disc"]], np.int64)


def _same(a, b, what):
    for k in a:
        if isinstance(a[k], np.ndarray):
            x, y = a[k], b[k]
            if x.dtype == np.float32:
                x, y = x.view(np.uint32), np.asarray(y, np.float32).view(np.uint32)
            assert np.array_equal(x, y), (what, k)


def _worker(rank, world, port, mode, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        data = _data()
        if mode in ("bins", "bins_region_cost"):
            # read-count balanced cuts, or length + a cost per contig boundary (what bench.py --gpus N uses)
            plan = shard.plan_bins(LENS, world, weights=[b["start"] for b in data]) if mode == "bins" else \
                shard.plan_bins(LENS, world, region_cost=max(1, min(LENS) // 3))
            mine, owned = [], []
            for b in data:
                sub, n_own = shard.select_shard(b, plan[rank], halo=400)
                if sub is not None and sub["n_reads"]:
                    mine.append(sub)
                    own_mask = (sub["meta"] & shard.META_HALO) == 0
                    if own_mask.any():
                        owned.append(B.select_reads(sub, own_mask))
            intervals = plan[rank]
        else:
            tids = shard.plan_contigs(LENS, world)[rank]
            mine = [data[t] for t in tids]
            owned = mine
            intervals = [(t, 0, LENS[t]) for t in tids]
        rows = _oracle_rows(mine) if mine else None
        part = {m: shard.owned_rows(r, intervals) for m, r in rows.items()} if rows else None
        lp = torch.from_numpy(_oracle_lpmd(owned))
        dist.all_reduce(lp)  # the one real exchange of the path: LPMD's four int64 counters
        gathered = [None] * world
        dist.all_gather_object(gathered, part)
        if rank == 0:
            whole = _oracle_rows(data)
            for m in whole:
                keys = ("tid", "pos", "p2", "p3", "p4") if m == "pm" else ("tid", "pos")
                merged = shard.merge_rows([g[m] for g in gathered if g is not None], keys=keys)
                assert merged["n"] == len(whole[m]["tid"]) and merged["n"] > 100, (m, merged["n"], len(whole[m]["tid"]))
                _same(whole[m], merged, m)
            assert np.array_equal(lp.numpy(), _oracle_lpmd(data))
            q.put("ok")
    except Exception as e:  # surface the failure in the parent
        if rank == 0:
            q.put(f"rank0 failed: {e!r}")
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["bins", "bins_region_cost", "contigs"])
def test_world_size_2_gloo(mode):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert q.get() == "ok"


def test_plans_cover_the_genome_exactly():
    for world in (1, 2, 3, 8):
        for plan in (shard.plan_bins(LENS, world), shard.plan_bins(LENS, world, weights=[b["start"] for b in _data()]),
                     shard.plan_bins(LENS, world, region_cost=min(LENS) // 3), shard.plan_bins(LENS, world, region_cost=10 * max(LENS))):
            cover = np.zeros(sum(LENS), np.int32)
            base = np.concatenate([[0], np.cumsum(LENS)])
            for r in plan:
                for tid, lo, hi in r:
                    assert 0 <= lo < hi <= LENS[tid]
                    cover[base[tid] + lo: base[tid] + hi] += 1
            assert (cover == 1).all()
        c = shard.plan_contigs(LENS, world)
        assert sorted(t for r in c for t in r) == list(range(len(LENS)))


def test_cpp_host_plans_equal_the_python_planner():
    """`metheor --gpus N` (host/run.cpp plan_bins / plan_contigs, through mthh_plan_shards) and shard.py agree."""
    from metheor_b200 import host, synth_gpu
    rng = np.random.default_rng(3)
    cases = [[l for _, l in synth_gpu.HG38], [1000], [5, 7, 11, 13], list(rng.integers(1, 10_000_000, 40))]
    for rl in cases:
        for w in (1, 2, 3, 4, 8, 16):
            a = host.plan_shards(rl, w)
            b = shard.plan_bins([int(x) for x in rl], w)
            assert [[tuple(int(v) for v in x) for x in r] for r in a] == [[tuple(int(v) for v in x) for x in r] for r in b], (rl, w)
            # the intervals of all ranks tile the genome exactly
            cover = sorted(x for r in a for x in r)
            for t, l in enumerate(rl):
                iv = [(lo, hi) for tt, lo, hi in cover if tt == t]
                assert iv[0][0] == 0 and iv[-1][1] == l and all(iv[k][1] == iv[k + 1][0] for k in range(len(iv) - 1)), (rl, w, t)
            c = host.plan_shards(rl, w, True)
            assert [[t for t, _, _ in r] for r in c] == shard.plan_contigs([int(x) for x in rl], w)
            # bins balanced on bases + region_cost x contigs (METHEOR_SHARD_REGION_COST in the C++ host)
            for cost in (1, 5_000, 30_000_000):
                os.environ["METHEOR_SHARD_REGION_COST"] = str(cost)
                try:
                    a = host.plan_shards(rl, w)
                finally:
                    del os.environ["METHEOR_SHARD_REGION_COST"]
                b = shard.plan_bins([int(x) for x in rl], w, region_cost=cost)
                assert [[tuple(int(v) for v in x) for x in r] for r in a] == [[tuple(int(v) for v in x) for x in r] for r in b], (rl, w, cost)
                cover = sorted(x for r in a for x in r)
                for t, l in enumerate(rl):
                    iv = [(lo, hi) for tt, lo, hi in cover if tt == t]
                    assert iv[0][0] == 0 and iv[-1][1] == l and all(iv[k][1] == iv[k + 1][0] for k in range(len(iv) - 1)), (rl, w, t, cost)
