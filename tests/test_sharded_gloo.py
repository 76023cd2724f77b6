"""CPU suite, part 3: the N>1 host logic (row partition, padding, all-gather, merge
ordering) under torch.distributed `gloo`, world_size 2 and 3.  The shard-local
search and the merge are injected from the ORACLE here (tests only) -- on a GPU
the same ShardedCorpus is wired to the CUDA kernels (ShardedCorpus.on_gpu)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q_out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import oracle
    from conftest import unit_rows
    from semtools_b200 import capi
    from semtools_b200.sharded import ShardedCorpus, shard_bounds

    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(99)
    rows = unit_rows(rng, 1003)
    rows[1000] = rows[5]                       # cross-shard exact tie
    rows[400] = 0
    q = rows[5].copy()
    lo, hi = shard_bounds(1003, world, rank)

    def local_search(qv, top_k, max_distance, mode):
        if mode == capi.STB_MODE_SEARCH_DOCUMENTS and max_distance is not None:
            r, d = oracle.search_rows(rows[lo:hi], qv, top_k=top_k, max_distance=max_distance)    # threshold lifts top_k
        else:
            r, d = oracle.search_rows(rows[lo:hi], qv, top_k=top_k)
            if max_distance is not None:
                keep = d < max_distance
                r, d = r[keep], d[keep]
        out = np.zeros(len(r), dtype=capi.HIT_DTYPE)
        out["row"], out["distance"] = r + lo, d
        return out

    def all_gather(local):
        t = torch.from_numpy(local.view(np.float64).reshape(-1, 2).copy())
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return np.stack([o.numpy() for o in out]).view(capi.HIT_DTYPE).reshape(world, -1)

    def merge(lists, top_k):
        flat = lists.reshape(-1)
        flat = flat[flat["row"] != np.uint64(0xFFFFFFFFFFFFFFFF)]
        return flat[np.lexsort((flat["row"], flat["distance"]))][:top_k]

    sc = ShardedCorpus(rank, world, local_search, merge, all_gather)
    res = {}
    for k in (1, 3, 10, 600):
        got = sc.search(q, k)
        r, d = oracle.search_rows(rows, q, top_k=k)
        res[k] = bool(got["row"].tolist() == [int(x) for x in r] and np.array_equal(got["distance"], d))
    res["tie"] = sc.search(q, 2)["row"].tolist() == [5, 1000]
    # threshold mode (mod.rs:115-116): counts all-gather, then a payload padded to the largest count
    thr_ok = True
    for thr in (0.0, 0.3, 0.93, 1.0, 1.5):
        got = sc.search(q, 3, max_distance=thr, mode=capi.STB_MODE_SEARCH_DOCUMENTS)
        r, d = oracle.search_rows(rows, q, top_k=3, max_distance=thr)
        thr_ok = thr_ok and got["row"].tolist() == [int(x) for x in r] and np.array_equal(got["distance"], d)
    res["threshold"] = bool(thr_ok)
    res["threshold_sizes"] = [len(sc.search(q, 3, max_distance=t, mode=capi.STB_MODE_SEARCH_DOCUMENTS)) for t in (0.0, 0.93, 1.5)]

    # an APPROXIMATE shard-local search (IVF-PQ sharded by row: each shard may miss rows) --
    # the global answer must be exactly the (distance,row)-ordered union of the shard lists
    def approx_local(qv, top_k, max_distance, mode):
        keep = np.arange(lo, hi)[(np.arange(lo, hi) % 3) != 1]        # every shard "probes" 2/3 of its rows
        r, d = oracle.search_rows(rows[keep], qv, top_k=top_k)
        out = np.zeros(len(r), dtype=capi.HIT_DTYPE)
        out["row"], out["distance"] = keep[r], d
        return out

    sa = ShardedCorpus(rank, world, approx_local, merge, all_gather)
    keep_all = np.arange(1003)[(np.arange(1003) % 3) != 1]
    r, d = oracle.search_rows(rows[keep_all], q, top_k=7)
    got = sa.search(q, 7)
    res["approx_union"] = bool(got["row"].tolist() == keep_all[r].tolist() and np.array_equal(got["distance"], d))
    # batched queries: one all-gather of nq x k hits, per-query merge
    qs = np.stack([rows[5], rows[77], rows[1002], rows[400]])       # incl. a zero-row query
    def local_batch(queries, top_k):
        return [local_search(qv, top_k, None, None) for qv in queries]
    got_b = sc.search_batch(qs, 6, local_batch)
    ok = len(got_b) == 4
    for i in range(4):
        r, d = oracle.search_rows(rows, qs[i], top_k=6)
        ok = ok and got_b[i]["row"].tolist() == [int(x) for x in r] and np.array_equal(got_b[i]["distance"], d)
    res["batch"] = bool(ok)
    res["bounds"] = (lo, hi)
    q_out.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_search_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    port = 29650 + world + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q_out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q_out.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    bounds = sorted(res["bounds"] for _, res in results)
    assert bounds[0][0] == 0 and bounds[-1][1] == 1003
    assert all(a[1] == b[0] for a, b in zip(bounds[:-1], bounds[1:]))
    for rank, res in results:
        assert res["batch"], f"rank {rank}: sharded batch search differs from the oracle"
        assert res["approx_union"], f"rank {rank}: approximate shard lists merged wrongly"
        for k in (1, 3, 10, 600):
            assert res[k], (rank, k)
        assert res["tie"]
        assert res["threshold"], f"rank {rank}: sharded threshold-mode search differs from the oracle"
        assert res["threshold_sizes"][0] == 0 and res["threshold_sizes"][2] == 1003 and 0 < res["threshold_sizes"][1] < 1003


def test_shard_bounds_cover_and_are_contiguous():
    from semtools_b200.sharded import shard_bounds, pad_hits
    from semtools_b200 import capi
    for n in (0, 1, 7, 8, 100_000_000):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(x[1] == y[0] for x, y in zip(b[:-1], b[1:]))
    h = np.zeros(2, dtype=capi.HIT_DTYPE)
    p = pad_hits(h, 5)
    assert len(p) == 5 and np.isinf(p["distance"][2:]).all()
