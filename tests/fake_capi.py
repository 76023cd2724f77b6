"""CPU stand-in for semtools_b200.capi, for ONE purpose: running bench.py's control flow without a GPU
(tests/bench_sim.py).  TEST INFRASTRUCTURE -- it answers every call with the oracle (exact search) or
with numpy, it is never imported by the product and measures nothing.  "Device pointers" are addresses
of CPU tensors; the fused exchanges are played by gloo all-gathers, so a multi-rank run exercises the
same sequence of bench.py calls, collectives and deadlines as the real thing.

Fault injection for the deadline tests: FAKE_CAPI_STALL=<rank>:<method> makes that rank sleep forever
inside the named Exchange / IvfPq method; FAKE_CAPI_RAISE=<rank>:<method> makes it raise StbError;
FAKE_CAPI_TIMEOUT=<rank>:search_batch_dev makes that rank report a peer time-out (proof flag 2)."""
import ctypes
import os
import time

import numpy as np

import oracle                                         # checker: exact search / pooling on the CPU
from semtools_b200.capi import (HIT_DTYPE, STB_DIM, STB_ERR_STATE, STB_MODE_SEARCH_DOCUMENTS,   # noqa: F401  (constants only)
                                STB_MODE_STORE_QUERY, StbError)

_PAD_ROW = np.uint64(0xFFFFFFFFFFFFFFFF)
_TIERS = {"f32": 0, "h16": 1, "q8": 2}


def _rank():
    return int(os.environ.get("RANK", "0"))


def _fault(method):
    for var, act in (("FAKE_CAPI_STALL", "stall"), ("FAKE_CAPI_RAISE", "raise"), ("FAKE_CAPI_TIMEOUT", "timeout")):
        v = os.environ.get(var, "")
        if v and v.split(":")[0] == str(_rank()) and v.split(":")[1] == method:
            if act == "raise":
                raise StbError(STB_ERR_STATE, f"injected failure in {method}")
            if act == "timeout":
                return True                                  # the caller reports "a peer never arrived"
            while True:
                time.sleep(1.0)
    return False


def _view(ptr, shape, dtype):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    buf = (ctypes.c_char * n).from_address(int(ptr))
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def _pad(hits, k):
    out = np.zeros(k, dtype=HIT_DTYPE)
    out["distance"], out["row"] = np.inf, _PAD_ROW
    out[: len(hits)] = hits[:k]
    return out


def _merge(lists, k):
    allh = np.concatenate([np.asarray(x, dtype=HIT_DTYPE).reshape(-1) for x in lists])
    allh = allh[allh["row"] != _PAD_ROW]
    order = np.lexsort((allh["row"], allh["distance"]))
    return allh[order][:k]


def _all_gather(arr):
    """gloo all-gather of an equal-shaped numpy array (world 1: identity)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [arr]
    t = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1).copy())
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, t)
    return [p.numpy().view(arr.dtype).reshape(arr.shape) for p in parts]


class Context:
    def __init__(self, device=0, stream=None):
        self.device, self.stream, self.launches = device, int(stream or 1), 0

    def close(self):
        pass

    def sync(self):
        pass

    def counters(self):
        return {"kernel_launches": self.launches, "fallback_searches": 0}

    def ticket_check(self):
        return 0, 0

    def hits_merge(self, lists, top_k):
        return _merge(list(np.asarray(lists, dtype=HIT_DTYPE)), top_k)

    def hits_merge_dev(self, lists_dev, n_lists, per_list, top_k, out_dev):
        lists = _view(lists_dev, (n_lists, per_list), HIT_DTYPE)
        _view(out_dev, (top_k,), HIT_DTYPE)[:] = _pad(_merge(list(lists), top_k), top_k)
        self.launches += 1

    def hits_merge_batch_dev(self, lists_dev, n_lists, nq, per_list, top_k, out_dev):
        lists = _view(lists_dev, (n_lists, nq, per_list), HIT_DTYPE)
        out = _view(out_dev, (nq, top_k), HIT_DTYPE)
        for q in range(nq):
            out[q] = _pad(_merge(list(lists[:, q]), top_k), top_k)
        self.launches += 1


class Corpus:
    def __init__(self, ctx, capacity_rows=1024, row_base=0):
        self.ctx, self.row_base = ctx, int(row_base)
        self.rows = np.zeros((0, STB_DIM), dtype=np.float32)
        self.built = set()
        self.stats = {t: [0, 0] for t in _TIERS}

    def close(self):
        self.rows = np.zeros((0, STB_DIM), dtype=np.float32)

    def __len__(self):
        return len(self.rows)

    def append(self, rows):
        self.rows = np.concatenate([self.rows, np.ascontiguousarray(rows, dtype=np.float32)])

    def append_dev(self, rows_dev, n):
        self.append(_view(rows_dev, (n, STB_DIM), np.float32).copy())

    def read(self, first=0, n=None):
        n = len(self) - first if n is None else n
        return self.rows[first:first + n].copy()

    def prepare(self, what=3):
        if what & 1:
            self.built.add("q8")
        if what & 2:
            self.built.add("h16")

    def prepare_batch(self):
        self.built.add("h16")

    def tier_stats(self):
        return {t: {"tries": v[0], "proven": v[1], "built_rows": len(self) if (t == "f32" or t in self.built) else 0}
                for t, v in self.stats.items()}

    def _tier(self, k):
        want = os.environ.get("STB_SCAN_TIER", "q8")
        for t in ("q8", "h16"):
            if _TIERS[t] <= _TIERS.get(want, 2) and t in self.built and (t != "q8" or k <= 16):
                return t
        return "f32"

    def _exact(self, q, k, max_distance=None):
        if len(self.rows) == 0 or k == 0:
            return np.zeros(0, dtype=HIT_DTYPE)
        r, d = oracle.search_rows(self.rows, np.ascontiguousarray(q, dtype=np.float32), top_k=k, max_distance=max_distance)
        out = np.zeros(len(r), dtype=HIT_DTYPE)
        out["distance"], out["row"] = d, np.asarray(r, dtype=np.uint64) + np.uint64(self.row_base)
        return out

    def search(self, q, top_k=3, max_distance=None, mode=STB_MODE_SEARCH_DOCUMENTS, row_ranges=None, cap=None):
        t = self._tier(top_k)
        self.stats[t][0] += 1; self.stats[t][1] += 1
        self.ctx.launches += 1
        return self._exact(q, top_k, max_distance)

    def search_topk_dev(self, q_dev, top_k, out_hits_dev, out_status_dev):
        q = _view(q_dev, (STB_DIM,), np.float32)
        hits = self.search(q, top_k)
        _view(out_hits_dev, (top_k,), HIT_DTYPE)[:] = _pad(hits, top_k)
        _view(out_status_dev, (4,), np.uint32)[:] = [len(hits), 1, len(hits), 32 | (_TIERS[self._tier(top_k)] << 16)]

    def search_many(self, queries, top_k=10, xchg=None):
        if xchg is None:
            return [self.search(q, top_k) for q in queries]
        res = [xchg.search(self, q, top_k) for q in queries]
        return [r[0] for r in res], np.array([r[1] for r in res], dtype=bool)

    def search_batch(self, queries, top_k=10):
        self.ctx.launches += 5
        return [self._exact(q, top_k) for q in queries]

    def search_batch_dev(self, q_dev, nq, top_k, out_hits_dev, out_status_dev):
        qs = _view(q_dev, (nq, STB_DIM), np.float32)
        out = _view(out_hits_dev, (nq, top_k), HIT_DTYPE)
        st = _view(out_status_dev, (nq, 2), np.uint32)
        for i in range(nq):
            h = self._exact(qs[i], top_k)
            out[i] = _pad(h, top_k)
            st[i] = [len(h), 1]
        self.ctx.launches += 5


class Exchange:
    HANDLE_BYTES = 64

    def __init__(self, ctx, world, rank, max_k, max_nq=0):
        self.ctx, self.world, self.rank, self.max_k, self.max_nq = ctx, world, rank, max_k, max_nq

    def close(self):
        pass

    def local_handle(self):
        return bytes([self.rank]) * self.HANDLE_BYTES

    def connect(self, handles):
        assert len(handles) == self.world and all(len(h) == self.HANDLE_BYTES for h in handles)

    def _global(self, corpus, q, k):
        local = _pad(corpus.search(q, k), k)
        return _merge(_all_gather(local), k)

    def search(self, corpus, q, top_k):
        _fault("search")
        return self._global(corpus, np.ascontiguousarray(q, dtype=np.float32), top_k), True

    def search_topk(self, corpus, q_dev, top_k, out_hits_dev, out_status_dev):
        _fault("search_topk")
        hits = self._global(corpus, _view(q_dev, (STB_DIM,), np.float32), top_k)
        _view(out_hits_dev, (top_k,), HIT_DTYPE)[:] = _pad(hits, top_k)
        _view(out_status_dev, (4,), np.uint32)[:] = [len(hits), 1, len(hits), 32 | (_TIERS[corpus._tier(top_k)] << 16)]

    def search_batch_dev(self, corpus, q_dev, nq, top_k, out_hits_dev, out_status_dev):
        timed_out = _fault("search_batch_dev")
        assert self.max_nq >= nq
        local = np.zeros((nq, top_k), dtype=HIT_DTYPE)
        st_local = np.zeros((nq, 2), dtype=np.uint32)
        t = np.zeros((nq, top_k, 2), dtype=np.float64)
        corpus.search_batch_dev(q_dev, nq, top_k, local.ctypes.data, st_local.ctypes.data)
        del t
        parts = _all_gather(local)
        out = _view(out_hits_dev, (nq, top_k), HIT_DTYPE)
        st = _view(out_status_dev, (nq, 2), np.uint32)
        for i in range(nq):
            h = _merge([p[i] for p in parts], top_k)
            out[i] = _pad(h, top_k)
            st[i] = [len(h), 2 if timed_out else 1]
        self.ctx.launches += 2


class IvfPq:
    """Answers with the exact scan (recall 1.0): the sim checks bench.py's plumbing, not the index."""

    def __init__(self, corpus, nlist=1024, train_rows=65536, iters=8):
        self.corpus, self.nlist = corpus, nlist

    def close(self):
        pass

    def stats(self):
        return {"rows": len(self.corpus), "nlist": self.nlist, "max_list": max(1, len(self.corpus) // max(self.nlist, 1)),
                "index_bytes": len(self.corpus) * 32}

    def search(self, q, nprobe=64, top_k=10, rerank=256):
        _fault("ivf_search")
        return self.corpus._exact(q, top_k), min(len(self.corpus), 1000)

    def search_dev(self, q_dev, nprobe, top_k, rerank, out_hits_dev, out_status_dev):
        _fault("ivf_search_dev")
        hits = self.corpus._exact(_view(q_dev, (STB_DIM,), np.float32), top_k)
        _view(out_hits_dev, (top_k,), HIT_DTYPE)[:] = _pad(hits, top_k)
        _view(out_status_dev, (2,), np.uint32)[:] = [len(hits), min(len(self.corpus), 1000)]
        self.corpus.ctx.launches += 2


class Table:
    def __init__(self, ctx, E, weights=None, mapping=None, normalize=True):
        self.ctx, self.E = ctx, np.ascontiguousarray(E, dtype=np.float32)

    def close(self):
        pass


def embed(ctx, table, offsets, ids, out=True, append_to=None):
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    ids = np.ascontiguousarray(ids, dtype=np.uint32)
    rows = oracle.embed_csr(table.E, offsets, ids)
    if append_to is not None:
        append_to.append(rows)
    ctx.launches += 1
    return rows if out else None


def embed_dev(ctx, table, offsets_dev, ids_dev, n_lines, out_dev):
    offsets = _view(offsets_dev, (n_lines + 1,), np.uint64)
    ids = _view(ids_dev, (int(offsets[-1]),), np.uint32) if int(offsets[-1]) else np.zeros(0, np.uint32)
    _view(out_dev, (n_lines, STB_DIM), np.float32)[:] = oracle.embed_csr(table.E, offsets.copy(), ids.copy())
    ctx.launches += 1


def embed_status(ctx):
    pass
