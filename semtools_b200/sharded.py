"""Row-sharded search across ranks: one process per GPU, contiguous
row blocks, ONE exchange per query -- an all-gather of k stb_hit (16 B) per rank --
then K4 (stb_hits_merge) on every rank.

The collective is whatever `torch.distributed` backend the host initialised (NCCL
over NVLink on GPUs).  The shard-local search and the merge are injected callables
so the partition / padding / ordering logic can be exercised on CPU with `gloo`
(tests/test_sharded_gloo.py injects the oracle there -- tests only).  The product
wiring, `ShardedCorpus.on_gpu`, uses libsemtools_b200 for both.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable

import numpy as np

from . import capi

PAD_ROW = np.uint64(0xFFFFFFFFFFFFFFFF)


def shard_bounds(n_rows: int, world: int, rank: int):
    """GPU g owns rows [g*ceil(N/G), min(N,(g+1)*ceil(N/G)))  (SURVEY 8e)."""
    per = (n_rows + world - 1) // world
    return min(rank * per, n_rows), min((rank + 1) * per, n_rows)


def pad_hits(hits: np.ndarray, k: int) -> np.ndarray:
    out = np.zeros(k, dtype=capi.HIT_DTYPE)
    out["distance"] = np.inf
    out["row"] = PAD_ROW
    out[: len(hits)] = hits[:k]
    return out


@dataclass
class ShardedCorpus:
    rank: int
    world: int
    local_search: Callable      # (q, top_k, max_distance, mode) -> HIT_DTYPE array, global rows
    merge: Callable             # ((world, k) HIT_DTYPE, top_k) -> HIT_DTYPE array
    all_gather: Callable        # (k,) HIT_DTYPE -> (world, k) HIT_DTYPE

    def search(self, q, top_k: int, max_distance=None, mode=capi.STB_MODE_STORE_QUERY):
        """Global result of one query over all shards.

        top_k caps the result (STORE_QUERY always; SEARCH_DOCUMENTS without a threshold): ONE
        exchange of k hits per rank, then the K4 merge.
        SEARCH_DOCUMENTS with `max_distance` lifts the cap (src/search/mod.rs:115-116: every line
        under the threshold is returned): the per-shard result size is data-dependent, so the ranks
        first all-gather their COUNTS, then a payload padded to the largest count, and
        every rank merges by (distance,row) -- the reference's stable sort over (doc,line) order."""
        threshold_all = mode == capi.STB_MODE_SEARCH_DOCUMENTS and max_distance is not None
        if top_k == 0 and not threshold_all:
            return np.zeros(0, dtype=capi.HIT_DTYPE)
        if not threshold_all:
            local = pad_hits(self.local_search(q, top_k, max_distance, mode), top_k)
            lists = self.all_gather(local)
            return self.merge(lists, top_k)
        local = np.ascontiguousarray(self.local_search(q, top_k, max_distance, mode), dtype=capi.HIT_DTYPE)
        cnt = np.zeros(1, dtype=capi.HIT_DTYPE)
        cnt["row"] = len(local)                                   # counts ride in the row field of one 16-byte hit
        counts = self.all_gather(cnt)["row"].reshape(-1).astype(np.int64)
        widest = int(counts.max())
        if widest == 0:
            return np.zeros(0, dtype=capi.HIT_DTYPE)
        lists = self.all_gather(pad_hits(local, widest))          # (world, widest), padding = (+inf, PAD_ROW)
        return self.merge(lists, int(counts.sum()))

    def search_batch(self, queries, top_k: int, local_search_batch: Callable):
        """Batched queries (K2) over the row shards: every rank answers all nq queries on its
        shard (`local_search_batch(queries, top_k)` -> list of HIT_DTYPE arrays with global
        rows, e.g. `Corpus.search_batch`), ONE all-gather of nq x top_k hits (16 B each), then the
        per-query merge.  Returns a list of nq HIT_DTYPE arrays."""
        queries = np.ascontiguousarray(queries, dtype=np.float32)
        nq = len(queries)
        if nq == 0 or top_k == 0:
            return [np.zeros(0, dtype=capi.HIT_DTYPE) for _ in range(nq)]
        local = np.stack([pad_hits(h, top_k) for h in local_search_batch(queries, top_k)])   # (nq, k)
        lists = self.all_gather(local.reshape(-1)).reshape(self.world, nq, top_k)
        return [self.merge(np.ascontiguousarray(lists[:, i, :]), top_k) for i in range(nq)]

    @classmethod
    def on_gpu(cls, ctx: capi.Context, corpus: capi.Corpus, dist, device):
        """Product wiring: CUDA kernels for search + merge, NCCL all-gather."""
        def local_search(q, top_k, max_distance, mode):
            return corpus.search(q, top_k, max_distance, mode)

        def merge(lists, top_k):
            # the K4 kernel sorts up to 4096 hits in shared memory; threshold-mode results beyond that
            # are ordered on the host (ordering bookkeeping, no distance is computed here)
            if lists.size <= 4096:
                return ctx.hits_merge(lists, top_k)
            flat = lists.reshape(-1)
            flat = flat[flat["row"] != PAD_ROW]
            return flat[np.lexsort((flat["row"], flat["distance"]))][:top_k]

        return cls(dist.get_rank(), dist.get_world_size(), local_search, merge, _device_all_gather(dist, device))

    @classmethod
    def on_gpu_ivfpq(cls, ctx: capi.Context, index: "capi.IvfPq", dist, device, nprobe: int = 64, rerank: int = 256):
        """K5 sharded by ROW: every rank holds an IVF-PQ index over its own row
        block (its corpus was created with row_base = the block's first global row, so the
        exactly re-ranked hits already carry global rows); per-rank top-k, the same 16 B x k
        all-gather, the same K4 merge.  Approximate per shard => approximate globally; the
        merge itself is exact, so recall is the mean of the shard recalls' union."""
        def local_search(q, top_k, max_distance, mode):
            if max_distance is not None:
                raise ValueError("the IVF-PQ probe has no distance threshold")
            return index.search(q, nprobe=nprobe, top_k=top_k, rerank=rerank)[0]

        return cls(dist.get_rank(), dist.get_world_size(), local_search, ctx.hits_merge, _device_all_gather(dist, device))


def _device_all_gather(dist, device):
    import torch
    world = dist.get_world_size()

    def all_gather(local):
        t = torch.from_numpy(local.view(np.float64).reshape(-1, 2).copy())
        if dist.get_backend() == "gloo":                   # CPU collectives (tests; bench.py's one-GPU functional mode)
            parts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, t)
            out = torch.stack(parts)
        else:
            t = t.to(device)
            out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=device)
            dist.all_gather_into_tensor(out, t)
            out = out.cpu()
        return np.ascontiguousarray(out.numpy()).view(capi.HIT_DTYPE).reshape(world, -1)

    return all_gather
