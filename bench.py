#!/usr/bin/env python
"""bench.py -- queries/sec of the semtools `search` scan on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps 200 --warmup 10
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference            # the reference's CPU scan on the host cores

A "step" is one query: one pass of the hot path (K1 cosine scan + running top-k +
exact re-rank [+ NCCL all-gather + K4 merge when sharded]) over the whole corpus.
Workload: the 10M-line x 256 f32 corpus BASELINE.json's metric is quoted on, top-k 10,
single query (the HBM-bound brute-force scan; corpus = 10.24 GB >> 126 MB L2, so every
step streams from HBM and no L2 flush is needed).  N > 1 shards the SAME corpus
row-wise across ranks ("strong" scaling), one process per GPU.

PyTorch is used here for plumbing only: device allocation of the synthetic corpus,
CUDA events on the launching stream, torch.distributed (NCCL) for the 160-byte
per-rank top-k exchange.  All compute is libsemtools_b200.so via its C ABI.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CHUNK = 1_000_000
SEED = 0x5E117003
METRIC = "queries/sec over 10M-line corpus, top-k=10"


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=200)
    p.add_argument("--warmup", type=int, default=10)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--rows", type=int, default=10_000_000)
    p.add_argument("--topk", type=int, default=10)
    p.add_argument("--cpu-sample-rows", type=int, default=1_000_000)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--ivfpq-rows", type=int, default=4_000_000,
                   help="rows of the IVF-PQ side section (BASELINE configs[4] names 100M; 4M keeps the default run short)")
    p.add_argument("--ivfpq-sharded", action="store_true",
                   help="N>1: also measure IVF-PQ sharded by row (off by default: collectives on every rank)")
    p.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                   help="N>1: p2p = fused in-kernel exchange over NVLink peer memory; nccl = all-gather + merge kernel")
    return p.parse_args()


def workload_name(rows, topk):
    return (f"{rows}-line corpus x 256 f32 (N(0,1) rows L2-normalised, 0.1% duplicate + 0.01% zero rows), "
            f"single query, top-k={topk}, brute-force cosine scan (BASELINE configs[1] kernel at the "
            f"metric's corpus size)")


# ------------------------------------------------------------------ synthetic data ---
def gen_chunk_torch(torch, dev, chunk_id, rows):
    """Chunk `chunk_id` of the global corpus: identical whatever the rank count."""
    g = torch.Generator(device=dev)
    g.manual_seed(SEED + chunk_id)
    x = torch.randn((rows, 256), generator=g, device=dev, dtype=torch.float32)
    x /= x.norm(dim=1, keepdim=True)
    n_dup, n_zero = max(rows // 1000, 1), max(rows // 10000, 1)
    idx = torch.randint(0, rows, (2 * n_dup + n_zero,), generator=g, device=dev)
    x[idx[:n_dup]] = x[idx[n_dup:2 * n_dup]]
    x[idx[2 * n_dup:]] = 0.0
    return x


def gen_chunk_numpy(chunk_id, rows):
    rng = np.random.default_rng(SEED + chunk_id)
    x = rng.standard_normal((rows, 256), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    n_dup, n_zero = max(rows // 1000, 1), max(rows // 10000, 1)
    idx = rng.integers(0, rows, 2 * n_dup + n_zero)
    x[idx[:n_dup]] = x[idx[n_dup:2 * n_dup]]
    x[idx[2 * n_dup:]] = 0.0
    return np.ascontiguousarray(x, dtype=np.float32)


def gen_queries(n):
    rng = np.random.default_rng(SEED - 1)
    q = rng.standard_normal((n, 256)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return np.ascontiguousarray(q, dtype=np.float32)


# ------------------------------------------------------------------ clock sampling ---
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [s.strip() for s in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------ CPU baseline -----
def cpu_baseline(sample_rows_arr, queries, rows_total, topk, threads):
    """Times oracle/cpu_baseline.c (the reference's scan restated) on a bounded sample
    and scales linearly in rows (optimistic for the CPU: its sort is N log N)."""
    import oracle
    n = sample_rows_arr.shape[0]
    oracle.baseline_search(sample_rows_arr[:1000], queries[0], topk, threads=threads)   # load lib
    times = []
    for i in range(3):
        t0 = time.perf_counter()
        oracle.baseline_search(sample_rows_arr, queries[i % len(queries)], topk, threads=threads)
        times.append(time.perf_counter() - t0)
    t = float(np.median(times))
    scale = rows_total / n
    return 1.0 / (t * scale), t


def run_reference(args):
    """--impl reference: the reference's own CPU path.  The Rust reference cannot be
    built here (no cargo/rustc; model2vec-rs / simsimd not vendored), so this is the
    oracle port of search_documents (oracle/cpu_baseline.c), single-threaded like the
    reference's loop (src/search/mod.rs:84-104)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    sample = min(args.cpu_sample_rows, args.rows)
    rows = gen_chunk_numpy(0, sample)
    queries = gen_queries(8)
    steps = max(1, min(args.steps, 5))
    for i in range(min(args.warmup, 1)):
        oracle.baseline_search(rows, queries[i], args.topk, threads=1)
    t0 = time.perf_counter()
    for i in range(steps):
        oracle.baseline_search(rows, queries[i % 8], args.topk, threads=1)
    dt = (time.perf_counter() - t0) / steps
    scale = args.rows / sample
    qps = 1.0 / (dt * scale)
    omp_threads = oracle.baseline_threads()
    t1 = time.perf_counter()
    oracle.baseline_search(rows, queries[0], args.topk, threads=omp_threads)
    dt_omp = time.perf_counter() - t1
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": 0,
        "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * scale * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args.rows, args.topk), "rows": args.rows, "top_k": args.topk},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": 1, "kind": "port",
                         "sample": f"{sample} of {args.rows} rows per step, time scaled x{scale:g} (linear; "
                                   f"the reference's full sort is N log N so this favours the CPU)",
                         "host_cores": os.cpu_count(),
                         "all_cores_not_reference_behaviour": {"threads": omp_threads,
                                                               "value": 1.0 / (dt_omp * scale)}},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ K2 side bench -----
def bench_batch(torch, dev, ctx, stream, corpus, rows, k, nq=1024, iters=5):
    """BASELINE configs[2]: 10M-line corpus, batch of 1024 queries, top-k=10, 1xB200 through
    the tcgen05 path (stb_search_batch*).  FLOPs = 2*Q*N*256."""
    from semtools_b200 import capi
    qh = gen_queries(nq + 64)[64:]                       # distinct from the single-query set
    q_dev = torch.from_numpy(qh).to(dev)
    hits = torch.zeros((nq, k, 2), dtype=torch.float64, device=dev)
    st = torch.zeros((nq, 2), dtype=torch.int32, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); corpus.prepare_batch(); e1.record(stream); torch.cuda.synchronize(dev)
    shadow_ms = e0.elapsed_time(e1)
    for _ in range(2):
        corpus.search_batch_dev(q_dev.data_ptr(), nq, k, hits.data_ptr(), st.data_ptr())
    torch.cuda.synchronize(dev)
    l0 = ctx.counters()["kernel_launches"]
    e0.record(stream)
    for _ in range(iters):
        corpus.search_batch_dev(q_dev.data_ptr(), nq, k, hits.data_ptr(), st.data_ptr())
    e1.record(stream); torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / iters
    launches = (ctx.counters()["kernel_launches"] - l0) / iters
    proven = int((st[:, 1] == 1).sum().item())
    t0 = time.perf_counter()
    res = corpus.search_batch(qh, top_k=k)               # host queries in, host hits out
    e2e_s = time.perf_counter() - t0
    # parity: a few queries against the single-query exact path
    agree = all(np.array_equal(res[i], corpus.search(qh[i], top_k=k)) for i in range(4))
    flops = 2.0 * nq * rows * 256
    return {"workload": f"{rows}-line corpus, batch of {nq} queries, top-k={k} (BASELINE configs[2])",
            "value": nq / (ms * 1e-3), "unit": "queries/s", "ms_per_batch": ms, "dtype": "bf16 candidates + f64 exact re-rank",
            "gemm_TFLOPs_pipeline": flops / (ms * 1e-3) / 1e12, "queries_proven_exact": proven, "queries": nq,
            "gpu_launches_per_batch": launches, "shadow_build_ms": shadow_ms,
            "e2e": {"value": nq / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": nq * 1024,
                    "d2h_bytes_per_step": nq * (16 * k + 8), "ms_per_batch": e2e_s * 1e3},
            "agrees_with_single_query_path": bool(agree)}


def bench_batch_sharded(torch, dist, dev, ctx, stream, corpus, rows, k, world, nq=1024, iters=5):
    """configs[2] sharded row-wise: every rank runs K2 on its shard, the nq x k hits are
    all-gathered (NCCL, 160 KB per rank) and merged per query on every rank."""
    qh = gen_queries(nq + 64)[64:]
    q_dev = torch.from_numpy(qh).to(dev)
    hits = torch.zeros((nq, k, 2), dtype=torch.float64, device=dev)
    st = torch.zeros((nq, 2), dtype=torch.int32, device=dev)
    gathered = torch.zeros((world, nq, k, 2), dtype=torch.float64, device=dev)
    st_all = torch.zeros((world, nq, 2), dtype=torch.int32, device=dev)
    merged = torch.zeros((nq, k, 2), dtype=torch.float64, device=dev)
    corpus.prepare_batch()

    def one():
        corpus.search_batch_dev(q_dev.data_ptr(), nq, k, hits.data_ptr(), st.data_ptr())
        dist.all_gather_into_tensor(gathered, hits)
        dist.all_gather_into_tensor(st_all, st)
        ctx.hits_merge_batch_dev(gathered.data_ptr(), world, nq, k, k, merged.data_ptr())

    for _ in range(2):
        one()
    dist.barrier(); torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        one()
    e1.record(stream)
    dist.barrier(); torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ref = merged.clone()
    dist.broadcast(ref, src=0)
    agree = torch.tensor([int(torch.equal(ref.view(torch.int64), merged.view(torch.int64)))], device=dev)
    dist.all_reduce(agree, op=dist.ReduceOp.MIN)
    return {"workload": f"{rows}-line corpus row-sharded x{world}, batch of {nq} queries, top-k={k} (configs[2], sharded)",
            "value": nq / (ms * 1e-3), "unit": "queries/s", "ms_per_batch": ms,
            "gemm_TFLOPs_pipeline": 2.0 * nq * rows * 256 / (ms * 1e-3) / 1e12,
            "queries_proven_exact": int((st_all[:, :, 1].min(dim=0).values == 1).sum().item()), "queries": nq,
            "ranks_agree": bool(agree.item()), "exchange": "nccl all_gather of nq x k hits + stb_hits_merge_batch_dev"}


def bench_shadow_scan(torch, dev, ctx, stream, corpus, q_dev, k, rows, peak_gbs, steps=50):
    """K1 with STB_SCAN_SHADOW=1: candidate scores from the 16-bit normalised shadow (half the HBM
    bytes), exact f64 re-rank and proof as in the default path.  Same device-timed loop as
    `value`, plus a bit-for-bit comparison with the default (f32-row) scan on every query."""
    n_q = q_dev.shape[0]
    corpus.prepare_batch()                                   # builds the shadow if K2 has not already
    ref = torch.zeros((n_q, k, 2), dtype=torch.float64, device=dev)
    got = torch.zeros((n_q, k, 2), dtype=torch.float64, device=dev)
    st = torch.zeros((n_q, 4), dtype=torch.int32, device=dev)
    for i in range(n_q):
        corpus.search_topk_dev(q_dev[i].data_ptr(), k, ref[i].data_ptr(), st[i].data_ptr())
    torch.cuda.synchronize(dev)
    os.environ["STB_SCAN_SHADOW"] = "1"
    try:
        for i in range(n_q):
            corpus.search_topk_dev(q_dev[i].data_ptr(), k, got[i].data_ptr(), st[i].data_ptr())
        torch.cuda.synchronize(dev)
        proven = int((st[:, 1] == 1).sum().item())
        same = bool(torch.equal(ref.view(torch.int64), got.view(torch.int64)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            corpus.search_topk_dev(q_dev[i % n_q].data_ptr(), k, got[i % n_q].data_ptr(), st[i % n_q].data_ptr())
        e1.record(stream); torch.cuda.synchronize(dev)
        qh = q_dev.cpu().numpy()
        for i in range(3):
            corpus.search(qh[i], top_k=k)
        t0 = time.perf_counter()
        for i in range(steps):
            corpus.search(qh[i % n_q], top_k=k)              # host query in, host hits out (stb_search)
        e2e_s = time.perf_counter() - t0
    finally:
        os.environ.pop("STB_SCAN_SHADOW", None)
    ms = e0.elapsed_time(e1) / steps
    return {"workload": f"{rows}-line corpus, single query, top-k={k}, candidates from the 16-bit shadow (512 B/row)",
            "value": 1e3 / ms, "unit": "queries/s", "ms_per_step": ms, "steps": steps,
            "queries_proven_exact": proven, "queries": n_q, "bit_identical_to_f32_scan": same,
            "e2e": {"value": steps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": 1024, "d2h_bytes_per_step": 16 * k + 16},
            "roofline": {"bound": "hbm", "algorithmic_GBps": rows * 1024 / (ms * 1e-3) / 1e9,
                         "frac_of_measured_peak_algorithmic": rows * 1024 / (ms * 1e-3) / 1e9 / peak_gbs,
                         "bytes_read_per_query": rows * 512,
                         "frac_of_measured_peak_bytes_read": rows * 512 / (ms * 1e-3) / 1e9 / peak_gbs},
            "note": "opt-in (STB_SCAN_SHADOW=1); unproven queries fall back to the f32 scan in stb_search"}


# ------------------------------------------------------------------ K5 side bench -----
def bench_ivfpq(torch, dev, ctx, rows=4_000_000, nlist=4096, nprobe=64, n_centers=None, spread=0.6):
    """IVF-PQ (self-specified: the reference has no IVF_PQ, so no parity -- recall@10 against
    the exact scan is the quality metric).  Clustered synthetic corpus (random unit vectors
    have no neighbourhood structure for an IVF to exploit): rows = normalise(center + noise)."""
    from semtools_b200 import capi
    n_centers = n_centers or max(rows // 100, 1000)               # ~100 rows per natural cluster
    g = torch.Generator(device=dev); g.manual_seed(SEED + 5)
    centers = torch.randn((n_centers, 256), generator=g, device=dev); centers /= centers.norm(dim=1, keepdim=True)
    c = capi.Corpus(ctx, rows)
    for i in range(0, rows, CHUNK):
        n = min(CHUNK, rows - i)
        idx = torch.randint(0, n_centers, (n,), generator=g, device=dev)
        x = centers[idx] + spread / 16.0 * torch.randn((n, 256), generator=g, device=dev)
        x /= x.norm(dim=1, keepdim=True)
        torch.cuda.synchronize(dev); c.append_dev(x.data_ptr(), n)
    idx = torch.randint(0, n_centers, (64,), generator=g, device=dev)
    q = centers[idx] + spread / 16.0 * torch.randn((64, 256), generator=g, device=dev); q /= q.norm(dim=1, keepdim=True)
    qh = q.cpu().numpy()
    del x, centers
    t0 = time.perf_counter()
    index = capi.IvfPq(c, nlist=nlist, train_rows=262144, iters=8)
    build_s = time.perf_counter() - t0
    exact = [c.search(qh[i], top_k=10) for i in range(64)]
    t0 = time.perf_counter()
    for i in range(64):
        c.search(qh[i], top_k=10)
    exact_ms = (time.perf_counter() - t0) / 64 * 1e3
    rec, scanned = [], []
    t0 = time.perf_counter()
    for i in range(64):
        got, ns = index.search(qh[i], nprobe=nprobe, top_k=10, rerank=512)
        rec.append(len(set(got["row"].tolist()) & set(exact[i]["row"].tolist())) / 10.0); scanned.append(ns)
    ms = (time.perf_counter() - t0) / 64 * 1e3
    st = index.stats()
    index.close(); c.close()
    return {"workload": f"{rows} clustered rows, nlist={nlist}, nprobe={nprobe}, m=32x8bit, rerank=512, top-k=10",
            "parity": "unpinned (no IVF_PQ exists in the reference); quality = recall vs exact scan",
            "recall_at_10": float(np.mean(rec)), "min_recall": float(np.min(rec)), "build_s": build_s,
            "ms_per_query_e2e": ms, "exact_scan_ms_per_query_e2e": exact_ms, "scanned_rows_per_query": float(np.mean(scanned)),
            "code_bytes_per_query": float(np.mean(scanned)) * 32, "index_bytes": st["index_bytes"], "max_list": st["max_list"]}


def bench_ivfpq_sharded(torch, dist, dev, ctx, world, rank, rows=4_000_000, nlist=4096, nprobe=64, spread=0.6):
    """IVF-PQ sharded by ROW: the clustered corpus of `bench_ivfpq` split into `world`
    contiguous row blocks, one index per rank over its block (nlist lists each), per-rank
    top-k, all-gather of k hits, K4 merge (semtools_b200.sharded.ShardedCorpus.on_gpu_ivfpq).
    Off by default (--ivfpq-sharded): every rank takes part in collectives."""
    from semtools_b200 import capi
    from semtools_b200.sharded import ShardedCorpus, shard_bounds
    n_centers = max(rows // 100, 1000)
    lo, hi = shard_bounds(rows, world, rank)
    g = torch.Generator(device=dev); g.manual_seed(SEED + 5)                  # same centers on every rank
    centers = torch.randn((n_centers, 256), generator=g, device=dev); centers /= centers.norm(dim=1, keepdim=True)
    c = capi.Corpus(ctx, max(hi - lo, 1), row_base=lo)
    for chunk_id in range(lo // CHUNK, (hi + CHUNK - 1) // CHUNK):            # chunk-seeded: identical rows for any world size
        gc = torch.Generator(device=dev); gc.manual_seed(SEED + 1000 + chunk_id)
        idx = torch.randint(0, n_centers, (CHUNK,), generator=gc, device=dev)
        x = centers[idx] + spread / 16.0 * torch.randn((CHUNK, 256), generator=gc, device=dev)
        x /= x.norm(dim=1, keepdim=True)
        a, b = max(lo, chunk_id * CHUNK), min(hi, (chunk_id + 1) * CHUNK)
        part = x[a - chunk_id * CHUNK: b - chunk_id * CHUNK].contiguous()
        torch.cuda.synchronize(dev); c.append_dev(part.data_ptr(), b - a)
    gq = torch.Generator(device=dev); gq.manual_seed(SEED + 6)
    idx = torch.randint(0, n_centers, (64,), generator=gq, device=dev)
    q = centers[idx] + spread / 16.0 * torch.randn((64, 256), generator=gq, device=dev); q /= q.norm(dim=1, keepdim=True)
    qh = q.cpu().numpy()
    del x, centers
    t0 = time.perf_counter()
    index = capi.IvfPq(c, nlist=nlist, train_rows=262144, iters=8)
    build_s = time.perf_counter() - t0
    exact = ShardedCorpus.on_gpu(ctx, c, dist, dev)
    approx = ShardedCorpus.on_gpu_ivfpq(ctx, index, dist, dev, nprobe=nprobe, rerank=512)
    want = [exact.search(qh[i], 10) for i in range(64)]
    dist.barrier()
    t0 = time.perf_counter()
    got = [approx.search(qh[i], 10) for i in range(64)]
    dist.barrier()
    ms = (time.perf_counter() - t0) / 64 * 1e3
    rec = [len(set(got[i]["row"].tolist()) & set(want[i]["row"].tolist())) / 10.0 for i in range(64)]
    index.close(); c.close()
    return {"workload": f"{rows} clustered rows row-sharded x{world}, nlist={nlist} per shard, nprobe={nprobe}, rerank=512, top-k=10",
            "recall_at_10": float(np.mean(rec)), "min_recall": float(np.min(rec)), "ms_per_query_e2e": ms,
            "build_s": build_s, "exchange": "nccl all_gather of k hits + stb_hits_merge (host-staged, as ShardedCorpus.on_gpu)"}


# ------------------------------------------------------------------ K3 side bench -----
def bench_embed(torch, dev, ctx, stream, V=500_000, n_lines=1_000_000):
    """K3 on SURVEY 8d's synthetic ingestion batch: V=500k x 256 table (0.5 GB), line
    lengths ~ clamp(round(LogNormal(2.5,0.8)),0,2048), ids ~ Zipf(1.1).  Algorithmic
    bytes = sum_i (1028*T_i + 1024)."""
    from semtools_b200 import capi
    rng = np.random.default_rng(SEED + 77)
    E = (rng.standard_normal((V, 256), dtype=np.float32) * np.float32(0.1))
    T = np.clip(np.round(rng.lognormal(2.5, 0.8, n_lines)), 0, 2048).astype(np.int64)
    offsets = np.concatenate([[0], np.cumsum(T)]).astype(np.uint64)
    ids = ((rng.zipf(1.1, int(T.sum())) - 1) % V).astype(np.uint32)
    table = capi.Table(ctx, E)
    off_d = torch.from_numpy(offsets.view(np.int64)).to(dev)
    ids_d = torch.from_numpy(ids.view(np.int32)).to(dev)
    out_d = torch.empty((n_lines, 256), dtype=torch.float32, device=dev)
    torch.cuda.synchronize(dev)
    for _ in range(3):
        capi.embed_dev(ctx, table, off_d.data_ptr(), ids_d.data_ptr(), n_lines, out_d.data_ptr())
    capi.embed_status(ctx)
    iters = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        capi.embed_dev(ctx, table, off_d.data_ptr(), ids_d.data_ptr(), n_lines, out_d.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / iters
    alg_bytes = float((1028 * T + 1024).sum())
    corpus = capi.Corpus(ctx, n_lines)
    t0 = time.perf_counter()
    capi.embed(ctx, table, offsets, ids, out=False, append_to=corpus)       # host CSR in, rows stay in HBM
    e2e_s = time.perf_counter() - t0
    import oracle
    n_chk = 2000
    exp = oracle.embed_csr(E, offsets[:n_chk + 1], ids[:int(offsets[n_chk])])
    bit_exact = bool(np.array_equal(corpus.read(0, n_chk).view(np.uint32), exp.view(np.uint32)))
    corpus.close(); table.close()
    return {"kernel": "stb_embed_kernel", "lines": n_lines, "tokens": int(T.sum()), "table_rows": V,
            "ms": ms, "lines_per_s": n_lines / (ms * 1e-3), "achieved_GBps": alg_bytes / (ms * 1e-3) / 1e9,
            "algorithmic_bytes": alg_bytes, "bound": "hbm/l2 (random 1 KiB gathers, Zipf ids)",
            "e2e_lines_per_s": n_lines / e2e_s, "e2e_h2d_bytes": int(ids.nbytes + offsets.nbytes),
            "bit_exact_vs_oracle_first_2000": bit_exact}


# ------------------------------------------------------------------ our arm ----------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from semtools_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # a dedicated non-default stream shared by torch (events, NCCL ordering) and the library
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = capi.Context(local_rank, stream.cuda_stream)
    assert ctx.stream == stream.cuda_stream
    k = args.topk

    # ---- corpus shard of this rank (strong scaling: the SAME global corpus) --------
    per = (args.rows + world - 1) // world
    lo, hi = min(rank * per, args.rows), min((rank + 1) * per, args.rows)
    corpus = capi.Corpus(ctx, max(hi - lo, 1), row_base=lo)
    for c in range(lo // CHUNK, (max(hi, lo + 1) - 1) // CHUNK + 1):
        c_lo, c_hi = c * CHUNK, min((c + 1) * CHUNK, args.rows)
        a, b = max(lo, c_lo), min(hi, c_hi)
        if a >= b:
            continue
        x = gen_chunk_torch(torch, dev, c, c_hi - c_lo)
        torch.cuda.synchronize(dev)
        sl = x[a - c_lo:b - c_lo]
        corpus.append_dev(sl.data_ptr(), b - a)
        del x, sl
    torch.cuda.empty_cache()
    assert len(corpus) == hi - lo

    n_q = 64
    queries_h = gen_queries(n_q)
    q_dev = torch.from_numpy(queries_h).to(dev)
    n_slots = args.steps + args.warmup
    local_hits = torch.zeros((n_slots, k, 2), dtype=torch.float64, device=dev)   # stb_hit = (f64, u64)
    status = torch.zeros((n_slots, 4), dtype=torch.int32, device=dev)
    gathered = torch.zeros((world, k, 2), dtype=torch.float64, device=dev) if world > 1 else None
    final_hits = torch.zeros((n_slots, k, 2), dtype=torch.float64, device=dev)

    # ---- exchange wiring (N > 1) ---------------------------------------------------------
    xchg, exchange = None, "none"
    if world > 1:
        exchange = args.exchange
        if exchange == "p2p":
            try:
                xchg = capi.Exchange(ctx, world, rank, k)
                handles = [None] * world
                dist.all_gather_object(handles, xchg.local_handle())
                xchg.connect(handles)
                ok = torch.ones(1, device=dev)
            except capi.StbError as e:
                print(f"[rank {rank}] p2p exchange unavailable ({e}); using nccl", file=sys.stderr)
                ok = torch.zeros(1, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() == 0:
                xchg, exchange = None, "nccl"

    def step(i):
        if xchg is not None:       # ONE kernel: scan + NVLink peer-memory exchange + global merge
            xchg.search_topk(corpus, q_dev[i % n_q].data_ptr(), k, final_hits[i].data_ptr(), status[i].data_ptr())
            return
        corpus.search_topk_dev(q_dev[i % n_q].data_ptr(), k, local_hits[i].data_ptr(), status[i].data_ptr())
        if world > 1:
            dist.all_gather_into_tensor(gathered, local_hits[i])
            ctx.hits_merge_dev(gathered.data_ptr(), world, k, k, final_hits[i].data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- value: inputs resident in HBM, device-timed ---------------------------------
    # clock sampling starts before the warm-up (nvidia-smi needs ~0.2 s to come up) and runs
    # through the timed region; both are under the same load
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)
    for i in range(args.warmup):
        step(i)
    barrier()
    launches0 = ctx.counters()["kernel_launches"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for i in range(args.warmup, args.warmup + args.steps):
        step(i)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.counters()["kernel_launches"] - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    st = status[args.warmup:].cpu().numpy()
    n_expect = min(k, args.rows) if xchg is not None else min(k, hi - lo)
    all_complete = bool((st[:, 1] == 1).all() and (st[:, 0] == n_expect).all())

    ranks_agree = None
    if world > 1:
        mine = final_hits[args.warmup:args.warmup + min(args.steps, 16)].contiguous().view(torch.int64)
        ref = mine.clone()
        dist.broadcast(ref, src=0)
        agree = torch.tensor([int(torch.equal(ref, mine))], device=dev)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN)
        ranks_agree = bool(agree.item())

    # ---- e2e: host query in, host hits out, every step synchronous -------------------
    q_pin = torch.from_numpy(queries_h).pin_memory()
    out_pin = torch.zeros((k, 2), dtype=torch.float64).pin_memory()
    e2e_steps = max(10, min(args.steps, 100))

    def e2e_step(i):
        if world == 1:
            return corpus.search(queries_h[i % n_q], top_k=k)        # the C-ABI call a host makes
        if xchg is not None:
            return xchg.search(corpus, queries_h[i % n_q], k)[0]     # stb_search_xchg: the same call, sharded
        q_dev[i % n_q].copy_(q_pin[i % n_q], non_blocking=True)
        step(i % n_slots)
        out_pin.copy_(final_hits[i % n_slots], non_blocking=True)
        torch.cuda.synchronize(dev)
        return out_pin

    for i in range(3):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())

    # ---- parity spot-check of the benchmarked configuration (not timed) ----------------
    check = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle
        n_s = min(args.cpu_sample_rows, hi - lo)
        sample = corpus.read(0, n_s)
        cs = capi.Corpus(ctx, n_s)
        cs.append(sample)
        got = cs.search(queries_h[0], top_k=k)
        r, d = oracle.search_rows(sample, queries_h[0], top_k=k)
        check = bool(got["row"].tolist() == [int(x) for x in r] and np.array_equal(got["distance"], d))
        cs.close()

    # ---- K3 (embed gather/pool/normalise) secondary measurement, N=1 only ---------------
    def side(fn, *a, **kw):
        """Side sections never take the headline line down with them."""
        try:
            return fn(*a, **kw)
        except Exception as e:                                     # noqa: BLE001 - reported in the JSON line
            return {"error": f"{type(e).__name__}: {e}"}

    k3 = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        k3 = side(bench_embed, torch, dev, ctx, stream)

    # ---- K2 (BASELINE configs[2]: batch of 1024 queries, tensor-core path), N=1 only ------
    k2 = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        k2 = side(bench_batch, torch, dev, ctx, stream, corpus, args.rows, k)
    if world > 1 and not args.no_cpu_baseline:
        k2 = bench_batch_sharded(torch, dist, dev, ctx, stream, corpus, args.rows, k, world)   # every rank takes part: no side() here, a rank that swallowed an error would leave the others in the all-gather

    k5s = None
    if world > 1 and args.ivfpq_sharded:
        k5s = bench_ivfpq_sharded(torch, dist, dev, ctx, world, rank, rows=args.ivfpq_rows)

    # ---- K5 (BASELINE configs[4] at single-GPU scale: IVF-PQ probe, recall-measured) ------
    k5 = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        k5 = side(bench_ivfpq, torch, dev, ctx, rows=args.ivfpq_rows)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy, burst)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        ms_step = ms_max / args.steps
        rows_per_gpu = hi - lo
        # the dominant kernel is stb_scan_topk_kernel: algorithmic bytes = 1024 * rows it scans
        # (SURVEY 8d K1); its launch duration = step time when it is the only kernel (N=1).
        kernel_ms = ms_step
        achieved = rows_per_gpu * 1024 / (kernel_ms * 1e-3) / 1e9
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))["stb_scan_topk_kernel"]
            if tr["rows"] == rows_per_gpu and tr["top_k"] == k:
                traffic = tr["traffic_bytes"]     # from the committed ncu --set full capture
        except (OSError, KeyError, ValueError):
            pass
        line = {
            "metric": METRIC, "value": args.steps / (ms_max * 1e-3), "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args.rows, k), "rows": args.rows, "rows_per_gpu": rows_per_gpu,
                       "top_k": k, "parallelism": f"row-shard x{world}", "exchange": exchange,
                       "l2": "corpus shard >> 126 MB L2, no flush needed" if rows_per_gpu * 1024 > 4 * 126e6
                             else "WARNING shard fits partly in L2",
                       "distinct_queries": n_q},
            "clocks": clocks,
            "e2e": {"value": e2e_steps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": 1024,
                    "d2h_bytes_per_step": 16 * k + 16, "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3,
                    "timing": "wall clock around synchronous per-query host calls: pinned H2D of the query, kernel(s), "
                              "D2H of the hits, stream sync (N=1: stb_search; N>1: stb_search_xchg on every rank, or the "
                              "NCCL sequence with pinned copies when --exchange nccl)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "stb_scan_topk_kernel<E=1,U=2>", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": rows_per_gpu * 1024,
                         "note": "duration = CUDA-event step time / steps on the launching stream"
                                 + ("" if world == 1 else " (includes all-gather + merge, so a lower bound on the kernel)")},
            "all_results_proven_exact": all_complete,
            "ranks_agree": ranks_agree,
            "parity_spot_check": check,
        }
        if k2 is not None and "error" in k2:
            line["batch1024"] = k2
        elif k2 is not None:
            tpeak = float(peaks.get("bf16_tflops", 1590.0)) * world        # whole-job FLOP rate vs N GPUs' peak
            k2["roofline"] = {"bound": "tensor", "kernel": "stb_batch_gemm_kernel (tcgen05.mma kind::f16, bf16 in / f32 TMEM)",
                              "achieved": k2["gemm_TFLOPs_pipeline"], "peak": tpeak, "unit": "TFLOP/s",
                              "frac": k2["gemm_TFLOPs_pipeline"] / tpeak,
                              "peak_source": "MEASURED_PEAKS.json bf16_tflops (cuBLAS burst)" if "bf16_tflops" in peaks else "fallback 1590",
                              "note": "achieved = 2*Q*N*256 FLOP / whole-pipeline batch time (shadow(q)+GEMM+select+finish); "
                                      "the GEMM kernel alone is faster, see profiles/"}
            line["batch1024"] = k2
        if k5 is not None:
            line["ivfpq"] = k5
        if k5s is not None:
            line["ivfpq_sharded"] = k5s
        if k3 is not None and "error" in k3:
            line["k3_embed"] = k3
        elif k3 is not None:
            k3["frac"] = k3["achieved_GBps"] / peak
            line["k3_embed"] = k3
        if world == 1 and not args.no_cpu_baseline:
            n_s = min(args.cpu_sample_rows, rows_per_gpu)
            sample = corpus.read(0, n_s)
            v1, t1 = cpu_baseline(sample, queries_h, args.rows, k, 1)
            import oracle
            nt = oracle.baseline_threads()
            vN, tN = cpu_baseline(sample, queries_h, args.rows, k, nt)
            line["cpu_baseline"] = {
                "value": v1, "unit": "queries/s", "cores": 1, "kind": "port",
                "sample": f"first {n_s} of {args.rows} rows, 3 queries, median {t1:.3f} s, scaled x{args.rows / n_s:g} "
                          f"(linear in rows; favours the CPU, whose full-result sort is N log N)",
                "host_cores": os.cpu_count(),
                "all_cores_not_reference_behaviour": {"threads": nt, "value": vN}}
        if world == 1 and not args.no_cpu_baseline:
            # K2 pipeline v2 (sampled threshold -> emitting epilogue -> exact finish; opt-in in the
            # library until validated).  Measured LAST so nothing above depends on it.
            os.environ["STB_BATCH_V2"] = "1"
            try:
                v2 = side(bench_batch, torch, dev, ctx, stream, corpus, args.rows, k)
            finally:
                os.environ.pop("STB_BATCH_V2", None)
            if "error" not in v2:
                tpeak = float(peaks.get("bf16_tflops", 1590.0))
                v2["roofline_frac_pipeline"] = v2["gemm_TFLOPs_pipeline"] / tpeak
                v2["note"] = "opt-in pipeline (STB_BATCH_V2=1): same C-ABI call, same results; not the default path yet"
            line["batch1024_v2"] = v2
            line["k1_shadow_scan"] = side(bench_shadow_scan, torch, dev, ctx, stream, corpus, q_dev, k, args.rows, peak)
            # K5 fused search (two launches, one sync; opt-in STB_IVFPQ_V2=1), same workload as `ivfpq`
            os.environ["STB_IVFPQ_V2"] = "1"
            try:
                k5v2 = side(bench_ivfpq, torch, dev, ctx, rows=args.ivfpq_rows)
            finally:
                os.environ.pop("STB_IVFPQ_V2", None)
            if "error" not in k5v2:
                k5v2["note"] = ("opt-in fused search (STB_IVFPQ_V2=1); the exact-scan column inside this section "
                                "is unaffected by the switch")
            line["ivfpq_v2"] = k5v2
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
