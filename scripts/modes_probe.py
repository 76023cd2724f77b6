"""Times the non-headline modes of stb_search end to end (host query in, host hits out):
store query with row ranges, threshold mode, large top_k, fallback.  python scripts/modes_probe.py [rows]"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semtools_b200 import capi
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device("cuda:0")
ctx = capi.Context(0)
g = torch.Generator(device=dev); g.manual_seed(9)
c = capi.Corpus(ctx, rows)
for i in range(0, rows, 1_000_000):
    n = min(1_000_000, rows - i)
    x = torch.randn((n, 256), generator=g, device=dev); x /= x.norm(dim=1, keepdim=True)
    torch.cuda.synchronize(); c.append_dev(x.data_ptr(), n)
q = torch.randn((8, 256), generator=g, device=dev); q /= q.norm(dim=1, keepdim=True)
qh = q.cpu().numpy()
c.prepare()                                  # q8 + h16 copies; STB_SCAN_TIER=f32 in the environment pins the f32 rows

def timeit(fn, reps=8):
    fn(0); t0 = time.perf_counter()
    for i in range(reps): r = fn(i % 8)
    return (time.perf_counter() - t0) / reps * 1e3, r

out = {"rows": rows, "tier": os.environ.get("STB_SCAN_TIER", "q8 (default)")}
ms, r = timeit(lambda i: c.search(qh[i], top_k=10)); out["topk10_ms"] = ms
ms, r = timeit(lambda i: c.search(qh[i], top_k=10, mode=capi.STB_MODE_STORE_QUERY, row_ranges=[[0, rows]])); out["store_1_range_ms"] = ms
rng = np.random.default_rng(0)
starts = np.sort(rng.choice(np.arange(0, rows - 200, 200), 20_000, replace=False))
ranges = np.stack([starts, starts + rng.integers(50, 200, len(starts))], axis=1).astype(np.uint64)
sel = int((ranges[:, 1] - ranges[:, 0]).sum())
ms, r = timeit(lambda i: c.search(qh[i], top_k=10, mode=capi.STB_MODE_STORE_QUERY, row_ranges=ranges)); out["store_20k_ranges_ms"] = ms; out["store_20k_ranges_rows"] = sel
d10 = c.search(qh[0], top_k=1000)["distance"]
thr = float(d10[-1])
ms, r = timeit(lambda i: c.search(qh[0], top_k=3, max_distance=thr)); out["threshold_ms"] = ms; out["threshold_hits"] = len(r)
ms, r = timeit(lambda i: c.search(qh[i], top_k=1000), reps=4); out["topk1000_ms"] = ms
ms, r = timeit(lambda i: c.search(qh[i], top_k=96), reps=4); out["topk96_ms"] = ms
print(json.dumps(out))
