"""K5 probe: build + recall/time grid.  python scripts/ivfpq_probe.py [rows] [nlist] [n_centers] [spread]"""
import os, sys, json, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semtools_b200 import capi
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
nlist = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
n_centers = int(sys.argv[3]) if len(sys.argv) > 3 else 20_000
spread = float(sys.argv[4]) if len(sys.argv) > 4 else 0.6
dev = torch.device("cuda:0")
ctx = capi.Context(0)
g = torch.Generator(device=dev); g.manual_seed(3)
centers = torch.randn((n_centers, 256), generator=g, device=dev); centers /= centers.norm(dim=1, keepdim=True)
c = capi.Corpus(ctx, rows)
for i in range(0, rows, 1_000_000):
    n = min(1_000_000, rows - i)
    idx = torch.randint(0, n_centers, (n,), generator=g, device=dev)
    x = centers[idx] + spread / 16.0 * torch.randn((n, 256), generator=g, device=dev)
    x /= x.norm(dim=1, keepdim=True)
    torch.cuda.synchronize(); c.append_dev(x.data_ptr(), n)
idx = torch.randint(0, n_centers, (64,), generator=g, device=dev)
q = centers[idx] + spread / 16.0 * torch.randn((64, 256), generator=g, device=dev); q /= q.norm(dim=1, keepdim=True)
qh = q.cpu().numpy()
t0 = time.perf_counter(); index = capi.IvfPq(c, nlist=nlist, train_rows=min(rows, 262144), iters=8); ctx.sync()
print(json.dumps({"rows": rows, "nlist": nlist, "build_s": time.perf_counter() - t0, **index.stats()}))
exact = [c.search(qh[i], top_k=10) for i in range(len(qh))]
t0 = time.perf_counter()
for i in range(len(qh)): c.search(qh[i], top_k=10)
t_exact = (time.perf_counter() - t0) / len(qh)
for nprobe in (8, 32, 64, 128):
    if nprobe > nlist: continue
    for rerank in (128, 512, 2048):
        rec, scanned = [], []
        t0 = time.perf_counter()
        for i in range(len(qh)):
            got, ns = index.search(qh[i], nprobe=nprobe, top_k=10, rerank=rerank)
            rec.append(len(set(got["row"].tolist()) & set(exact[i]["row"].tolist())) / 10.0); scanned.append(ns)
        dt = (time.perf_counter() - t0) / len(qh)
        print(json.dumps({"nprobe": nprobe, "rerank": rerank, "recall@10": float(np.mean(rec)), "min_recall": float(np.min(rec)),
                          "scanned_rows": float(np.mean(scanned)), "ms_per_query": dt * 1e3, "exact_ms_per_query": t_exact * 1e3}))
