"""K3 probe on SURVEY 8d's ingestion batch: python scripts/embed_probe.py [lines] [V]"""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semtools_b200 import capi
n_lines = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
V = int(sys.argv[2]) if len(sys.argv) > 2 else 500_000
dev = torch.device("cuda:0")
s = torch.cuda.Stream(dev); torch.cuda.set_stream(s)
ctx = capi.Context(0, s.cuda_stream)
rng = np.random.default_rng(7)
E = (rng.standard_normal((V, 256), dtype=np.float32) * np.float32(0.1))
T = np.clip(np.round(rng.lognormal(2.5, 0.8, n_lines)), 0, 2048).astype(np.int64)
offsets = np.concatenate([[0], np.cumsum(T)]).astype(np.uint64)
ids = ((rng.zipf(1.1, int(T.sum())) - 1) % V).astype(np.uint32)
table = capi.Table(ctx, E)
off_d = torch.from_numpy(offsets.view(np.int64)).to(dev); ids_d = torch.from_numpy(ids.view(np.int32)).to(dev)
out_d = torch.empty((n_lines, 256), dtype=torch.float32, device=dev)
torch.cuda.synchronize()
for _ in range(3):
    capi.embed_dev(ctx, table, off_d.data_ptr(), ids_d.data_ptr(), n_lines, out_d.data_ptr())
capi.embed_status(ctx)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s)
for _ in range(10):
    capi.embed_dev(ctx, table, off_d.data_ptr(), ids_d.data_ptr(), n_lines, out_d.data_ptr())
e1.record(s); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
alg = float((1028 * T + 1024).sum())
print(json.dumps({"lines": n_lines, "tokens": int(T.sum()), "ms": ms, "lines_per_s": n_lines / ms * 1e3, "GBps": alg / ms / 1e6}))
