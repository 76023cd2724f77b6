"""K2 timing probe: python scripts/batch_probe.py [rows] [nq] [iters]"""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semtools_b200 import capi
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
k = 10
dev = torch.device("cuda:0")
s = torch.cuda.Stream(dev); torch.cuda.set_stream(s)
ctx = capi.Context(0, s.cuda_stream)
g = torch.Generator(device=dev); g.manual_seed(1)
c = capi.Corpus(ctx, rows)
for i in range(0, rows, 1_000_000):
    n = min(1_000_000, rows - i)
    x = torch.randn((n, 256), generator=g, device=dev); x /= x.norm(dim=1, keepdim=True)
    torch.cuda.synchronize(); c.append_dev(x.data_ptr(), n)
del x
q = torch.randn((nq, 256), generator=g, device=dev); q /= q.norm(dim=1, keepdim=True)
hits = torch.zeros((nq, k, 2), dtype=torch.float64, device=dev); st = torch.zeros((nq, 2), dtype=torch.int32, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s); c.prepare_batch(); e1.record(s); torch.cuda.synchronize()
print("shadow build ms", e0.elapsed_time(e1))
for _ in range(2):
    c.search_batch_dev(q.data_ptr(), nq, k, hits.data_ptr(), st.data_ptr())
torch.cuda.synchronize()
l0 = ctx.counters()["kernel_launches"]
e0.record(s)
for _ in range(iters):
    c.search_batch_dev(q.data_ptr(), nq, k, hits.data_ptr(), st.data_ptr())
e1.record(s); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
flops = 2.0 * nq * rows * 256
# cross-check 4 queries against the single-query path
h1 = torch.zeros((k, 2), dtype=torch.float64, device=dev); s1 = torch.zeros(4, dtype=torch.int32, device=dev)
agree = True
for i in range(4):
    c.search_topk_dev(q[i].data_ptr(), k, h1.data_ptr(), s1.data_ptr()); torch.cuda.synchronize()
    agree &= bool(torch.equal(h1.view(torch.int64), hits[i].view(torch.int64)))
print(json.dumps({"rows": rows, "nq": nq, "ms_per_batch": ms, "qps": nq / ms * 1e3, "TFLOPs": flops / ms / 1e9,
                  "complete": int((st[:, 1] == 1).sum().item()), "agree_with_k1": agree,
                  "launches_per_batch": (ctx.counters()["kernel_launches"] - l0) / iters}))
