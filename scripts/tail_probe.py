"""Small-corpus probe of K1 (fixed overhead / tail).  python scripts/tail_probe.py [rows] [iters]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semtools_b200 import capi
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0")
s = torch.cuda.Stream(dev); torch.cuda.set_stream(s)
ctx = capi.Context(0, s.cuda_stream)
g = torch.Generator(device=dev); g.manual_seed(1)
x = torch.randn((rows, 256), generator=g, device=dev); x /= x.norm(dim=1, keepdim=True)
c = capi.Corpus(ctx, rows); torch.cuda.synchronize(); c.append_dev(x.data_ptr(), rows)
if os.environ.get("STB_PROBE_PREPARE", "1") == "1":  # build the q8 / h16 copies (STB_SCAN_TIER picks the tier that is scanned)
    c.prepare()
q = torch.randn((16, 256), generator=g, device=dev); q /= q.norm(dim=1, keepdim=True)
hits = torch.zeros((16, 10, 2), dtype=torch.float64, device=dev); st = torch.zeros((16, 4), dtype=torch.int32, device=dev)
for i in range(iters):
    c.search_topk_dev(q[i % 16].data_ptr(), 10, hits[i % 16].data_ptr(), st[i % 16].data_ptr())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s)
for i in range(iters):
    c.search_topk_dev(q[i % 16].data_ptr(), 10, hits[i % 16].data_ptr(), st[i % 16].data_ptr())
e1.record(s); torch.cuda.synchronize()
h32 = None
if os.environ.get("STB_SCAN_TIER", "q8") != "f32":   # cross-check the reduced-width scan against the f32 scan
    got = hits.clone(); os.environ["STB_SCAN_TIER"] = "f32"
    for i in range(16):
        c.search_topk_dev(q[i].data_ptr(), 10, hits[i].data_ptr(), st[i].data_ptr())
    torch.cuda.synchronize()
    h32 = bool(torch.equal(got.view(torch.int64), hits.view(torch.int64)))
print("rows", rows, "us/query", e0.elapsed_time(e1) / iters * 1e3, "complete", bool((st[:, 1] == 1).all()), "same_as_f32_scan", h32)
