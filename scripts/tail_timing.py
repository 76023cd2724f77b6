"""Phase timestamps of K1 (needs a -DSTB_TAIL_TIMING library via STB_LIB_PATH)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semtools_b200 import capi
dev = torch.device("cuda:0")
s = torch.cuda.Stream(dev); torch.cuda.set_stream(s)
ctx = capi.Context(0, s.cuda_stream)
g = torch.Generator(device=dev); g.manual_seed(1)
for rows in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "65536,1250000").split(",")]:
    c = capi.Corpus(ctx, rows)
    for i in range(0, rows, 1_000_000):
        n = min(1_000_000, rows - i)
        x = torch.randn((n, 256), generator=g, device=dev); x /= x.norm(dim=1, keepdim=True)
        torch.cuda.synchronize(); c.append_dev(x.data_ptr(), n)
    q = torch.randn((4, 256), generator=g, device=dev); q /= q.norm(dim=1, keepdim=True)
    hits = torch.zeros((4, 10, 2), dtype=torch.float64, device=dev); st = torch.zeros((4, 4), dtype=torch.int32, device=dev)
    for i in range(5):
        c.search_topk_dev(q[i % 4].data_ptr(), 10, hits[i % 4].data_ptr(), st[i % 4].data_ptr())
    torch.cuda.synchronize()
    acc = []
    for rep in range(5):
        out = (C.c_uint64 * 8)()
        capi._check(capi.lib().stb_debug_timestamps(ctx._h, 1, out))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        c.search_topk_dev(q[rep % 4].data_ptr(), 10, hits[rep % 4].data_ptr(), st[rep % 4].data_ptr())
        e1.record(s); torch.cuda.synchronize()
        capi._check(capi.lib().stb_debug_timestamps(ctx._h, 0, out))
        t = [int(v) for v in out]
        acc.append([e0.elapsed_time(e1) * 1e3] + [(t[i] - t[0]) / 1e3 for i in range(1, 8)])
    a = np.median(np.array(acc), axis=0)
    print(f"rows {rows}: event {a[0]:.1f} us | since first CTA start: scan_end {a[1]:.1f}  cta_merge_end {a[2]:.1f}  ticket {a[3]:.1f}  keys_loaded {a[6]:.1f}  pass0_done {a[7]:.1f}  select_done {a[4]:.1f}  rerank_done {a[5]:.1f}")
    c.close()
