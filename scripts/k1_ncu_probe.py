"""One tier of K1 for ncu: python scripts/k1_ncu_probe.py <f32|h16|q8> [rows] [launches]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from semtools_b200 import capi
tier = sys.argv[1]
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
n = int(sys.argv[3]) if len(sys.argv) > 3 else 6
os.environ["STB_SCAN_TIER"] = tier
dev = torch.device("cuda:0")
s = torch.cuda.Stream(dev); torch.cuda.set_stream(s)
ctx = capi.Context(0, s.cuda_stream)
g = torch.Generator(device=dev); g.manual_seed(1)
c = capi.Corpus(ctx, rows)
for i in range(0, rows, 1_000_000):
    m = min(1_000_000, rows - i)
    x = torch.randn((m, 256), generator=g, device=dev); x /= x.norm(dim=1, keepdim=True)
    torch.cuda.synchronize(); c.append_dev(x.data_ptr(), m)
del x
c.prepare()
q = torch.randn((n, 256), generator=g, device=dev); q /= q.norm(dim=1, keepdim=True)
hits = torch.zeros((n, 10, 2), dtype=torch.float64, device=dev); st = torch.zeros((n, 4), dtype=torch.int32, device=dev)
for i in range(n):
    c.search_topk_dev(q[i].data_ptr(), 10, hits[i].data_ptr(), st[i].data_ptr())
torch.cuda.synchronize()
print(tier, rows, st.cpu().numpy()[:, :2].tolist())
