"""Times the K1 scan kernel of every variant library (scripts/build_variants.sh).
Each variant runs in its own process (one .so per process).  GPU only.
    python scripts/scan_sweep.py [rows] [out.jsonl]
"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, os, sys, numpy as np, torch
sys.path.insert(0, os.environ["STB_ROOT"])
from semtools_b200 import capi
rows = int(sys.argv[1]); k = int(sys.argv[2])
dev = torch.device("cuda:0")
s = torch.cuda.Stream(dev); torch.cuda.set_stream(s)
ctx = capi.Context(0, s.cuda_stream)
g = torch.Generator(device=dev); g.manual_seed(1)
c = capi.Corpus(ctx, rows)
for i in range(0, rows, 1_000_000):
    n = min(1_000_000, rows - i)
    x = torch.randn((n, 256), generator=g, device=dev); x /= x.norm(dim=1, keepdim=True)
    torch.cuda.synchronize(); c.append_dev(x.data_ptr(), n)
q = torch.randn((16, 256), generator=g, device=dev); q /= q.norm(dim=1, keepdim=True)
hits = torch.zeros((16, k, 2), dtype=torch.float64, device=dev); st = torch.zeros((16, 4), dtype=torch.int32, device=dev)
def run(n):
    for i in range(n):
        c.search_topk_dev(q[i % 16].data_ptr(), k, hits[i % 16].data_ptr(), st[i % 16].data_ptr())
run(5); torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s); run(30); e1.record(s); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 30)
ok = bool((st[:, 1] == 1).all().item())
print(json.dumps({"lib": os.path.basename(os.environ["STB_LIB_PATH"]), "rows": rows, "k": k, "ms": best,
                  "GBps": rows * 1024 / best / 1e6, "complete": ok, "row0": int(hits[0, 0, 1].view(torch.int64).item())}))
'''


def main():
    rows_list = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [4_000_000]
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "sweep.jsonl")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    libs = sorted(glob.glob(os.path.join(ROOT, "semtools_b200", "lib", "variants", "*.so")))
    with open(out, "a") as f:
        for lib in libs:
            for rows in rows_list:
                k = 10
                env = dict(os.environ, STB_LIB_PATH=lib, STB_ROOT=ROOT)
                r = subprocess.run([sys.executable, "-c", CHILD, str(rows), str(k)], env=env, capture_output=True,
                                   text=True, timeout=300)
                line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else json.dumps(
                    {"lib": os.path.basename(lib), "error": r.stderr[-400:]})
                print(line)
                f.write(line + "\n")
                f.flush()


if __name__ == "__main__":
    main()
