"""Generates the committed golden fixtures in this directory.

No reference implementation can run here (Rust toolchain and crates absent), so
the fixtures come from (a) PUBLIC known-answer vectors and (b) the independent
numpy restatement oracle/oracle_np.py, which shares no code with the C oracle or
the CUDA kernels.  The C oracle and the GPU path are both tested against them.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle_np as onp  # noqa: E402

# (a) FNV-1a 64-bit known answers: the public FNV test-suite vectors
# (Fowler/Noll/Vo reference test_fnv.c, fnv1a_64 table) for the constants at
# reference src/workspace/store.rs:652-653.
FNV_KAT = {
    "": "cbf29ce484222325", "a": "af63dc4c8601ec8c", "b": "af63df4c8601f1a5",
    "c": "af63de4c8601eff2", "d": "af63d94c8601e773", "e": "af63d84c8601e5c0",
    "f": "af63db4c8601ead9", "fo": "08985907b541d342", "foo": "dcb27518fed9d577",
    "foob": "dd120e790c2512af", "fooba": "cac165afa2fef40a", "foobar": "85944171f73967e8",
}
for k, v in FNV_KAT.items():
    assert format(onp.fnv1a64(k.encode()), "016x") == v, (k, v)

ids = {
    "doc": {p: str(onp.fnv1a64(p.encode())) for p in ["/test/doc1.txt", "/test/doc2.txt", "a/ü.md"]},
    "line": {f"{p}|{n}": str(onp.line_id(p, n))
             for p in ["/test/doc1.txt", "/test/doc2.txt"] for n in [0, 1, 255, 256, 65536, -1]},
}
json.dump({"fnv1a64": FNV_KAT, "ids": ids}, open(os.path.join(HERE, "fnv1a64.json"), "w"), indent=1)

# (b) small seeded search / pooling cases scored by the numpy restatement
rng = np.random.default_rng(0x5E117001)
rows = rng.standard_normal((600, 256)).astype(np.float32)
rows /= np.linalg.norm(rows, axis=1, keepdims=True)
rows = rows.astype(np.float32)
rows[17] = rows[400]            # exact duplicates -> tie broken by row order
rows[401] = rows[400]
rows[33] = 0.0                  # empty line -> zero vector -> distance exactly 1
rows[34] = 0.0
q = (rows[400] + 0.35 * rng.standard_normal(256)).astype(np.float32)
doc_offsets = np.array([0, 1, 120, 120, 380, 600], dtype=np.uint64)   # incl. 1-line and empty docs
cases = {}
for name, kw in {
    "top3_n3": dict(n_lines=3, top_k=3),
    "top10_n0": dict(n_lines=0, top_k=10),
    "top40_n5": dict(n_lines=5, top_k=40),
    "thr_0p9": dict(n_lines=1, top_k=3, max_distance=0.9),
    "thr_0p0": dict(n_lines=1, top_k=3, max_distance=0.0),
    "thr_1p0000001": dict(n_lines=2, top_k=1, max_distance=1.0000001),
}.items():
    res = onp.search_documents(rows, doc_offsets, q, **kw)
    cases[name] = dict(kw=kw, res=[[float(t[0])] + [int(x) for x in t[1:]] for t in res])
np.savez_compressed(os.path.join(HERE, "search_small.npz"), rows=rows, q=q, doc_offsets=doc_offsets)
json.dump(cases, open(os.path.join(HERE, "search_small.json"), "w"))

E = (rng.standard_normal((600, 256)) * 0.1).astype(np.float32)
weights = rng.uniform(0.2, 1.5, 600).astype(np.float32)
mapping = rng.permutation(600).astype(np.uint32)
lines = [rng.integers(0, 600, t).astype(np.uint32) for t in [0, 1, 2, 7, 8, 9, 31, 64, 200]]
pooled = {}
for variant in ["plain", "weights", "mapping", "both", "nonorm"]:
    w = weights if variant in ("weights", "both") else None
    m = mapping if variant in ("mapping", "both") else None
    pooled[variant] = np.stack([onp.pool_ids(E, l, w, m, normalize=(variant != "nonorm")) for l in lines])
offsets = np.cumsum([0] + [len(l) for l in lines]).astype(np.uint64)
np.savez_compressed(os.path.join(HERE, "pool_small.npz"), E=E, weights=weights, mapping=mapping,
                    offsets=offsets, ids=np.concatenate(lines).astype(np.uint32),
                    **{f"out_{k}": v for k, v in pooled.items()})
print("golden fixtures written")
