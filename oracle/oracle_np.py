"""Independent numpy restatement of the same maths, used ONLY to cross-check
semtools_oracle.c in tests (two independently written oracles must agree
bit for bit).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import numpy as np


def fnv1a64(data: bytes) -> int:
    """src/workspace/store.rs:651-661."""
    h = 0xCBF29CE484222325
    for b in data:
        h ^= b
        h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def line_id(path: str, line_number: int) -> int:
    """src/workspace/store.rs:82-89 (i32 little-endian suffix)."""
    return fnv1a64(path.encode("utf-8") + int(line_number).to_bytes(4, "little", signed=True))


def pool_ids(E, ids, weights=None, mapping=None, normalize=True):
    """model2vec-rs 0.1.3 pool_ids: sequential f32, unfused mul/add."""
    E = np.asarray(E, dtype=np.float32)
    s = np.zeros(E.shape[1], dtype=np.float32)
    cnt = 0
    for tok in np.asarray(ids, dtype=np.uint32):
        tok = int(tok)
        row = int(mapping[tok]) if mapping is not None and tok < len(mapping) else tok
        scale = np.float32(weights[tok]) if weights is not None and tok < len(weights) else np.float32(1.0)
        s = (s + (E[row] * scale).astype(np.float32)).astype(np.float32)
        cnt += 1
    s = (s / np.float32(max(cnt, 1))).astype(np.float32)
    if normalize:
        ss = np.float32(0.0)
        sq = (s * s).astype(np.float32)
        for v in sq:                      # sequential fold, f32
            ss = np.float32(ss + v)
        norm = max(np.float32(np.sqrt(ss)), np.float32(1e-12))
        s = (s / norm).astype(np.float32)
    return s


def cosine(a, b) -> float:
    """Canonical: f64 accumulation in index order (math.fsum would NOT be it)."""
    a = np.asarray(a, dtype=np.float32).astype(np.float64)
    b = np.asarray(b, dtype=np.float32).astype(np.float64)
    ab = a2 = b2 = 0.0
    for x, y in zip(a.tolist(), b.tolist()):
        ab += x * y
        a2 += x * x
        b2 += y * y
    if a2 == 0.0 and b2 == 0.0:
        return 0.0
    if ab == 0.0:
        return 1.0
    r = 1.0 - ab / (np.sqrt(a2) * np.sqrt(b2))
    return float(r) if r > 0.0 else 0.0


def search_documents(rows, doc_offsets, q, n_lines=3, top_k=3, max_distance=None):
    """src/search/mod.rs:77-120 with Python's stable sort."""
    thr = 100.0 if max_distance is None else max_distance
    res = []
    for doc in range(len(doc_offsets) - 1):
        lo, hi = int(doc_offsets[doc]), int(doc_offsets[doc + 1])
        for idx in range(hi - lo):
            d = cosine(q, rows[lo + idx])
            if d < thr:
                start = max(0, idx - n_lines)
                end = min(hi - lo, idx + n_lines + 1)
                res.append((d, lo + idx, doc, idx, start, end))
    res.sort(key=lambda t: t[0])          # stable
    return res if max_distance is not None else res[:top_k]
