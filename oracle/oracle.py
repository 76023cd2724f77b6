"""ctypes front-end of oracle/semtools_oracle.c and oracle/cpu_baseline.c.

TEST INFRASTRUCTURE ONLY (see package docstring).  Each wrapper names the
reference lines the C function restates.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)


def build(force: bool = False) -> None:
    """Compile the oracle and the CPU baseline with the committed Makefile."""
    libs = [os.path.join(_BUILD, n) for n in ("liboracle.so", "libcpubaseline.so")]
    if force or not all(os.path.exists(p) for p in libs):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)


_lib = None
_base = None


def _ptr(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(os.path.join(_BUILD, "liboracle.so"))
        L.orc_fnv1a64.restype = C.c_uint64
        L.orc_fnv1a64.argtypes = [C.c_char_p, C.c_uint64]
        L.orc_doc_id.restype = C.c_uint64
        L.orc_doc_id.argtypes = [C.c_char_p, C.c_uint64]
        L.orc_line_id.restype = C.c_uint64
        L.orc_line_id.argtypes = [C.c_char_p, C.c_uint64, C.c_int32]
        L.orc_prepare_ids.restype = C.c_uint64
        L.orc_prepare_ids.argtypes = [u32p, C.c_uint64, C.c_int, C.c_uint32, C.c_int, C.c_uint64]
        L.orc_pool_ids.restype = C.c_int
        L.orc_pool_ids.argtypes = [f32p, C.c_uint64, C.c_uint32, f32p, C.c_uint64, u32p,
                                   C.c_uint64, C.c_int, u32p, C.c_uint64, f32p]
        L.orc_embed_csr.restype = C.c_int
        L.orc_embed_csr.argtypes = [f32p, C.c_uint64, C.c_uint32, f32p, C.c_uint64, u32p,
                                    C.c_uint64, C.c_int, u64p, u32p, C.c_uint64, f32p]
        L.orc_cosine_f32.restype = C.c_double
        L.orc_cosine_f32.argtypes = [f32p, f32p, C.c_uint64]
        L.orc_cosine_f32_serial32.restype = C.c_double
        L.orc_cosine_f32_serial32.argtypes = [f32p, f32p, C.c_uint64]
        L.orc_distances.restype = None
        L.orc_distances.argtypes = [f32p, C.c_uint64, C.c_uint32, f32p, f64p]
        L.orc_search_documents.restype = C.c_int
        L.orc_search_documents.argtypes = [f32p, C.c_uint64, C.c_uint32, u64p, C.c_uint64, f32p,
                                           C.c_uint64, C.c_uint64, C.c_int, C.c_double,
                                           C.c_uint64, u64p, u64p, u64p, u64p, u64p, f64p, u64p]
        L.orc_store_search.restype = C.c_int
        L.orc_store_search.argtypes = [f32p, C.c_uint64, C.c_uint32, u64p, C.c_uint64, f32p,
                                       C.c_uint64, C.c_int, C.c_float, C.c_uint64, u64p, f32p, u64p]
        L.orc_selftest_unfused.restype = C.c_int
        L.orc_num_threads.restype = C.c_int
        if L.orc_selftest_unfused() != 1:
            raise RuntimeError("oracle was built with FMA contraction; rebuild with -ffp-contract=off")
        _lib = L
    return _lib


def baseline_lib():
    global _base
    if _base is None:
        build()
        L = C.CDLL(os.path.join(_BUILD, "libcpubaseline.so"))
        L.orc_baseline_search.restype = C.c_int
        L.orc_baseline_search.argtypes = [f32p, C.c_uint64, f32p, C.c_uint64, C.c_int, C.c_double,
                                          C.c_int, C.c_uint64, u64p, f64p, u64p]
        L.orc_baseline_threads.restype = C.c_int
        L.orc_baseline_isa.restype = C.c_char_p
        _base = L
    return _base


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# -- a12 ---------------------------------------------------------------------
def fnv1a64(data: bytes) -> int:
    """fnv1a_hash, src/workspace/store.rs:651-661."""
    return int(lib().orc_fnv1a64(data, len(data)))


def doc_id(path: str) -> int:
    """DocMeta::id, src/workspace/store.rs:75-80."""
    b = path.encode("utf-8")
    return int(lib().orc_doc_id(b, len(b)))


def line_id(path: str, line_number: int) -> int:
    """LineEmbedding::id, src/workspace/store.rs:82-89."""
    b = path.encode("utf-8")
    return int(lib().orc_line_id(b, len(b), line_number))


# -- a3 / a4 / a5 ------------------------------------------------------------
def prepare_ids(ids, unk_id=None, max_length=None) -> np.ndarray:
    """retain(!= unk) then truncate(max_length): tail of encode_with_args
    (model2vec-rs 0.1.3; call sites src/search/mod.rs:69,138)."""
    a = np.ascontiguousarray(ids, dtype=np.uint32).copy()
    n = lib().orc_prepare_ids(_ptr(a, u32p), a.size, int(unk_id is not None),
                              int(unk_id or 0), int(max_length is not None), int(max_length or 0))
    return a[:n]


def pool_ids(E, ids, weights=None, mapping=None, normalize=True) -> np.ndarray:
    """StaticModel::pool_ids (model2vec-rs 0.1.3), see semtools_oracle.c."""
    E = _f32(E)
    ids = np.ascontiguousarray(ids, dtype=np.uint32)
    w = _f32(weights) if weights is not None else None
    m = np.ascontiguousarray(mapping, dtype=np.uint32) if mapping is not None else None
    out = np.empty(E.shape[1], dtype=np.float32)
    rc = lib().orc_pool_ids(_ptr(E, f32p), E.shape[0], E.shape[1], _ptr(w, f32p),
                            0 if w is None else w.size, _ptr(m, u32p), 0 if m is None else m.size,
                            int(normalize), _ptr(ids, u32p), ids.size, _ptr(out, f32p))
    if rc != 0:
        raise IndexError("token row outside the embedding table")
    return out


def embed_csr(E, offsets, ids, weights=None, mapping=None, normalize=True) -> np.ndarray:
    """encode_with_args minus tokenisation over a CSR batch (src/search/mod.rs:69)."""
    E = _f32(E)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    ids = np.ascontiguousarray(ids, dtype=np.uint32)
    w = _f32(weights) if weights is not None else None
    m = np.ascontiguousarray(mapping, dtype=np.uint32) if mapping is not None else None
    n = offsets.size - 1
    out = np.empty((n, E.shape[1]), dtype=np.float32)
    rc = lib().orc_embed_csr(_ptr(E, f32p), E.shape[0], E.shape[1], _ptr(w, f32p),
                             0 if w is None else w.size, _ptr(m, u32p), 0 if m is None else m.size,
                             int(normalize), _ptr(offsets, u64p), _ptr(ids, u32p), n, _ptr(out, f32p))
    if rc != 0:
        raise IndexError("token row outside the embedding table")
    return out


# -- a6 ----------------------------------------------------------------------
def cosine(a, b) -> float:
    """f32::cosine (simsimd; src/search/mod.rs:86), canonical f64 arithmetic."""
    a, b = _f32(a), _f32(b)
    if a.size != b.size:
        raise ValueError("length mismatch (reference returns None and skips the line)")
    return float(lib().orc_cosine_f32(_ptr(a, f32p), _ptr(b, f32p), a.size))


def cosine_serial32(a, b) -> float:
    a, b = _f32(a), _f32(b)
    return float(lib().orc_cosine_f32_serial32(_ptr(a, f32p), _ptr(b, f32p), a.size))


def distances(rows, q) -> np.ndarray:
    rows, q = _f32(rows), _f32(q)
    out = np.empty(rows.shape[0], dtype=np.float64)
    lib().orc_distances(_ptr(rows, f32p), rows.shape[0], rows.shape[1], _ptr(q, f32p), _ptr(out, f64p))
    return out


# -- a7 ----------------------------------------------------------------------
@dataclass
class Hits:
    row: np.ndarray          # global row = position in (doc order, line order)
    doc: np.ndarray
    match_line: np.ndarray   # 0-based line inside the doc (SearchResult.match_line)
    start: np.ndarray        # SearchResult.start
    end: np.ndarray          # SearchResult.end (exclusive)
    distance: np.ndarray     # f64


def search_documents(rows, doc_offsets, q, n_lines=3, top_k=3, max_distance=None) -> Hits:
    """search_documents, src/search/mod.rs:77-120."""
    rows, q = _f32(rows), _f32(q)
    n = rows.shape[0]
    D = rows.shape[1] if rows.ndim == 2 else q.size
    doc_offsets = np.ascontiguousarray(doc_offsets, dtype=np.uint64)
    n_docs = doc_offsets.size - 1
    cap = n if max_distance is not None else min(n, top_k)
    cap = max(cap, 1)
    o_row, o_doc, o_ml, o_s, o_e = (np.zeros(cap, dtype=np.uint64) for _ in range(5))
    o_d = np.zeros(cap, dtype=np.float64)
    o_n = C.c_uint64(0)
    rc = lib().orc_search_documents(_ptr(rows, f32p), n, D, _ptr(doc_offsets, u64p), n_docs,
                                    _ptr(q, f32p), n_lines, top_k, int(max_distance is not None),
                                    float(max_distance or 0.0), cap, _ptr(o_row, u64p),
                                    _ptr(o_doc, u64p), _ptr(o_ml, u64p), _ptr(o_s, u64p),
                                    _ptr(o_e, u64p), _ptr(o_d, f64p), C.byref(o_n))
    if rc != 0:
        raise RuntimeError(f"orc_search_documents rc={rc}")
    m = int(o_n.value)
    return Hits(o_row[:m], o_doc[:m], o_ml[:m], o_s[:m], o_e[:m], o_d[:m])


def search_rows(rows, q, top_k=3, max_distance=None):
    """search_documents over one implicit document: (rows, distances)."""
    rows = _f32(rows)
    h = search_documents(rows, [0, rows.shape[0]], q, 0, top_k, max_distance)
    return h.row, h.distance


# -- a11 ---------------------------------------------------------------------
def store_search(rows, ranges, q, top_k, max_distance=None):
    """Store::search_line_embeddings semantics, src/workspace/store.rs:481-546."""
    rows, q = _f32(rows), _f32(q)
    ranges = np.ascontiguousarray(ranges, dtype=np.uint64).reshape(-1, 2)
    cap = max(int(top_k), 1)
    o_row = np.zeros(cap, dtype=np.uint64)
    o_d = np.zeros(cap, dtype=np.float32)
    o_n = C.c_uint64(0)
    rc = lib().orc_store_search(_ptr(rows, f32p), rows.shape[0], rows.shape[1], _ptr(ranges, u64p),
                                ranges.shape[0], _ptr(q, f32p), top_k, int(max_distance is not None),
                                float(max_distance or 0.0), cap, _ptr(o_row, u64p), _ptr(o_d, f32p),
                                C.byref(o_n))
    if rc != 0:
        raise RuntimeError(f"orc_store_search rc={rc}")
    m = int(o_n.value)
    return o_row[:m], o_d[:m]


# -- timed baseline ----------------------------------------------------------
def baseline_search(rows, q, top_k=10, max_distance=None, threads=1):
    """The reference scan as a timed CPU baseline (cpu_baseline.c)."""
    rows, q = _f32(rows), _f32(q)
    assert rows.shape[1] == 256
    n = rows.shape[0]
    cap = n if max_distance is not None else max(min(n, top_k), 1)
    o_row = np.zeros(cap, dtype=np.uint64)
    o_d = np.zeros(cap, dtype=np.float64)
    o_n = C.c_uint64(0)
    rc = baseline_lib().orc_baseline_search(_ptr(rows, f32p), n, _ptr(q, f32p), top_k,
                                            int(max_distance is not None), float(max_distance or 0.0),
                                            threads, cap, _ptr(o_row, u64p), _ptr(o_d, f64p), C.byref(o_n))
    if rc != 0:
        raise RuntimeError(f"orc_baseline_search rc={rc}")
    m = min(int(o_n.value), cap)
    return o_row[:m], o_d[:m]


def baseline_isa() -> str:
    """SIMD kernel the timed baseline dispatches to on this host ("avx512f" or "avx2+fma")."""
    return baseline_lib().orc_baseline_isa().decode()


def baseline_threads() -> int:
    return int(baseline_lib().orc_baseline_threads())
