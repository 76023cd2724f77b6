"""CPU suite, part 1: pin the oracle.

The C oracle (oracle/semtools_oracle.c) is checked against (a) public FNV-1a
known answers, (b) the committed golden fixtures scored by the independent numpy
restatement, (c) the fixtures the reference's own tests hold for this path
(src/workspace/store.rs:753-757,814-850; src/search/mod.rs:251-463 re-expressed
over synthetic embeddings).  Numeric parity with model2vec/simsimd themselves is
UNPINNED (no reference binary, no reference numeric vectors) -- see DESIGN.md.
"""
import json
import os

import numpy as np
import pytest

import oracle
from oracle import oracle_np as onp
from conftest import unit_rows

G = os.path.join(os.path.dirname(__file__), "golden")


def test_oracle_build_is_unfused():
    assert oracle.lib().orc_selftest_unfused() == 1


def test_fnv1a64_known_answers():
    kat = json.load(open(os.path.join(G, "fnv1a64.json")))
    for s, h in kat["fnv1a64"].items():
        assert format(oracle.fnv1a64(s.encode()), "016x") == h
    for p, h in kat["ids"]["doc"].items():
        assert str(oracle.doc_id(p)) == h
    for key, h in kat["ids"]["line"].items():
        p, n = key.rsplit("|", 1)
        assert str(oracle.line_id(p, int(n))) == h


def test_line_ids_differ_like_reference_test():
    # src/workspace/store.rs:1003-1022 asserts only inequality
    assert oracle.line_id("/test/doc1.txt", 0) != oracle.line_id("/test/doc1.txt", 1)
    assert oracle.line_id("/test/doc1.txt", 0) != oracle.line_id("/test/doc2.txt", 0)


def test_search_golden_cases():
    z = np.load(os.path.join(G, "search_small.npz"))
    cases = json.load(open(os.path.join(G, "search_small.json")))
    for name, c in cases.items():
        h = oracle.search_documents(z["rows"], z["doc_offsets"], z["q"], **c["kw"])
        exp = c["res"]
        assert len(h.row) == len(exp), name
        for i, (d, row, doc, idx, start, end) in enumerate(exp):
            assert (h.row[i], h.doc[i], h.match_line[i], h.start[i], h.end[i]) == (row, doc, idx, start, end), name
            assert h.distance[i] == d, name          # bit-identical f64


def test_pool_golden_cases():
    z = np.load(os.path.join(G, "pool_small.npz"))
    for variant in ["plain", "weights", "mapping", "both", "nonorm"]:
        w = z["weights"] if variant in ("weights", "both") else None
        m = z["mapping"] if variant in ("mapping", "both") else None
        out = oracle.embed_csr(z["E"], z["offsets"], z["ids"], w, m, normalize=(variant != "nonorm"))
        assert np.array_equal(out.view(np.uint32), z[f"out_{variant}"].view(np.uint32)), variant


def test_pool_empty_line_is_zero_vector():
    E = np.ones((4, 256), dtype=np.float32)
    assert np.all(oracle.pool_ids(E, []) == 0.0)
    with pytest.raises(IndexError):
        oracle.pool_ids(E, [4])


def test_prepare_ids_unk_and_truncate():
    ids = np.array([5, 1, 5, 2, 3, 5, 4], dtype=np.uint32)
    assert oracle.prepare_ids(ids, unk_id=5, max_length=3).tolist() == [1, 2, 3]
    assert oracle.prepare_ids(ids, unk_id=None, max_length=None).tolist() == ids.tolist()
    assert oracle.prepare_ids(ids, unk_id=5, max_length=2048).tolist() == [1, 2, 3, 4]


def test_cosine_special_cases():
    z = np.zeros(256, dtype=np.float32)
    a = np.full(256, 0.1, dtype=np.float32)
    assert oracle.cosine(z, z) == 0.0             # a2 == b2 == 0 -> 0
    assert oracle.cosine(a, z) == 1.0             # ab == 0 -> 1
    e0 = np.zeros(256, dtype=np.float32); e0[0] = 1
    e1 = np.zeros(256, dtype=np.float32); e1[1] = 1
    assert oracle.cosine(e0, e1) == 1.0
    assert oracle.cosine(a, a) == 0.0 or oracle.cosine(a, a) < 1e-15
    assert oracle.cosine(a, -a) == pytest.approx(2.0, abs=1e-15)
    with pytest.raises(ValueError):
        oracle.cosine(a, a[:100])                 # reference: None -> line skipped


def test_backend_choice_is_inside_tolerance():
    """simsimd dispatches to different accumulation widths; the canonical f64
    oracle and an f32-serial backend must agree far inside north_star's 1e-5."""
    rng = np.random.default_rng(3)
    rows = unit_rows(rng, 2000)
    q = unit_rows(rng, 1)[0]
    worst = max(abs(oracle.cosine(q, r) - oracle.cosine_serial32(q, r)) for r in rows)
    assert worst < 2e-6


def test_c_oracle_matches_numpy_restatement_random():
    rng = np.random.default_rng(11)
    rows = unit_rows(rng, 300)
    q = unit_rows(rng, 1)[0]
    d = oracle.distances(rows, q)
    for i in range(0, 300, 7):
        assert d[i] == onp.cosine(q, rows[i])
    offs = [0, 50, 50, 299, 300]
    for kw in [dict(n_lines=2, top_k=7), dict(n_lines=0, top_k=400), dict(n_lines=3, top_k=2, max_distance=0.95)]:
        h = oracle.search_documents(rows, offs, q, **kw)
        exp = onp.search_documents(rows, offs, q, **kw)
        assert [int(r) for r in h.row] == [t[1] for t in exp]
        assert [float(x) for x in h.distance] == [t[0] for t in exp]


# ---- reference test fixtures re-expressed (SURVEY 8c "fixtures to carry over") ----
def test_store_fixture_constant_vectors_path_filter():
    """src/workspace/store.rs:753-757 + :814-850: three collinear constant vectors,
    query [0.1;256] restricted to doc1, top_k=1, max_distance=0.1 -> doc1 line 0,
    distance < 0.1."""
    rows = np.stack([np.full(256, v, dtype=np.float32) for v in (0.1, 0.5, 0.75)])
    q = np.full(256, 0.1, dtype=np.float32)
    r, d = oracle.store_search(rows, [[0, 1]], q, top_k=1, max_distance=0.1)
    assert r.tolist() == [0] and d[0] < 0.1
    r, d = oracle.store_search(rows, [[0, 3]], q, top_k=3, max_distance=0.1)
    assert r.tolist() == [0, 1, 2]              # collinear: all ~0, ties by row
    assert oracle.store_search(rows, [], q, top_k=1)[0].size == 0       # :489 empty subset
    assert oracle.store_search(rows, [[0, 3]], q, top_k=0)[0].size == 0  # :489 top_k == 0


def test_search_structural_cases_from_reference_tests():
    rng = np.random.default_rng(5)
    rows = unit_rows(rng, 6)
    q = unit_rows(rng, 1)[0]
    # sorted ascending (:271-273)
    h = oracle.search_documents(rows, [0, 3, 5, 6], q, 3, 3)
    assert np.all(np.diff(h.distance) >= 0) and len(h.row) == 3
    # all < threshold (:290-292), threshold lifts top_k (:115-116)
    h = oracle.search_documents(rows, [0, 6], q, 3, 1, max_distance=1.5)
    assert np.all(h.distance < 1.5) and len(h.row) == 6
    # len <= top_k (:312)
    assert len(oracle.search_documents(rows, [0, 6], q, 3, 2).row) == 2
    # context 3 lines for n_lines=1 away from the edges (:331-334)
    h = oracle.search_documents(rows, [0, 6], q, 1, 6)
    for i in range(6):
        if 1 <= h.match_line[i] <= 4:
            assert h.end[i] - h.start[i] == 3
    # window clamped to a 2-line file (:350-356)
    h = oracle.search_documents(rows[:2], [0, 2], rows[0], 5, 1)
    assert (h.start[0], h.end[0]) == (0, 2)
    # empty documents -> empty (:389-390)
    assert len(oracle.search_documents(rows[:0], [0], q, 3, 3).row) == 0


def test_stable_tie_order_and_zero_rows():
    rng = np.random.default_rng(9)
    rows = unit_rows(rng, 50)
    rows[40] = rows[3]; rows[41] = rows[3]; rows[10] = 0; rows[5] = 0
    q = rows[3].copy()
    h = oracle.search_documents(rows, [0, 50], q, 0, 50)
    assert h.row[:3].tolist() == [3, 40, 41]
    d = dict(zip(h.row.tolist(), h.distance.tolist()))
    assert d[5] == 1.0 and d[10] == 1.0
    i5, i10 = h.row.tolist().index(5), h.row.tolist().index(10)
    assert i5 < i10


def test_cpu_baseline_agrees_with_oracle_on_indices():
    rng = np.random.default_rng(21)
    rows = unit_rows(rng, 5000)
    q = unit_rows(rng, 1)[0]
    r0, d0 = oracle.search_rows(rows, q, top_k=10)
    for threads in (1, 4):
        r1, d1 = oracle.baseline_search(rows, q, top_k=10, threads=threads)
        assert r1.tolist() == r0.tolist()
        assert np.max(np.abs(d1 - d0)) < 1e-5
