"""Workspace store: the reference's own unit tests (src/workspace/store.rs:717-1375,
src/workspace/mod.rs:104-312) re-expressed against semtools_b200.workspace.  Everything
except the nearest-neighbour query is host logic and runs on CPU; the query tests are
marked gpu (the store never computes a distance on the CPU)."""
import json
import os
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

import numpy as np
import pytest

from semtools_b200 import capi
from semtools_b200.search import SearchConfig
from semtools_b200.workspace import (CURRENT_EMBEDDING_VERSION, DocMeta, LineEmbedding, Store, Workspace,
                                     WorkspaceConfig, _rust_lines, search_with_workspace)


def meta(path, size=1024, mtime=1234567890, version=CURRENT_EMBEDDING_VERSION):
    return DocMeta(path, size, mtime, version)


def sample_lines():
    """store.rs:753-757: constant 256-d vectors 0.1 / 0.5 / 0.75."""
    return [LineEmbedding(f"/test/doc{i + 1}.txt", 0, np.full(256, v, dtype=np.float32))
            for i, v in enumerate((0.1, 0.5, 0.75))]


def sample_docs():
    return [meta("/test/doc1.txt", 1024), meta("/test/doc2.txt", 2048), meta("/test/doc3.txt", 512, 1234567891)]


# ------------------------------------------------------------------ mod.rs tests ---------
def test_workspace_config_default():                     # mod.rs:112-119
    c = WorkspaceConfig()
    assert (c.name, c.root_dir, c.in_batch_size, c.oversample_factor) == ("default", "", 5000, 3)


def test_workspace_active_and_paths(monkeypatch, tmp_path):   # mod.rs:121-219
    monkeypatch.setenv("HOME", str(tmp_path))
    monkeypatch.delenv("SEMTOOLS_WORKSPACE", raising=False)
    with pytest.raises(RuntimeError, match="No active workspace"):
        Workspace.active()
    monkeypatch.setenv("SEMTOOLS_WORKSPACE", "")
    with pytest.raises(RuntimeError, match="No active workspace"):
        Workspace.active()
    monkeypatch.setenv("SEMTOOLS_WORKSPACE", "test-workspace")
    assert Workspace.active() == "test-workspace"
    assert Workspace.active("flag-wins") == "flag-wins"
    assert Workspace.root_path("w").endswith(os.path.join(".semtools", "workspaces", "w"))
    assert Workspace.active_path().endswith(os.path.join(".semtools", "workspaces", "test-workspace"))
    assert Workspace._config_path_for("w").endswith(os.path.join("workspaces", "w", "config.json"))


def test_workspace_save_and_open_roundtrip(monkeypatch, tmp_path):   # mod.rs:222-312
    monkeypatch.setenv("HOME", str(tmp_path))
    monkeypatch.setenv("SEMTOOLS_WORKSPACE", "ws1")
    ws = Workspace.open()                                 # no config on disk -> defaults filled in
    assert ws.config.name == "ws1" and ws.config.root_dir == Workspace.root_path("ws1")
    ws.config.in_batch_size = 1000
    ws.config.oversample_factor = 5
    ws.save()
    on_disk = json.load(open(Workspace._config_path_for("ws1")))
    assert set(on_disk) == {"name", "root_dir", "in_batch_size", "oversample_factor"}
    again = Workspace.open()
    assert (again.config.in_batch_size, again.config.oversample_factor) == (1000, 5)
    open(Workspace._config_path_for("ws1"), "w").write("{not json")
    assert Workspace.open().config.in_batch_size == 5000  # unreadable config -> defaults (mod.rs:36-39)


# ------------------------------------------------------------------ store.rs tests -------
def test_store_creation_and_stats(tmp_path):              # store.rs:763-812
    s = Store.open(str(tmp_path))
    assert os.path.isdir(os.path.join(str(tmp_path), "flat.b200"))
    st = s.get_stats()
    assert st.total_documents == 0 and st.has_index
    s.upsert_document_metadata(sample_docs())
    assert s.get_stats().total_documents == 3 and s.count_documents() == 3


def test_get_all_document_paths_and_existing_docs(tmp_path):   # store.rs:853-915
    s = Store.open(str(tmp_path))
    s.upsert_document_metadata(sample_docs())
    assert sorted(s.get_all_document_paths()) == ["/test/doc1.txt", "/test/doc2.txt", "/test/doc3.txt"]
    ex = s.get_existing_docs(["/test/doc1.txt", "/test/doc3.txt", "/nonexistent.txt"])
    assert len(ex) == 2 and ex["/test/doc1.txt"].size_bytes == 1024 and ex["/test/doc3.txt"].size_bytes == 512


def test_delete_documents(tmp_path):                      # store.rs:918-948
    s = Store.open(str(tmp_path))
    s.upsert_document_metadata(sample_docs())
    s.upsert_line_embeddings(sample_lines())
    s.delete_documents(["/test/doc1.txt", "/test/doc2.txt"])
    assert s.get_all_document_paths() == ["/test/doc3.txt"]
    assert s.count_line_embeddings() == 1
    s.delete_documents([])                                # :361-363 no-op


def test_delete_metadata_only_for_current_version(tmp_path):   # store.rs:263-271 filter on _version
    s = Store.open(str(tmp_path))
    s.upsert_document_metadata([meta("/old.txt", version=1), meta("/new.txt")])
    s.delete_document_metadata(["/old.txt", "/new.txt"])
    assert s.get_all_document_paths() == ["/old.txt"]


def test_upsert_replaces_by_id(tmp_path):                 # store.rs:951-1000
    s = Store.open(str(tmp_path))
    s.upsert_document_metadata([meta("/test/doc1.txt", 1024)])
    s.upsert_document_metadata([meta("/test/doc1.txt", 4096, 1234567999)])
    assert s.count_documents() == 1 and s.get_existing_docs(["/test/doc1.txt"])["/test/doc1.txt"].size_bytes == 4096
    s.upsert_line_embeddings(sample_lines())
    le = LineEmbedding("/test/doc1.txt", 0, np.full(256, 0.9, dtype=np.float32))
    s.upsert_line_embeddings([le, LineEmbedding("/test/doc1.txt", 1, np.full(256, 0.2, dtype=np.float32))])
    assert s.count_line_embeddings() == 4                 # one replaced, one appended
    assert np.all(s._emb[0] == np.float32(0.9))


def test_ids(tmp_path):                                   # store.rs:1003-1022 (+ FNV known answer)
    a, b = LineEmbedding("/test/doc1.txt", 0), LineEmbedding("/test/doc1.txt", 1)
    assert a.id() != b.id() and a.id() != LineEmbedding("/test/doc2.txt", 0).id()
    assert DocMeta("a", 0, 0).id() == 0xAF63DC4C8601EC8C


def test_persistence_roundtrip(tmp_path):
    s = Store.open(str(tmp_path))
    s.upsert_document_metadata(sample_docs())
    s.upsert_line_embeddings(sample_lines())
    t = Store.open(str(tmp_path))
    assert t.count_documents() == 3 and t.count_line_embeddings() == 3
    assert np.array_equal(t._emb, s._emb) and t._paths == s._paths
    raw = np.fromfile(os.path.join(str(tmp_path), "flat.b200", "line_embeddings.f32"), dtype=np.float32)
    assert raw.size == 3 * 256                            # the file IS the corpus matrix


def test_analyze_document_states(tmp_path):               # store.rs:1044-1287 (6 cases)
    s = Store.open(str(tmp_path / "ws"))
    f1, f2, f3 = (tmp_path / n for n in ("a.txt", "b.txt", "c.txt"))
    f1.write_text("hello\nworld\n"); f2.write_text("x\n"); f3.write_text("old\n")
    states = s.analyze_document_states([str(f1), str(f2), str(tmp_path / "missing.txt")])
    assert [st.kind for st in states] == ["New", "New"]    # missing file skipped (:578-581)
    assert states[0].info.content == "hello\nworld\n" and states[0].info.meta.size_bytes == 12
    s.upsert_document_metadata([st.info.meta for st in states])
    assert [st.kind for st in s.analyze_document_states([str(f1), str(f2)])] == ["Unchanged", "Unchanged"]
    f2.write_text("changed content\n")                     # size differs
    assert [st.kind for st in s.analyze_document_states([str(f1), str(f2)])] == ["Unchanged", "Changed"]
    st3 = os.stat(f3)
    s.upsert_document_metadata([DocMeta(str(f3), st3.st_size, int(st3.st_mtime), 1)])   # version mismatch
    assert s.analyze_document_states([str(f3)])[0].kind == "Changed"
    s.upsert_document_metadata([DocMeta(str(f3), st3.st_size, int(st3.st_mtime) - 5)])  # mtime differs
    assert s.analyze_document_states([str(f3)])[0].kind == "Changed"
    assert s.analyze_document_states([]) == []


def test_rust_lines_semantics():                          # search/mod.rs:55 str::lines()
    assert _rust_lines("") == []
    assert _rust_lines("a\nb") == ["a", "b"]
    assert _rust_lines("a\r\nb\n") == ["a", "b"]
    assert _rust_lines("a\n\nb\n\n") == ["a", "", "b", ""]
    assert _rust_lines("\n") == [""]
    # a bare CR is text, not a line break; only "\r\n" / "\n" terminate; an unterminated last
    # line keeps its trailing CR (Python text mode used to split on the bare CR)
    assert _rust_lines("one\rstill one\ntwo\r\nthree") == ["one\rstill one", "two", "three"]
    assert _rust_lines("a\r") == ["a\r"]
    assert _rust_lines("a\r\n") == ["a"]
    assert _rust_lines("a\r\r\n") == ["a\r"]


def test_read_to_string_does_not_translate_newlines(tmp_path):
    from semtools_b200.workspace import read_to_string
    p = tmp_path / "cr.txt"
    p.write_bytes(b"one\rstill one\ntwo\r\nthree")
    assert read_to_string(str(p)) == "one\rstill one\ntwo\r\nthree"
    assert len(_rust_lines(read_to_string(str(p)))) == 3
    bad = tmp_path / "bad.txt"
    bad.write_bytes(b"\xff\xfe")
    with pytest.raises(UnicodeDecodeError):
        read_to_string(str(bad))


# ------------------------------------------------------------------ GPU query tests ------
@pytest.mark.gpu
def test_search_line_embeddings_reference_fixture(tmp_path, ctx):   # store.rs:814-850
    s = Store.open(str(tmp_path), ctx)
    s.upsert_line_embeddings(sample_lines())
    q = np.full(256, 0.1, dtype=np.float32)
    res = s.search_line_embeddings(q, ["/test/doc1.txt"], 1, 0.1)
    assert len(res) == 1 and res[0].line_number == 0 and res[0].path == "/test/doc1.txt" and res[0].distance < 0.1
    assert s.search_line_embeddings(q, [], 1, 0.1) == []
    assert s.search_line_embeddings(q, ["/test/doc1.txt"], 0, 0.1) == []
    assert s.search_line_embeddings(q, ["/nope.txt"], 3, None) == []


@pytest.mark.gpu
def test_search_with_workspace_flow_matches_oracle(tmp_path, ctx, monkeypatch):
    """search/mod.rs:146-216 end to end with a synthetic embedder: only New/Changed
    files are (re-)embedded, interleaved upserts keep the path filter exact."""
    import oracle
    monkeypatch.setenv("HOME", str(tmp_path))
    monkeypatch.setenv("SEMTOOLS_WORKSPACE", "wsx")
    rng = np.random.default_rng(0)
    vocab = {}

    def embed_lines(lines):
        calls.append(len(lines))
        out = np.zeros((len(lines), 256), dtype=np.float32)
        for i, l in enumerate(lines):
            if l not in vocab:
                v = rng.standard_normal(256).astype(np.float32)
                vocab[l] = v / np.linalg.norm(v)
            out[i] = vocab[l]
        return out

    files = []
    for d in range(4):
        p = tmp_path / f"doc{d}.txt"
        p.write_text("".join(f"doc{d} line {i}\n" for i in range(30 + d)))
        files.append(str(p))
    calls = []
    q = embed_lines(["doc2 line 7"])[0]
    calls.clear()
    cfg = SearchConfig(n_lines=3, top_k=3)
    res = search_with_workspace(files, q, embed_lines, cfg, ctx=ctx)
    assert sum(calls) == 30 + 31 + 32 + 33
    assert (res[0].path, res[0].line_number) == (files[2], 7) and res[0].distance < 1e-6
    calls.clear()
    res2 = search_with_workspace(files[1:3], q, embed_lines, cfg, ctx=ctx)     # unchanged: nothing re-embedded
    assert calls == [] and [(r.path, r.line_number) for r in res2][0] == (files[2], 7)
    time.sleep(1.1)
    open(files[0], "a").write("doc2 line 7\n")                                 # changed file gains the query line
    res3 = search_with_workspace(files, q, embed_lines, cfg, ctx=ctx)
    assert calls == [31]
    # exact tie: ordered by store row (the appended line of doc0 is a NEW id -> last row)
    assert [(r.path, r.line_number) for r in res3[:2]] == [(files[2], 7), (files[0], 30)]
    # cross-check against the oracle over the store's own matrix + path filter
    st = Store.open(Workspace.open().config.root_dir, ctx)
    ranges = st._ranges_for(files[1:3])
    r, d = oracle.store_search(st._emb, ranges, q, 3)
    got = st.search_line_embeddings(q, files[1:3], 3)
    assert [(g.path, g.line_number) for g in got] == [(st._paths[st._rows[int(x)][0]], int(st._rows[int(x)][1])) for x in r]
    assert [np.float32(g.distance) for g in got] == list(d)


def test_incremental_flush_appends_and_patches_in_place(tmp_path):
    """Upserts append to / patch the row files; only deletions rewrite them.  The files always
    equal what a from-scratch write of the same state produces."""
    def emb(v):
        return np.full(256, v, dtype=np.float32)

    s = Store.open(str(tmp_path))
    s.upsert_line_embeddings([LineEmbedding("a.txt", i, emb(i)) for i in range(4)])
    assert s.full_rewrites == 1                                    # first flush creates the files
    s = Store.open(str(tmp_path))
    s.upsert_line_embeddings([LineEmbedding("b.txt", i, emb(10 + i)) for i in range(3)])      # append
    s.upsert_line_embeddings([LineEmbedding("a.txt", 2, emb(99.0)), LineEmbedding("c.txt", 0, emb(20))])   # patch + append
    s.upsert_document_metadata([DocMeta("a.txt", 1, 2)])           # metadata only
    assert s.full_rewrites == 0
    d = tmp_path / "flat.b200"
    rows = np.fromfile(d / "rows.i32", dtype=np.int32).reshape(-1, 2)
    e = np.fromfile(d / "line_embeddings.f32", dtype=np.float32).reshape(-1, 256)
    assert rows.tolist() == [[0, 0], [0, 1], [0, 2], [0, 3], [1, 0], [1, 1], [1, 2], [2, 0]]
    assert e[:, 0].tolist() == [0, 1, 99, 3, 10, 11, 12, 20]
    s2 = Store.open(str(tmp_path))                                 # a fresh load sees the same state
    assert np.array_equal(s2._emb, s._emb) and np.array_equal(s2._rows, s._rows) and s2._paths == s._paths
    s2.delete_documents(["b.txt"])                                 # rows move -> rewrite
    assert s2.full_rewrites == 1
    s3 = Store.open(str(tmp_path))
    assert s3.count_line_embeddings() == 5 and s3._emb[:, 0].tolist() == [0, 1, 99, 3, 20]
    # deletions commit a NEW generation of row files; the old one is gone
    meta = json.load(open(d / "store.json"))
    assert meta["files"]["rows"] != "rows.i32" and not (d / "rows.i32").exists() and meta["rows"] == 5
    # debris behind the committed rows (an interrupted flush) is cut off, not appended to
    with open(d / meta["files"]["rows"], "ab") as f:
        f.write(b"\x07" * 8)
    s3.upsert_line_embeddings([LineEmbedding("d.txt", 0, emb(30))])
    assert s3.full_rewrites == 0
    s4 = Store.open(str(tmp_path))
    assert s4.count_line_embeddings() == 6 and s4._emb[:, 0].tolist() == [0, 1, 99, 3, 20, 30]
    assert s4._rows[-1].tolist() == [s4._path_idx["d.txt"], 0]


def test_store_is_crash_consistent_and_validates_what_it_loads(tmp_path):
    """rows reach the files before store.json is replaced, so a crash in between must
    leave a store that opens (the uncommitted rows are ignored), and a row file that lost
    committed rows or points outside the path table must fail with a clear error, not IndexError."""
    def emb(v):
        return np.full(256, v, dtype=np.float32)
    s = Store.open(str(tmp_path))
    s.upsert_line_embeddings([LineEmbedding("a.txt", i, emb(i)) for i in range(3)])
    s.upsert_document_metadata([DocMeta("a.txt", 1, 2)])
    d = tmp_path / "flat.b200"
    # "crash" after appending two rows of a NEW path (index 1) but before the commit
    with open(d / "rows.i32", "ab") as f:
        f.write(np.array([[1, 0], [1, 1]], dtype=np.int32).tobytes())
    with open(d / "line_embeddings.f32", "ab") as f:
        f.write(np.stack([emb(7), emb(8)]).tobytes())
    t = Store.open(str(tmp_path))
    assert t.count_line_embeddings() == 3 and t._paths == ["a.txt"]
    t.upsert_line_embeddings([LineEmbedding("b.txt", 0, emb(50))])
    u = Store.open(str(tmp_path))
    assert u._emb[:, 0].tolist() == [0, 1, 2, 50] and u._paths == ["a.txt", "b.txt"]
    assert os.path.getsize(d / "rows.i32") == 4 * 8
    # committed rows missing -> clear error
    with open(d / "rows.i32", "r+b") as f:
        f.truncate(2 * 8)
    with pytest.raises(RuntimeError, match="truncated"):
        Store.open(str(tmp_path))
    # path index outside the table -> clear error
    np.array([[0, 0], [0, 1], [0, 2], [9, 0]], dtype=np.int32).tofile(d / "rows.i32")
    with pytest.raises(RuntimeError, match="corrupt"):
        Store.open(str(tmp_path))


def test_two_writers_do_not_drop_each_others_rows(tmp_path):
    """two `search -w` processes used to overwrite each other.  Every mutation now runs
    under an exclusive flock and first reloads when the on-disk generation moved."""
    def emb(v):
        return np.full(256, v, dtype=np.float32)
    a, b = Store.open(str(tmp_path)), Store.open(str(tmp_path))        # both loaded the same (empty) state
    a.upsert_line_embeddings([LineEmbedding("a.txt", i, emb(i)) for i in range(3)])
    a.upsert_document_metadata([DocMeta("a.txt", 1, 2)])
    b.upsert_line_embeddings([LineEmbedding("b.txt", i, emb(10 + i)) for i in range(2)])   # stale handle: reloads first
    b.upsert_document_metadata([DocMeta("b.txt", 3, 4)])
    a.upsert_line_embeddings([LineEmbedding("a.txt", 1, emb(77))])      # stale again: patch lands on the merged state
    c = Store.open(str(tmp_path))
    assert sorted(c.get_all_document_paths()) == ["a.txt", "b.txt"]
    assert c._emb[:, 0].tolist() == [0, 77, 2, 10, 11]
    assert [c._paths[i] for i in c._rows[:, 0]] == ["a.txt"] * 3 + ["b.txt"] * 2
    # separate processes, same protocol
    import subprocess, sys, textwrap
    code = textwrap.dedent("""
        import sys, numpy as np
        sys.path.insert(0, %r)
        from semtools_b200.workspace import Store, LineEmbedding
        s = Store.open(%r)
        tag = int(sys.argv[1])
        for j in range(20):
            s.upsert_line_embeddings([LineEmbedding("p%%d.txt" %% tag, j, np.full(256, tag * 100 + j, dtype=np.float32))])
    """) % (ROOT, str(tmp_path))
    procs = [subprocess.Popen([sys.executable, "-c", code, str(t)]) for t in (1, 2, 3)]
    assert all(p.wait(timeout=120) == 0 for p in procs)
    e = Store.open(str(tmp_path))
    assert e.count_line_embeddings() == 5 + 60
    vals = sorted(e._emb[5:, 0].tolist())
    assert vals == sorted(t * 100 + j for t in (1, 2, 3) for j in range(20))


def test_store_records_the_embedder_and_re_embeds_foreign_vectors(tmp_path):
    """the C++ host (WordLevel stand-in) and the Python host (HF tokenizer) share one store;
    change detection alone (size, mtime, _version) would silently mix their vectors."""
    f = tmp_path / "doc.txt"; f.write_text("hello\nworld\n")
    st = os.stat(f)
    a = Store.open(str(tmp_path), model_fingerprint="model2vec:A")
    a.upsert_line_embeddings([LineEmbedding(str(f), 0, np.ones(256, np.float32))])
    a.upsert_document_metadata([DocMeta(str(f), st.st_size, int(st.st_mtime))])
    assert json.load(open(tmp_path / "flat.b200" / "store.json"))["model"] == "model2vec:A"
    same = Store.open(str(tmp_path), model_fingerprint="model2vec:A")
    assert [s.kind for s in same.analyze_document_states([str(f)])] == ["Unchanged"]
    other = Store.open(str(tmp_path), model_fingerprint="wordlevel:B")
    assert [s.kind for s in other.analyze_document_states([str(f)])] == ["Changed"]
    anon = Store.open(str(tmp_path))                                  # no fingerprint given: reference behaviour
    assert [s.kind for s in anon.analyze_document_states([str(f)])] == ["Unchanged"]
    other.upsert_document_metadata([DocMeta(str(f), st.st_size, int(st.st_mtime))])    # re-embedded by B: the store is B's now
    assert json.load(open(tmp_path / "flat.b200" / "store.json"))["model"] == "wordlevel:B"
