"""Command layer + host model side (reference src/cmds/*.rs, src/json_mode.rs,
src/bin/semtools.rs, model2vec-rs encode front half).  Formatting, flags, tokenisation and
the workspace commands are host logic (CPU); the end-to-end search tests need the GPU."""
import io
import json
import os

import numpy as np
import pytest

import oracle
from semtools_b200 import cmds
from semtools_b200.__main__ import build_parser
from semtools_b200.model import StaticModel
from semtools_b200.search import RankedLine, SearchResult

WORDS = ["alpha", "beta", "gamma", "delta", "epsilon", "zeta", "eta", "theta", "iota", "kappa", "lambda", "mu",
         "hello", "world", "goodbye", "test", "line", "query", "apple", "banana", "orange", "grape", "fruit"]


@pytest.fixture(scope="module")
def model_dir(tmp_path_factory):
    """A tiny local model2vec directory: WordLevel tokenizer + random 256-d table."""
    from safetensors.numpy import save_file
    from tokenizers import Tokenizer
    from tokenizers.models import WordLevel
    from tokenizers.normalizers import Lowercase
    from tokenizers.pre_tokenizers import Whitespace
    d = tmp_path_factory.mktemp("model")
    vocab = {"[UNK]": 0, **{w: i + 1 for i, w in enumerate(WORDS)}}
    tok = Tokenizer(WordLevel(vocab, unk_token="[UNK]"))
    tok.pre_tokenizer = Whitespace()
    tok.save(str(d / "tokenizer.json"))
    rng = np.random.default_rng(0)
    E = (rng.standard_normal((len(vocab), 256)) * 0.1).astype(np.float32)
    save_file({"embeddings": E}, str(d / "model.safetensors"))
    (d / "config.json").write_text(json.dumps({"normalize": True}))
    return str(d), E, vocab


# ------------------------------------------------------------------ formatting (a14) --------
def test_rust_float_display():
    assert cmds.rust_display_f64(0.0) == "0" and cmds.rust_display_f64(1.0) == "1"
    assert cmds.rust_display_f64(0.1 + 0.2) == "0.30000000000000004"
    assert cmds.rust_display_f64(1e-7) == "0.0000001"                     # never scientific
    assert cmds.rust_display_f64(2.220446049250313e-16) == "0.0000000000000002220446049250313"
    assert cmds.rust_display_f32(np.float32(0.1)) == "0.1"                # f32 shortest, not 0.10000000149
    assert cmds.rust_display_f32(np.float32(1.0) - np.float32(0.9)) == "0.100000024"


def test_print_search_results_format():                    # cmds/search.rs:35-63
    r = SearchResult("dir/a.txt", ["l2", "l3", "l4"], 1, 4, 2, 0.25)
    assert cmds.format_search_results([r], False) == "dir/a.txt:1::4 (0.25)\n   2: l2\n   3: l3\n   4: l4\n\n"
    tty = cmds.format_search_results([r], True)
    assert "\x1b[43m\x1b[30m   3: l3\x1b[0m\n" in tty and "   2: l2\n" in tty
    big = SearchResult("f", ["x"], 12344, 12345, 12344, 1.0)
    assert cmds.format_search_results([big], False) == "f:12344::12345 (1)\n12345: x\n\n"      # {:4} is a minimum width


def test_serde_json_pretty_shapes():                       # json_mode.rs + serde_json::to_string_pretty
    r = SearchResult("a\"b.txt", ["x", "y\tz"], 0, 2, 1, 1e-7)
    s = cmds.to_string_pretty({"results": [cmds.search_result_to_json(r)]})
    assert json.loads(s)["results"][0] == {"filename": "a\"b.txt", "start_line_number": 0, "end_line_number": 2,
                                           "match_line_number": 1, "distance": 1e-7, "content": "x\ny\tz"}
    assert '"distance": 1e-7,' in s and s.startswith('{\n  "results": [\n    {\n      "filename": "a\\"b.txt",')
    assert cmds.to_string_pretty({"results": []}) == '{\n  "results": []\n}'
    assert cmds.to_string_pretty({"error": "m", "error_type": "NoInput"}) == '{\n  "error": "m",\n  "error_type": "NoInput"\n}'
    assert cmds._json_f64(1.0) == "1.0" and cmds._json_f64(0.00001) == "0.00001" and cmds._json_f64(0.000001) == "1e-6"
    assert cmds._json_f64(1e15) == "1000000000000000.0" and cmds._json_f64(1e16) == "1e16" and cmds._json_f64(1.5e16) == "1.5e16"   # ryu switches at 1e16


def test_workspace_result_rendering(tmp_path):             # cmds/search.rs:66-110,208-237
    f = tmp_path / "doc.txt"
    f.write_text("a\nb\nc\nd\n")
    rl = RankedLine(str(f), 3, float(np.float32(0.1)))
    txt = cmds.format_workspace_search_results([rl], 2, False)
    assert txt == f"{f}:1::6 (0.1)\n   2: b\n   3: c\n   4: d\n\n"       # end NOT clamped in the header
    js = cmds.workspace_results_to_json([rl], 2)[0]
    assert js["end_line_number"] == 6 and js["content"] == "b\nc\nd" and js["distance"] == float(np.float32(0.1))
    gone = RankedLine(str(tmp_path / "gone.txt"), 0, 0.5)
    assert "    [Error: Could not read file content]\n" in cmds.format_workspace_search_results([gone], 1, False)


def test_cli_flags_and_aliases():                          # src/bin/semtools.rs:52-83
    p = build_parser()
    a = p.parse_args(["search", "q", "f1", "f2"])
    assert (a.n_lines, a.top_k, a.max_distance, a.ignore_case, a.json, a.workspace) == (3, 3, None, False, False, None)
    a = p.parse_args(["search", "q", "--context", "5", "--threshold", "0.4", "-i", "-j", "-w", "ws", "--top-k", "7"])
    assert (a.n_lines, a.max_distance, a.ignore_case, a.json, a.workspace, a.top_k, a.files) == (5, 0.4, True, True, "ws", 7, [])
    a = p.parse_args(["search", "q", "-n", "1", "-m", "0.2"])
    assert (a.n_lines, a.max_distance) == (1, 0.2)
    assert p.parse_args(["workspace", "use", "w1"]).name == "w1"
    assert p.parse_args(["workspace", "--json", "status"]).json and p.parse_args(["workspace", "prune", "-j"]).json_sub


def test_workspace_commands(tmp_path, monkeypatch):        # cmds/workspace.rs
    monkeypatch.setenv("HOME", str(tmp_path))
    monkeypatch.delenv("SEMTOOLS_WORKSPACE", raising=False)
    out = io.StringIO()
    cmds.workspace_use_cmd("w1", False, out)
    assert out.getvalue().startswith("Workspace 'w1' configured.\nTo activate it, run:\n  export SEMTOOLS_WORKSPACE=w1\n\n")
    out = io.StringIO(); cmds.workspace_use_cmd("w1", True, out)
    assert json.loads(out.getvalue()) == {"name": "w1", "root_dir": os.path.join(str(tmp_path), ".semtools", "workspaces", "w1"),
                                          "total_documents": 0}
    with pytest.raises(RuntimeError):
        cmds.workspace_status_cmd(False, None)             # "No active workspace"
    out = io.StringIO(); cmds.workspace_status_cmd(False, "w1", out)
    assert out.getvalue().splitlines()[0] == "Active workspace: w1" and "Documents: 0" in out.getvalue()
    from semtools_b200.workspace import DocMeta, Store, Workspace
    st = Store.open(Workspace.root_path("w1"))
    keep = tmp_path / "keep.txt"; keep.write_text("x\n")
    st.upsert_document_metadata([DocMeta(str(keep), 2, 1), DocMeta(str(tmp_path / "gone.txt"), 2, 1)])
    out = io.StringIO(); cmds.workspace_prune_cmd(False, "w1", out)
    assert out.getvalue() == f"Found 1 stale documents:\n  - {tmp_path / 'gone.txt'}\nRemoved 1 stale documents from workspace.\n"
    out = io.StringIO(); cmds.workspace_prune_cmd(True, "w1", out)
    assert json.loads(out.getvalue()) == {"files_removed": 0, "files_remaining": 1}
    out = io.StringIO(); cmds.workspace_prune_cmd(False, "w1", out)
    assert out.getvalue() == "No stale documents found. Workspace is clean.\n"


def test_search_cmd_no_input_error():                      # cmds/search.rs:178-192
    out, err = io.StringIO(), io.StringIO()
    assert cmds.search_cmd("q", [], 3, 3, None, False, True, None, None, stdin_is_tty=True, out=out, err=err) == 1
    assert json.loads(err.getvalue()) == {"error": "No input provided. Either specify files as arguments or pipe input to stdin.",
                                          "error_type": "NoInput"}
    err = io.StringIO()
    cmds.search_cmd("q", [], 3, 3, None, False, False, None, None, stdin_is_tty=True, out=out, err=err)
    assert err.getvalue() == "Error: No input provided. Either specify files as arguments or pipe input to stdin.\n"


# ------------------------------------------------------------------ host model side (a1, a3, a5) ---
def test_model_loading_and_tokenisation(model_dir):
    d, E, vocab = model_dir
    m = StaticModel.from_pretrained(d)
    assert m.normalize and m.unk_token_id == 0 and np.array_equal(m.embeddings, E)
    lens = sorted(len(t) for t in vocab)
    assert m.median_token_length == lens[len(lens) // 2]
    off, ids = m.tokenize(["hello world", "", "zzz hello zzz", "apple banana orange"], 2048)
    assert off.tolist() == [0, 2, 2, 3, 6]                 # unk dropped, empty line -> no tokens
    assert ids.tolist() == [vocab["hello"], vocab["world"], vocab["hello"], vocab["apple"], vocab["banana"], vocab["orange"]]
    off, ids = m.tokenize(["alpha beta gamma delta"], 2)   # truncate(max_length) after unk removal
    assert ids.tolist() == [vocab["alpha"], vocab["beta"]]
    assert StaticModel.truncate_str("abcdef", 2, 2) == "abcd"
    with pytest.raises(FileNotFoundError):
        StaticModel.from_pretrained("minishlab/potion-multilingual-128M")


# ------------------------------------------------------------------ end to end (GPU) -------------------
@pytest.mark.gpu
def test_search_cmd_files_text_and_json(model_dir, tmp_path, ctx, monkeypatch):
    d, E, vocab = model_dir
    monkeypatch.delenv("SEMTOOLS_WORKSPACE", raising=False)
    model = StaticModel.from_pretrained(d, ctx=ctx)
    f1 = tmp_path / "file1.txt"; f1.write_text("hello world\ngoodbye world\ntest line\n\napple banana\n")
    f2 = tmp_path / "file2.txt"; f2.write_text("orange grape\r\nfruit fruit fruit\r\nHELLO world")
    files = [str(f1), str(f2)]
    # expected through the oracle: same tokenisation, oracle pooling, oracle search_documents
    lines = [["hello world", "goodbye world", "test line", "", "apple banana"], ["orange grape", "fruit fruit fruit", "HELLO world"]]
    flat = [l for doc in lines for l in doc]
    off, ids = model.tokenize(flat, 2048)
    rows = oracle.embed_csr(E, off, ids)
    qo, qi = model.tokenize(["apple fruit"], 512)
    q = oracle.pool_ids(E, qi)
    h = oracle.search_documents(rows, [0, 5, 8], q, n_lines=1, top_k=3)
    exp = ""
    for i in range(len(h.row)):
        doc, s, e, ml = int(h.doc[i]), int(h.start[i]), int(h.end[i]), int(h.match_line[i])
        exp += f"{files[doc]}:{s}::{e} ({cmds.rust_display_f64(float(h.distance[i]))})\n"
        exp += "".join(f"{s + j + 1:4}: {lines[doc][s + j]}\n" for j in range(e - s)) + "\n"
    out = io.StringIO()
    assert cmds.search_cmd("apple fruit", files, 1, 3, None, False, False, None, model, out=out) == 0
    assert out.getvalue() == exp
    out = io.StringIO()
    cmds.search_cmd("apple fruit", files, 1, 3, None, False, True, None, model, out=out)
    js = json.loads(out.getvalue())["results"]
    assert [r["match_line_number"] for r in js] == [int(x) for x in h.match_line]
    assert [r["distance"] for r in js] == [float(x) for x in h.distance]
    # -i lowercases query and lines for embedding, keeps the original text (mod.rs:63-67,456-462)
    out = io.StringIO()
    cmds.search_cmd("HELLO WORLD", files, 0, 2, None, True, False, None, model, out=out)
    assert out.getvalue().splitlines()[0].startswith(f"{files[0]}:0::1 (0") and "   3: HELLO world" in out.getvalue()
    # threshold returns everything under it (mod.rs:115-116); stdin document name (cmds/search.rs:156-160)
    out = io.StringIO()
    cmds.search_cmd("apple fruit", [], 0, 1, 1.5, False, True, None, model, stdin_lines=flat, stdin_is_tty=False, out=out)
    js = json.loads(out.getvalue())["results"]
    assert all(r["filename"] == "<stdin>" and r["distance"] < 1.5 for r in js) and len(js) > 1


@pytest.mark.gpu
def test_pipelined_ingestion_matches_oracle_and_preserves_order(model_dir, ctx):
    """encode_with_args tokenises batch i+1 on a producer thread while K3 pools batch i:
    row order and bits must be those of the sequential oracle path."""
    d, E, vocab = model_dir
    model = StaticModel.from_pretrained(d, ctx=ctx)
    rng = np.random.default_rng(3)
    lines = [" ".join(rng.choice(WORDS + ["zzz"], rng.integers(0, 12))) for _ in range(5000)]
    got = model.encode_with_args(lines, 2048, 700)          # 8 batches through the queue
    off, ids = model.tokenize(lines, 2048)
    assert np.array_equal(got.view(np.uint32), oracle.embed_csr(E, off, ids).view(np.uint32))
    from semtools_b200 import capi
    c = capi.Corpus(ctx, 16)
    model.encode_with_args(lines, 2048, 512, append_to=c)
    assert len(c) == 5000 and np.array_equal(c.read().view(np.uint32), got.view(np.uint32))
