"""C++ host layer (semtools_b200/host): same surface as the reference's search module for
this path.  CPU: it builds, its pure host functions self-test, error output matches
cmds/search.rs:178-192.  GPU: the CLI's text/JSON output is byte-identical to the Python
mirror (which tests/test_cmds.py ties to the oracle)."""
import io
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "semtools_b200", "lib", "semtools_b200_search")
WORDS = ["alpha", "beta", "gamma", "delta", "hello", "world", "goodbye", "test", "line", "apple", "banana", "orange",
         "grape", "fruit"]


@pytest.fixture(scope="module")
def binary():
    subprocess.run(["bash", os.path.join(ROOT, "scripts", "build_host.sh")], check=True, cwd=ROOT)
    return BIN


def test_cpp_host_builds_and_selftests(binary):
    r = subprocess.run([binary, "--selftest"], capture_output=True, text=True)
    assert r.returncode == 0 and "selftest ok" in r.stdout, r.stdout + r.stderr


def test_cpp_no_input_error_matches_reference_text(binary, tmp_path):
    (tmp_path / "v.txt").write_text("a\n"); (tmp_path / "t.f32").write_bytes(np.zeros(256, np.float32).tobytes())
    base = [binary, "--vocab", str(tmp_path / "v.txt"), "--table", str(tmp_path / "t.f32"), "q"]
    r = subprocess.run(base, capture_output=True, text=True, stdin=subprocess.DEVNULL)
    assert r.returncode == 1
    assert r.stderr == "Error: No input provided. Either specify files as arguments or pipe input to stdin.\n"
    r = subprocess.run(base + ["-j"], capture_output=True, text=True, stdin=subprocess.DEVNULL)
    assert json.loads(r.stderr) == {"error": "No input provided. Either specify files as arguments or pipe input to stdin.",
                                    "error_type": "NoInput"}


@pytest.mark.gpu
def test_cpp_cli_output_equals_python_mirror(binary, tmp_path, ctx, monkeypatch):
    from safetensors.numpy import save_file
    from tokenizers import Tokenizer
    from tokenizers.models import WordLevel
    from tokenizers.pre_tokenizers import Whitespace
    from semtools_b200 import cmds
    from semtools_b200.model import StaticModel
    monkeypatch.delenv("SEMTOOLS_WORKSPACE", raising=False)
    vocab = {"[UNK]": 0, **{w: i + 1 for i, w in enumerate(WORDS)}}
    d = tmp_path / "model"; d.mkdir()
    tok = Tokenizer(WordLevel(vocab, unk_token="[UNK]")); tok.pre_tokenizer = Whitespace()
    tok.save(str(d / "tokenizer.json"))
    rng = np.random.default_rng(1)
    E = (rng.standard_normal((len(vocab), 256)) * 0.1).astype(np.float32)
    save_file({"embeddings": E}, str(d / "model.safetensors"))
    (d / "config.json").write_text(json.dumps({"normalize": True}))
    # the C++ WordLevel tokenizer drops unknown words itself: give it the vocabulary without [UNK]
    # but keep ids aligned with the table (line number = id)
    (tmp_path / "vocab.txt").write_text("\n".join(["\x00unused-unk"] + WORDS) + "\n")
    (tmp_path / "table.f32").write_bytes(E.tobytes())
    f1 = tmp_path / "a.txt"; f1.write_text("hello world\ngoodbye world\ntest line\n\napple banana zzz\n")
    f2 = tmp_path / "b.txt"; f2.write_text("orange grape\r\nfruit fruit\r\nHELLO world")
    files = [str(f1), str(f2)]
    model = StaticModel.from_pretrained(str(d), ctx=ctx)
    base = [binary, "--vocab", str(tmp_path / "vocab.txt"), "--table", str(tmp_path / "table.f32")]
    for extra, kw in [([], dict(n_lines=3, top_k=3, max_distance=None, ignore_case=False, json=False)),
                      (["-n", "1", "--top-k", "5", "-j"], dict(n_lines=1, top_k=5, max_distance=None, ignore_case=False, json=True)),
                      (["--threshold", "1.2", "--context", "0"], dict(n_lines=0, top_k=3, max_distance=1.2, ignore_case=False, json=False)),
                      (["-i", "-n", "0", "--top-k", "2"], dict(n_lines=0, top_k=2, max_distance=None, ignore_case=True, json=False))]:
        q = "HELLO fruit" if "-i" in extra else "apple fruit"
        out = io.StringIO()
        cmds.search_cmd(q, files, kw["n_lines"], kw["top_k"], kw["max_distance"], kw["ignore_case"], kw["json"], None, model, out=out)
        r = subprocess.run(base + [q] + files + extra, capture_output=True, text=True, stdin=subprocess.DEVNULL)
        assert r.returncode == 0, r.stderr
        assert r.stdout == out.getvalue(), (extra, r.stdout, out.getvalue())
    # stdin document
    out = io.StringIO()
    lines = ["apple banana", "hello world", "fruit"]
    cmds.search_cmd("fruit", [], 1, 2, None, False, True, None, model, stdin_lines=lines, stdin_is_tty=False, out=out)
    r = subprocess.run(base + ["fruit", "-n", "1", "--top-k", "2", "-j"], capture_output=True, text=True, input="\n".join(lines) + "\n")
    assert r.stdout == out.getvalue()


# ---------------------------------------------------------------- C++ workspace / store mirror
WSBIN = os.path.join(ROOT, "semtools_b200", "lib", "semtools_b200_workspace")


def _fake_embedding(path: str, line: int, scale: float) -> np.ndarray:
    """Same LCG as fake_embedding() in semtools_workspace_main.cpp."""
    from semtools_b200 import capi
    h = capi.line_id(path, line)
    out = np.empty(256, np.float32)
    for i in range(256):
        h = (h * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        out[i] = np.float32(scale) * np.float32((h >> 40) - (1 << 23)) / np.float32(1 << 23)
    return out


def _python_store_script(d):
    """The mutations store_selftest() applies, through the Python Store."""
    from semtools_b200.workspace import DocMeta, LineEmbedding, Store
    s = Store.open(str(d))
    b = 'dir/b "q".txt'
    s.upsert_document_metadata([DocMeta("a.txt", 10, 1700000000), DocMeta(b, 20, 1700000001),
                                DocMeta("cé.txt", 30, 1700000002), DocMeta("old.txt", 5, 1600000000, 1)])
    ls = [LineEmbedding("a.txt", i, _fake_embedding("a.txt", i, 1.0)) for i in range(3)]
    ls += [LineEmbedding(b, i, _fake_embedding(b, i, 1.0)) for i in range(2)]
    ls += [LineEmbedding("cé.txt", i, _fake_embedding("cé.txt", i, 1.0)) for i in range(4)]
    ls += [LineEmbedding("old.txt", 0, _fake_embedding("old.txt", 0, 1.0))]
    s.upsert_line_embeddings(ls)
    s.upsert_line_embeddings([LineEmbedding("a.txt", 1, _fake_embedding("a.txt", 1, 0.5))])
    s.upsert_document_metadata([DocMeta("a.txt", 11, 1700000005)])
    s.delete_documents([b, "old.txt"])
    s.upsert_line_embeddings([LineEmbedding("a.txt", 3, _fake_embedding("a.txt", 3, 1.0))])
    return s


def test_cpp_store_and_python_store_write_the_same_files(binary, tmp_path):
    """Store mutations (upsert / replace-by-id / delete / version-gated metadata delete,
    store.rs:235-434) through the C++ mirror and the Python mirror leave identical stores, and
    each side reads what the other wrote."""
    from semtools_b200.workspace import Store
    dc, dp = tmp_path / "cpp", tmp_path / "py"
    r = subprocess.run([WSBIN, "store-selftest", str(dc)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    _python_store_script(dp)
    r2 = subprocess.run([WSBIN, "store-dump", str(dp)], capture_output=True, text=True)     # C++ reads Python's files
    assert r2.returncode == 0, r2.stderr
    assert r.stdout == r2.stdout
    assert "documents 3" in r.stdout and "line_embeddings 8" in r.stdout
    mc, mp = (json.loads((d / "flat.b200" / "store.json").read_text()) for d in (dc, dp))
    assert mc == mp                                                  # same commit record: rows, gen, files, paths, docs
    assert mc["rows"] == 8 and mc["files"]["rows"] != "rows.i32"     # the deletions committed a new generation of row files
    for name in mc["files"].values():
        assert (dc / "flat.b200" / name).read_bytes() == (dp / "flat.b200" / name).read_bytes()
    assert sorted(p.name for p in (dc / "flat.b200").iterdir()) == sorted(p.name for p in (dp / "flat.b200").iterdir())
    s = Store.open(str(dc))                                                               # Python reads C++'s files
    assert s.get_all_document_paths() == ["a.txt", "cé.txt", "old.txt"]
    assert s.count_line_embeddings() == 8
    assert [(s._paths[pi], int(ln)) for pi, ln in s._rows] == \
        [("a.txt", 0), ("a.txt", 1), ("a.txt", 2)] + [("cé.txt", i) for i in range(4)] + [("a.txt", 3)]
    assert np.array_equal(s._emb[1], _fake_embedding("a.txt", 1, 0.5))
    assert s._ranges_for(["a.txt"]).tolist() == [[0, 3], [7, 8]]


def test_cpp_workspace_commands_match_python_mirror(binary, tmp_path):
    """workspace use / status / prune (cmds/workspace.rs:11-176): C++ output == Python output."""
    from semtools_b200 import cmds
    from semtools_b200.workspace import DocMeta, Store
    env = dict(os.environ, HOME=str(tmp_path))
    env.pop("SEMTOOLS_WORKSPACE", None)
    r = subprocess.run([WSBIN, "status"], capture_output=True, text=True, env=env)
    assert r.returncode == 1 and r.stderr == "Error: No active workspace. Run: workspace use <name>\n"

    old_home, old_ws = os.environ.get("HOME"), os.environ.pop("SEMTOOLS_WORKSPACE", None)
    os.environ["HOME"] = str(tmp_path)
    try:
        for js in (False, True):
            flags = ["--json"] if js else []
            r = subprocess.run([WSBIN] + flags + ["use", "proj"], capture_output=True, text=True, env=env)
            buf = io.StringIO(); cmds.workspace_use_cmd("proj", js, out=buf)
            assert r.returncode == 0 and r.stdout == buf.getvalue()
        cfg = json.loads((tmp_path / ".semtools" / "workspaces" / "proj" / "config.json").read_text())
        assert cfg == {"name": "proj", "root_dir": str(tmp_path / ".semtools" / "workspaces" / "proj"),
                       "in_batch_size": 5000, "oversample_factor": 3}
        live = tmp_path / "live.txt"; live.write_text("x\n")
        s = Store.open(cfg["root_dir"])
        s.upsert_document_metadata([DocMeta(str(live), 2, 1), DocMeta(str(tmp_path / "gone.txt"), 3, 1)])
        for js in (False, True):
            flags = ["--json"] if js else []
            r = subprocess.run([WSBIN] + flags + ["--workspace", "proj", "status"], capture_output=True, text=True, env=env)
            buf = io.StringIO(); cmds.workspace_status_cmd(js, "proj", out=buf)
            assert r.returncode == 0 and r.stdout == buf.getvalue()
        env2 = dict(env, SEMTOOLS_WORKSPACE="proj")
        r = subprocess.run([WSBIN, "prune"], capture_output=True, text=True, env=env2)
        assert r.returncode == 0
        assert r.stdout == f"Found 1 stale documents:\n  - {tmp_path / 'gone.txt'}\nRemoved 1 stale documents from workspace.\n"
        assert Store.open(cfg["root_dir"]).get_all_document_paths() == [str(live)]
        r = subprocess.run([WSBIN, "--json", "prune"], capture_output=True, text=True, env=env2)
        buf = io.StringIO(); cmds.workspace_prune_cmd(True, "proj", out=buf)
        assert r.stdout == buf.getvalue() and json.loads(r.stdout) == {"files_removed": 0, "files_remaining": 1}
    finally:
        os.environ["HOME"] = old_home
        if old_ws is not None:
            os.environ["SEMTOOLS_WORKSPACE"] = old_ws


def test_cpp_rust_lines_equals_python_mirror_including_bare_cr(binary):
    """the Python host used to split on a bare CR, the C++ host did not; both now follow
    str::lines() (search/mod.rs:55) and share one store, so they must agree byte for byte."""
    import json as _json
    from semtools_b200.workspace import _rust_lines
    for text in ["one\rstill one\ntwo\r\nthree", "a\r", "a\r\n", "a\r\r\nb", "", "\n", "x\n\n", "t\tab \"q\" \\ z\n", "\r\n\r\n", "é\rü\n"]:
        r = subprocess.run([binary, "--lines"], input=text.encode("utf-8"), capture_output=True)
        assert r.returncode == 0
        got = [_json.loads(l) for l in r.stdout.decode("utf-8").splitlines()]
        assert got == _rust_lines(text), (text, got)


def test_cpp_to_lowercase_is_full_unicode(binary):
    """str::to_lowercase (mod.rs:63-67): full mapping, multi-char expansions, Final_Sigma,
    case-ignorable skipping -- byte-identical to Python's str.lower() (same UCD algorithm)."""
    samples = ["Hello WORLD", "ÀÉÎÕÜ Straße ẞ", "İstanbul I ı", "ΑΣ ΟΔΥΣΣΕΥΣ ΣΑΣ Σ", "ΑΣ.Β ΑΣ' Α­Σ ́Σ",
               "ǅ ǈ Ǌ ᾈ ᾼ", "ԱԲԳ ᲐᲑᲒ 𐐀𐐁 𞤀𞤁", "Ⅷ Ⓐ Ｆｕｌｌ", "ǰ ŉ ΐ", "日本語 emoji 😀 ÇA", "AͅΣ", "ΣΣ", "aΣʰb", ""]
    import random
    rng = random.Random(7)
    pool = [c for c in map(chr, list(range(0x20, 0x250)) + list(range(0x370, 0x530)) + list(range(0x1E00, 0x2000))
                           + [0x3A3, 0x3C2, 0x3C3, 0x2019, 0x27, 0x2E, 0xAD, 0x301, 0x345, 0x2B0, 0x10400, 0x1E900])]
    samples += ["".join(rng.choice(pool) for _ in range(40)) for _ in range(200)]
    text = "\n".join(samples)
    r = subprocess.run([binary, "--lower"], input=text.encode("utf-8"), capture_output=True)
    assert r.returncode == 0
    got = r.stdout.decode("utf-8").split("\n")
    want = text.lower().split("\n")
    # Final_Sigma looks across the '\n' separators identically on both sides (same whole string)
    assert got == want
    # malformed UTF-8 passes through byte for byte, ASCII around it is still lowered
    r = subprocess.run([binary, "--lower"], input=b"AB\xff\xc3(\xe2\x82CD", capture_output=True)
    assert r.stdout == b"ab\xff\xc3(\xe2\x82cd"


def test_cpp_workspace_renderers_match_python(binary, tmp_path):
    """print_workspace_search_results / workspace JSON (cmds/search.rs:66-110, 208-237): C++ == Python,
    including the f32 Display, the unclamped header end, unreadable files and CRLF lines."""
    import struct
    from semtools_b200 import cmds
    from semtools_b200.search import RankedLine
    f1 = tmp_path / "a b.txt"; f1.write_text("l0\nl1\r\nl2 \"q\"\nl3\n\nl5")
    f2 = tmp_path / "é.txt"; f2.write_text("only\n")
    gone = tmp_path / "gone.txt"
    ranked = [RankedLine(str(f1), 2, float(np.float32(0.1))), RankedLine(str(f2), 0, 0.0), RankedLine(str(f1), 5, float(np.float32(1.5e-7))),
              RankedLine(str(gone), 3, float(np.float32(0.33333334))), RankedLine(str(f1), 0, 1.0)]
    feed = "".join(f"{r.path}\t{r.line_number}\t{struct.unpack('<I', struct.pack('<f', r.distance))[0]}\n" for r in ranked)
    for n_lines in (0, 1, 3, 50):
        r = subprocess.run([WSBIN, "format-ranked", str(n_lines)], input=feed, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert r.stdout == cmds.format_workspace_search_results(ranked, n_lines, False)
        r = subprocess.run([WSBIN, "--json", "format-ranked", str(n_lines)], input=feed, capture_output=True, text=True)
        assert r.stdout == cmds.to_string_pretty({"results": cmds.workspace_results_to_json(ranked, n_lines)}) + "\n"
    r = subprocess.run([WSBIN, "--json", "format-ranked", "2"], input="", capture_output=True, text=True)
    assert r.stdout == cmds.to_string_pretty({"results": []}) + "\n"


@pytest.mark.gpu
def test_cpp_cli_workspace_mode_equals_python_mirror(binary, tmp_path, ctx, monkeypatch):
    """search with an active workspace (search/mod.rs:146-216 + cmds/search.rs:194-241): the C++ CLI
    and the Python mirror build their own workspaces over the same files and print the same
    bytes; a second C++ run re-uses the store (nothing re-embedded), an edited file is re-embedded."""
    from safetensors.numpy import save_file
    from tokenizers import Tokenizer
    from tokenizers.models import WordLevel
    from tokenizers.pre_tokenizers import Whitespace
    from semtools_b200 import cmds
    from semtools_b200.model import StaticModel
    vocab = {"[UNK]": 0, **{w: i + 1 for i, w in enumerate(WORDS)}}
    d = tmp_path / "model"; d.mkdir()
    tok = Tokenizer(WordLevel(vocab, unk_token="[UNK]")); tok.pre_tokenizer = Whitespace()
    tok.save(str(d / "tokenizer.json"))
    rng = np.random.default_rng(1)
    E = (rng.standard_normal((len(vocab), 256)) * 0.1).astype(np.float32)
    save_file({"embeddings": E}, str(d / "model.safetensors"))
    (d / "config.json").write_text(json.dumps({"normalize": True}))
    (tmp_path / "vocab.txt").write_text("\n".join(["\x00unused-unk"] + WORDS) + "\n")
    (tmp_path / "table.f32").write_bytes(E.tobytes())
    f1 = tmp_path / "a.txt"; f1.write_text("hello world\ngoodbye world\ntest line\n\napple banana zzz\n")
    f2 = tmp_path / "b.txt"; f2.write_text("orange grape\r\nfruit fruit\r\nHELLO world")
    files = [str(f1), str(f2)]
    model = StaticModel.from_pretrained(str(d), ctx=ctx)
    base = [binary, "--vocab", str(tmp_path / "vocab.txt"), "--table", str(tmp_path / "table.f32")]
    home_py, home_cpp = tmp_path / "home_py", tmp_path / "home_cpp"
    home_py.mkdir(); home_cpp.mkdir()
    env = dict(os.environ, HOME=str(home_cpp), SEMTOOLS_WORKSPACE="ws")
    monkeypatch.setenv("HOME", str(home_py))
    monkeypatch.setenv("SEMTOOLS_WORKSPACE", "ws")
    for extra, kw in [([], dict(n_lines=3, top_k=3, max_distance=None, json=False)),
                      (["-n", "1", "--top-k", "5", "-j"], dict(n_lines=1, top_k=5, max_distance=None, json=True)),
                      (["--threshold", "0.9", "--context", "0"], dict(n_lines=0, top_k=3, max_distance=0.9, json=False))]:
        out, err = io.StringIO(), io.StringIO()
        cmds.search_cmd("apple fruit", files, kw["n_lines"], kw["top_k"], kw["max_distance"], False, kw["json"], None, model,
                        out=out, err=err)
        r = subprocess.run(base + ["apple fruit"] + files + extra, capture_output=True, text=True, stdin=subprocess.DEVNULL, env=env)
        assert r.returncode == 0, r.stderr
        assert r.stdout == out.getvalue(), (extra, r.stdout, out.getvalue())
        assert r.stderr == err.getvalue()                    # "Updating workspace with ..." only on the first pass
    f1.write_text("hello world\napple fruit\n")              # changed: size differs -> re-embedded on both sides
    out, err = io.StringIO(), io.StringIO()
    cmds.search_cmd("apple fruit", files, 0, 1, None, False, False, None, model, out=out, err=err)
    r = subprocess.run(base + ["apple fruit"] + files + ["-n", "0", "--top-k", "1"], capture_output=True, text=True,
                       stdin=subprocess.DEVNULL, env=env)
    assert r.stdout == out.getvalue() and r.stdout.startswith(f"{f1}:1::2 (")
    assert "Updating workspace with 2 lines" in r.stderr and r.stderr == err.getvalue()


def test_cpp_tokenize_bench_hook_runs(binary):
    """Host tokenisation throughput hook; `--selftest` already checks that every
    thread count produces the serial CSR."""
    r = subprocess.run([binary, "--tokenize-bench", "20000", "2"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    rows = [json.loads(l) for l in r.stdout.strip().splitlines()]
    assert [x["threads"] for x in rows] == [1, 2]
    assert rows[0]["tokens"] == rows[1]["tokens"] > 100_000 and rows[0]["lines"] == 20000


def _synthetic_unigram(tmp_path, seed, normalizer, prepend_scheme="always", with_unk=True):
    """A SentencePiece-style tokenizer.json built with HF `tokenizers`: single characters plus random
    multi-character pieces with random log-probabilities, Metaspace pre-tokenizer."""
    import random
    from tokenizers import Tokenizer, models, normalizers, pre_tokenizers
    rnd = random.Random(seed)
    alphabet = list("abcdefghijklmnopqrstuvwxyz0123456789.,-") + ["é", "ü", "ß", "λ", "σ", "ς", "中", "文", "▁"]
    vocab = [("<unk>", 0.0)] if with_unk else []
    seen = {"<unk>"}
    for ch in alphabet[:-6] + ["▁"]:                      # some characters stay OUT of the vocabulary (-> unk)
        vocab.append((ch, -rnd.uniform(4.0, 9.0))); seen.add(ch)
    while len(vocab) < 1500:
        n = rnd.choice([2, 2, 3, 3, 4, 5, 6])
        tok = ("▁" if rnd.random() < 0.35 else "") + "".join(rnd.choice(alphabet[:-3]) for _ in range(n))
        if tok not in seen:
            seen.add(tok); vocab.append((tok, -rnd.uniform(2.0, 14.0)))
    vocab.append(("▁the", -2.5)); vocab.append(("ing", -3.0)); vocab.append(("ab", -3.25)); vocab.append(("abab", -6.5))   # exact score ties on purpose
    tk = Tokenizer(models.Unigram(vocab, unk_id=0 if with_unk else None, byte_fallback=False))
    if normalizer is not None:
        tk.normalizer = normalizer
    tk.pre_tokenizer = pre_tokenizers.Metaspace(replacement="▁", prepend_scheme=prepend_scheme, split=True)
    path = tmp_path / f"tokenizer_{seed}.json"
    tk.save(str(path))
    return tk, path, alphabet


@pytest.mark.parametrize("case", ["lower+multispace", "plain_first", "strip_prepend", "nfkc_ascii"])
def test_cpp_unigram_tokenizer_matches_hf_tokenizers_id_for_id(binary, tmp_path, case):
    """the C++ host tokenises with the model's tokenizer.json (Unigram + Metaspace
    subset) instead of a WordLevel stand-in.  Token ids -- with and without the unk-drop of
    encode_with_args (mod.rs:69) -- must equal HF `tokenizers` on random text, including unknown
    characters (fused unk), score ties, multiple spaces and non-ASCII lowercase."""
    import random
    from tokenizers import Regex, normalizers
    norm = {"lower+multispace": normalizers.Sequence([normalizers.Lowercase(), normalizers.Replace(Regex(" {2,}"), " ")]),
            "plain_first": None,
            "strip_prepend": normalizers.Sequence([normalizers.Strip(), normalizers.Replace("--", "-")]),
            "nfkc_ascii": normalizers.Sequence([normalizers.NFKC(), normalizers.Lowercase()])}[case]
    tk, path, alphabet = _synthetic_unigram(tmp_path, seed=len(case), normalizer=norm,
                                            prepend_scheme="first" if case == "plain_first" else "always")
    rnd = random.Random(99)
    ascii_only = case == "nfkc_ascii"
    pool = [a for a in alphabet if (ord(a[0]) < 128 or not ascii_only)] + ["X", "Q", "THE", "Σ", "ΑΣ", "ab", "abab", "ing"]
    if ascii_only:
        pool = [p for p in pool if all(ord(c) < 128 for c in p)]
    lines = ["", " ", "the thing", "  leading and   multiple   spaces  ", "abab abababab", "ΑΣ ΣΑΣ σας" if not ascii_only else "AS SAS sas"]
    for _ in range(300):
        words = ["".join(rnd.choice(pool) for _ in range(rnd.randint(1, 9))) for _ in range(rnd.randint(0, 12))]
        lines.append((" " * rnd.randint(0, 3)).join(words) if rnd.random() < 0.2 else " ".join(words))
    lines = [l for l in lines if "\n" not in l and "\r" not in l]
    r = subprocess.run([binary, "--encode", str(path)], input=("\n".join(lines) + "\n").encode("utf-8"), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    out = r.stdout.decode().splitlines()
    lens = sorted(len(t.encode("utf-8")) for t in tk.get_vocab(False).keys())
    assert out[0] == f"median_token_length {lens[len(lens) // 2]} vocab {tk.get_vocab_size()}"
    assert len(out) == len(lines) + 1
    unk = 0
    for line, got in zip(lines, out[1:]):
        want = tk.encode(line, add_special_tokens=False).ids
        raw, dropped = got.split("|")
        assert [int(x) for x in raw.split()] == want, (case, line)
        assert [int(x) for x in dropped.split()] == [i for i in want if i != unk], (case, line)
    if ascii_only:      # non-ASCII text under NFKC is refused, not silently mis-normalised
        r = subprocess.run([binary, "--encode", str(path)], input="café\n".encode("utf-8"), capture_output=True)
        assert "ERROR" in r.stdout.decode() and "Python host" in r.stdout.decode()


def test_cpp_tokenizer_rejects_what_it_does_not_implement(binary, tmp_path):
    from tokenizers import Tokenizer, models, pre_tokenizers
    tk = Tokenizer(models.WordPiece({"[UNK]": 0, "a": 1}, unk_token="[UNK]"))
    tk.pre_tokenizer = pre_tokenizers.Whitespace()
    p = tmp_path / "wp.json"; tk.save(str(p))
    r = subprocess.run([binary, "--encode", str(p)], input=b"a\n", capture_output=True)
    assert r.returncode != 0 and b"not supported by the C++ host" in r.stderr
