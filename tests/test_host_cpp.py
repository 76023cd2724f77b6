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
    saw_unk = 0
    for line, got in zip(lines, out[1:]):
        want = tk.encode(line, add_special_tokens=False).ids
        raw, dropped = got.split("|")
        assert [int(x) for x in raw.split()] == want, (case, line)
        # a Unigram model section names no `unk_token` (it has `unk_id`): encode_with_args drops nothing,
        # exactly as the Python host (model.py: unk_token_id is None) -- the two hosts share one store
        assert [int(x) for x in dropped.split()] == want, (case, line)
        saw_unk += 0 in want
    assert saw_unk > 20 or ascii_only
    # a model section that NAMES its unk token: that id is removed (ids.retain(|id| id != unk_token_id))
    j = json.loads(path.read_text(encoding="utf-8"))
    j["model"]["unk_token"] = "<unk>"
    named = tmp_path / "named_unk.json"
    named.write_text(json.dumps(j), encoding="utf-8")
    r = subprocess.run([binary, "--encode", str(named)], input=("\n".join(lines) + "\n").encode("utf-8"), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    for line, got in zip(lines, r.stdout.decode().splitlines()[1:]):
        want = tk.encode(line, add_special_tokens=False).ids
        assert [int(x) for x in got.split("|")[1].split()] == [i for i in want if i != 0], (case, line)
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


def _u(s):
    return s.encode("ascii").decode("unicode_escape")


def test_cpp_grapheme_clusters_match_uax29(binary):
    """HF tokenizers' Precompiled normalizer walks `graphemes(true)`; the C++ host cuts extended grapheme
    clusters with generated Unicode property tables (host/grapheme_break.inc).  Checked against the `regex`
    module's \\X on random strings over every break class (CR/LF/Control, Extend, ZWJ, regional indicators,
    Prepend, SpacingMark, Hangul jamo and syllables, emoji ZWJ sequences, Indic conjuncts)."""
    regex = pytest.importorskip("regex")
    import random
    rnd = random.Random(5)
    cands = [0x61, 0x41, 0x20, 0x0D, 0x09, 0x301, 0x302, 0x308, 0x200D, 0x200C, 0x1F1E6, 0x1F1FA, 0x1F1F8, 0x600, 0x110BD, 0x903, 0x93E, 0x915, 0x937,
             0x94D, 0x1100, 0x1161, 0x11A8, 0xAC00, 0xAC01, 0x1F468, 0x1F469, 0x1F466, 0x2764, 0xFE0F, 0x1F3FB, 0x1F600, 0xE0020, 0xE007F, 0x0E01, 0x0E33,
             0x0E48, 0x0BA8, 0x0BBF, 0x0D15, 0x0D4D, 0x0D30, 0x0A95, 0x0ACD, 0x0AB7, 0xFF21, 0x3042, 0x4E2D, 0x00E9, 0x1E9E, 0xFB01, 0x2122, 0x00AD, 0x061C,
             0x180E, 0x1F9D1, 0x1F91D, 0x2640, 0x1F3F4, 0xE0067, 0x0C15, 0x0C4D, 0x0C37, 0x1B05, 0x1B44, 0x11F02, 0x0D4E, 0x1193F, 0x11941, 0xA9, 0x3030]
    lines = ["".join(chr(rnd.choice(cands)) for _ in range(rnd.randint(1, 12))) for _ in range(3000)]
    lines += [_u(x) for x in ("e\\u0301\\u0302", "\\U0001f468\\u200d\\U0001f469\\u200d\\U0001f467\\u200d\\U0001f466", "\\U0001f1fa\\U0001f1f8\\U0001f1e6",
                              "\\u0915\\u094d\\u0937", "\\u0915\\u094d\\u200d\\u0937", "\\u1100\\u1161\\u11a8", "\\r\\rx", "a\\rb")]
    lines = [l for l in lines if "\n" not in l and not l.endswith("\r")]       # rust_lines strips a final \r
    r = subprocess.run([binary, "--graphemes"], input=("\n".join(lines) + "\n").encode("utf-8"), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    got = r.stdout.decode().split("\n")[:-1]
    assert len(got) == len(lines)
    for line, g in zip(lines, got):
        assert [int(x) for x in g.split()] == [len(x.encode("utf-8")) for x in regex.findall(r"\X", line)], [hex(ord(c)) for c in line]


@pytest.fixture(scope="module")
def nmt_nfkc_tokenizer(tmp_path_factory):
    """A tokenizer.json of the XLM-R family shape (what minishlab/potion-multilingual-128M carries): Unigram +
    Metaspace + Sequence[Precompiled(nmt_nfkc charsmap), Replace(" {2,}")].  The charsmap is the REAL one:
    sentencepiece writes it into any model trained with normalization_rule_name=nmt_nfkc."""
    spm = pytest.importorskip("sentencepiece")
    pb = pytest.importorskip("sentencepiece.sentencepiece_model_pb2")
    import random
    from tokenizers import Regex, Tokenizer
    from tokenizers.models import Unigram
    from tokenizers.normalizers import Precompiled, Replace, Sequence
    from tokenizers.pre_tokenizers import Metaspace
    d = tmp_path_factory.mktemp("spm")
    rnd = random.Random(3)
    words = _u("the quick brown fox na\\u00efve caf\\u00e9 stra\\u00dfe \\u00fcber \\u043f\\u0440\\u0438\\u0432\\u0435\\u0442 \\u043c\\u0438\\u0440 \\u0451\\u0436\\u0438\\u043a "
               "\\u03ba\\u03b1\\u03bb\\u03b7\\u03bc\\u03ad\\u03c1\\u03b1 \\u1f48\\u03b4\\u03c5\\u03c3\\u03c3\\u03b5\\u03cd\\u03c2 \\u3053\\u3093\\u306b\\u3061\\u306f \\u4e16\\u754c "
               "\\uff76\\uff9e\\uff77\\uff9e \\uff83\\uff7d\\uff84 \\u30d1\\u30fc\\u30c6\\u30a3\\u30fc \\u4f60\\u597d \\u6e2c\\u8a66 \\uc548\\ub155\\ud558\\uc138\\uc694 \\ud55c\\uad6d\\uc5b4 "
               "\\u0928\\u092e\\u0938\\u094d\\u0924\\u0947 \\u0915\\u094d\\u0937\\u0924\\u094d\\u0930\\u093f\\u092f \\u0645\\u0631\\u062d\\u0628\\u0627 \\u0634\\u0643\\u0631\\u0627 "
               "\\u0e2a\\u0e27\\u0e31\\u0e2a\\u0e14\\u0e35 \\u0e19\\u0e49\\u0e33 \\uff21\\uff22\\uff23 \\uff11\\uff12\\uff13 \\ufb01ne \\ufb02ow \\u2122 \\u00b2 \\u2462 \\u2167 \\u338f").split()
    lines = [" ".join(rnd.choice(words) for _ in range(rnd.randint(3, 12))) for _ in range(4000)]
    (d / "corpus.txt").write_text("\n".join(lines) + "\n", encoding="utf-8")
    spm.SentencePieceTrainer.train(input=str(d / "corpus.txt"), model_prefix=str(d / "m"), vocab_size=400, model_type="unigram",
                                   normalization_rule_name="nmt_nfkc", character_coverage=1.0, hard_vocab_limit=False, minloglevel=2)
    mp = pb.ModelProto()
    mp.ParseFromString((d / "m.model").read_bytes())
    charsmap = mp.normalizer_spec.precompiled_charsmap
    assert len(charsmap) > 100_000                                            # the nmt_nfkc table, not an empty one
    tk = Tokenizer(Unigram([(p.piece, p.score) for p in mp.pieces], unk_id=next(i for i, p in enumerate(mp.pieces) if p.type == 2), byte_fallback=False))
    tk.normalizer = Sequence([Precompiled(charsmap), Replace(Regex(" {2,}"), " ")])
    tk.pre_tokenizer = Metaspace(replacement=_u("\\u2581"), prepend_scheme="always")
    tk.save(str(d / "tokenizer.json"))
    return tk, str(d / "tokenizer.json"), lines


def test_cpp_precompiled_normalizer_and_ids_match_hf_tokenizers(binary, nmt_nfkc_tokenizer):
    """The C++ host applies the SentencePiece charsmap (double-array trie, grapheme by grapheme, the crate's
    quirks included) and then tokenises: normalizer output AND token ids must equal HF tokenizers' on
    multilingual text -- composed / decomposed accents, full- and half-width forms, ligatures, compatibility
    symbols, zero-width and control characters, NBSP / ideographic space, emoji sequences, Indic conjuncts,
    Hangul jamo, NUL."""
    import random
    tk, path, corpus_lines = nmt_nfkc_tokenizer
    rnd = random.Random(11)
    extra = ["e\\u0301\\u0302 cafe\\u0301", "\\ufb01\\ufb02 \\uff21\\uff22\\uff23\\uff11\\uff12\\uff13", "\\uff76\\uff9e \\uff8a\\uff9f \\uff73\\uff9e", "\\u3000ideographic\\u3000space",
             "a\\u00a0b\\u2009c\\u200bd\\u200ce", "tab\\there", "ctl\\x01\\x02x", "\\u00bd \\u00bc \\u2460 \\u3231 \\u337b \\u2103 \\u212b \\u2126",
             "\\U0001f1fa\\U0001f1f8\\U0001f1e6 \\U0001f468\\u200d\\U0001f469\\u200d\\U0001f467\\u200d\\U0001f466 \\u2764\\ufe0f \\U0001f44d\\U0001f3fd",
             "\\u0915\\u094d\\u0937 \\u0924\\u094d\\u0930 \\u091c\\u094d\\u091e", "\\uac01 \\u1100 \\u1161 \\u11a8 \\u1100\\u1161\\u11a8", "\\u01c4 \\u01c5 \\u01c6 \\u01c8", "\\u017f \\u1e9b \\ufb05",
             "e\\u0301\\u0301\\u0301\\u0301", "a\\u0300\\u0301\\u0302\\u0303b", "x\\rz", "  multiple   spaces  ", "\\u2460\\u2461 \\u00b2\\u00b3 \\u2122", "\\uff21\\u0301", "\\uff76\\uff9e\\uff76",
             "\\ufdfa \\ufdf2", "\\u1e9b\\u0323", "\\u0958 \\u0915\\u093c", "\\u314f \\u3131 \\u3132", "\\ufb2c \\ufb49", "\\U0001d400\\U0001d401\\U0001d402 \\U0001d7d8\\U0001d7d9", "\\u3300 \\u3301",
             "\\u0300\\u0301 \\u0343 \\u0344", "\\u00ad soft\\u00adhyphen", "\\ufeff bom", "\\u2028ls\\u2029", "\\u0085nel", "\\u212b \\u00c5", "\\u0e33 \\u0eb3", "\\u0f77 \\u0f79",
             "\\u2026  \\u2025 \\u2024", "\\ufe30 \\ufe50 \\ufe52", "\\u0000nul", "a\\u0000b", ""]
    tests = corpus_lines[:300] + [_u(x) for x in extra]
    cps = [0x65, 0x301, 0x302, 0xFB01, 0xFF21, 0xFF9E, 0xFF76, 0x3000, 0xA0, 0x200B, 0x200D, 0x1F468, 0x915, 0x94D, 0x937, 0x1100, 0x1161, 0x11A8, 0xAC00, 0x20, 0x20,
           0x61, 0x2122, 0xB2, 0x1E9B, 0x323, 0x9, 0x1, 0x7F, 0xAD, 0x2460, 0x3300, 0x1D400]
    tests += ["".join(chr(rnd.choice(cps)) for _ in range(rnd.randint(1, 14))) for _ in range(600)]
    tests = [t for t in tests if "\n" not in t and not t.endswith("\r")]
    blob = ("\n".join(tests) + "\n").encode("utf-8")
    r = subprocess.run([binary, "--normalize", path], input=blob, capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    got = r.stdout.decode("utf-8").split("\n")[:-1]
    assert len(got) == len(tests)
    changed = 0
    for t, g in zip(tests, got):
        want = tk.normalizer.normalize_str(t)
        assert g == want, ([hex(ord(c)) for c in t], [hex(ord(c)) for c in want], [hex(ord(c)) for c in g])
        changed += want != t
    assert changed > 300                                                       # the charsmap really was exercised
    r = subprocess.run([binary, "--encode", path], input=blob, capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    out = r.stdout.decode().split("\n")[1:-1]
    assert len(out) == len(tests)
    for t, g in zip(tests, out):
        assert [int(x) for x in g.split("|")[0].split()] == tk.encode(t, add_special_tokens=False).ids, [hex(ord(c)) for c in t]


def test_cpp_precompiled_with_an_empty_charsmap_is_the_identity(binary, tmp_path):
    from tokenizers import Tokenizer, models, pre_tokenizers
    tk = Tokenizer(models.Unigram([("<unk>", 0.0), (_u("\\u2581"), -2.0), ("a", -3.0), (_u("\\u00e9"), -3.5)], unk_id=0, byte_fallback=False))
    tk.pre_tokenizer = pre_tokenizers.Metaspace(replacement=_u("\\u2581"), prepend_scheme="always")
    p = tmp_path / "t.json"
    tk.save(str(p))
    j = json.loads(p.read_text(encoding="utf-8"))
    j["normalizer"] = {"type": "Precompiled", "precompiled_charsmap": None}     # what some exported tokenizer.json files carry
    p.write_text(json.dumps(j), encoding="utf-8")
    r = subprocess.run([binary, "--encode", str(p)], input=_u("a\\u00e9 a\n").encode("utf-8"), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    assert r.stdout.decode().splitlines()[1].split("|")[0].split() == ["1", "2", "3", "1", "2"]


@pytest.mark.parametrize("scheme", ["always", "first"])
def test_cpp_tokenizer_splits_added_tokens_like_hf(binary, tmp_path, scheme):
    """AddedVocabulary::extract_and_normalize: special / added tokens that occur literally in a line ("<s>" in an
    HTML document, "<mask>") become their ids before the model runs -- non-normalised ones matched on the raw
    text, normalised ones on the normalised text, leftmost-longest, lstrip / rstrip widening the match."""
    import random
    from tokenizers import AddedToken, Regex, Tokenizer, models, normalizers, pre_tokenizers
    rnd = random.Random(7)
    sp = _u("\\u2581")
    alphabet = list("abcdefghijklmnopqrstuvwxyz<>/") + [sp]
    vocab, seen = [("<unk>", 0.0)], {"<unk>"}
    for ch in alphabet:
        vocab.append((ch, -rnd.uniform(4, 9))); seen.add(ch)
    while len(vocab) < 600:
        t = (sp if rnd.random() < 0.35 else "") + "".join(rnd.choice(alphabet[:26]) for _ in range(rnd.choice([2, 3, 4, 5])))
        if t not in seen:
            seen.add(t); vocab.append((t, -rnd.uniform(2, 14)))
    tk = Tokenizer(models.Unigram(vocab, unk_id=0, byte_fallback=False))
    tk.normalizer = normalizers.Sequence([normalizers.Lowercase(), normalizers.Replace(Regex(" {2,}"), " ")])
    tk.pre_tokenizer = pre_tokenizers.Metaspace(replacement=sp, prepend_scheme=scheme, split=True)
    tk.add_special_tokens(["<s>", "</s>", "<pad>", AddedToken("<mask>", lstrip=True, special=True)])
    tk.add_tokens([AddedToken("Foo Bar", normalized=True), AddedToken("<x>", lstrip=True, rstrip=True, normalized=False), AddedToken("<s>>", normalized=False)])
    path = tmp_path / "added.json"
    tk.save(str(path))
    pool = ["<s>", "</s>", "<pad>", "<mask>", "foo bar", "FOO BAR", "Foo  Bar", "<x>", "<s>>", "<", ">", "abc", "the", " ", "  ", "x", "<unk>", "<S>", "</s", "s>"]
    lines = ["<s>hello</s>", "a <mask> b", "a  <mask>  b", "<x>", " <x> ", "a<x>b", "a  <x>   b", "foo bar", "xfoo bary", "<s>><s>", "<pad><pad>", "", "<s>", " <s>",
             "<s> ", "<mask><mask>", "a <mask>", "<mask> a"]
    lines += ["".join(rnd.choice(pool) for _ in range(rnd.randint(1, 8))) for _ in range(400)]
    r = subprocess.run([binary, "--encode", str(path)], input=("\n".join(lines) + "\n").encode("utf-8"), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    out = r.stdout.decode().split("\n")[1:-1]
    assert len(out) == len(lines)
    hit = 0
    for line, g in zip(lines, out):
        want = tk.encode(line, add_special_tokens=False).ids
        assert [int(x) for x in g.split("|")[0].split()] == want, (scheme, line)
        hit += any(i >= len(vocab) for i in want)
    assert hit > 200                                                           # added-token ids really occur
    j = json.loads(path.read_text(encoding="utf-8"))
    j["added_tokens"][0]["single_word"] = True                                  # not restated: refused, not mis-tokenised
    path.write_text(json.dumps(j), encoding="utf-8")
    r = subprocess.run([binary, "--encode", str(path)], input=b"a\n", capture_output=True)
    assert r.returncode != 0 and b"single_word" in r.stderr


@pytest.mark.parametrize("table_dtype", ["F32", "F16", "I8"])
def test_cpp_model_directory_loader_equals_the_python_loader(binary, tmp_path, table_dtype):
    """StaticModel::from_pretrained on a local directory (src/cmds/search.rs:123-128): the C++ host reads
    tokenizer.json + model.safetensors (F32 / F16 / I8 table upcast to f32, optional weights and mapping) +
    config.json itself.  Tensors, normalize flag and the store fingerprint must equal what the Python host
    (model.py) loads from the same directory -- the two hosts then share one workspace."""
    st = pytest.importorskip("safetensors.numpy")
    from semtools_b200 import capi
    from semtools_b200.model import StaticModel
    rng = np.random.default_rng(12)
    tk, tok_path, _ = _synthetic_unigram(tmp_path, seed=5, normalizer=None)
    V = tk.get_vocab_size()
    emb32 = (rng.standard_normal((V, 256)) * 0.1).astype(np.float32)
    emb = {"F32": emb32, "F16": emb32.astype(np.float16), "I8": np.clip(np.round(emb32 * 400), -127, 127).astype(np.int8)}[table_dtype]
    d = tmp_path / "model"
    d.mkdir()
    os.replace(tok_path, d / "tokenizer.json")
    tensors = {"embeddings": emb}
    if table_dtype != "F32":                                             # the newer model2vec layout: per-token weights + id mapping
        tensors["weights"] = rng.uniform(0.1, 2.0, V).astype(np.float64 if table_dtype == "I8" else np.float32)
        tensors["mapping"] = rng.permutation(V).astype(np.int64 if table_dtype == "I8" else np.int32)
    st.save_file(tensors, str(d / "model.safetensors"))
    (d / "config.json").write_text(json.dumps({"normalize": table_dtype != "F16", "hidden_dim": 256}))
    r = subprocess.run([binary, "--model-info", str(d)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = json.loads(r.stdout)
    m = StaticModel.from_pretrained(str(d))
    h = lambda a: f"{capi.fnv1a64(np.ascontiguousarray(a).tobytes()):016x}"
    assert got["V"] == V and got["normalize"] == m.normalize == (table_dtype != "F16")
    assert got["table_fnv"] == h(m.embeddings) and m.embeddings.dtype == np.float32
    if table_dtype == "F32":
        assert got["n_weights"] == 0 and got["n_mapping"] == 0 and m.weights is None and m.mapping is None
    else:
        assert got["n_weights"] == V and got["weights_fnv"] == h(m.weights)
        assert got["n_mapping"] == V and got["mapping_fnv"] == h(m.mapping)
    assert got["fingerprint"] == m.fingerprint() and got["fingerprint"].startswith(f"model2vec:{V}x256:n")
    # a directory without the table is an error, not a crash
    os.remove(d / "model.safetensors")
    r = subprocess.run([binary, "--model-info", str(d)], capture_output=True, text=True)
    assert r.returncode == 1 and "model.safetensors" in r.stderr


def test_damaged_commit_records_are_errors_not_crashes(binary, tmp_path):
    """A store.json that lost a key, has a field of the wrong type, a negative row count or a file name that
    leaves the store directory must produce a clear error in BOTH hosts (found by fuzzing the loaders under
    ASan / UBSan: the C++ loader used to dereference a missing key)."""
    from semtools_b200.workspace import Store
    d = tmp_path / "ws"
    r = subprocess.run([WSBIN, "store-selftest", str(d)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    meta_p = d / "flat.b200" / "store.json"
    good = json.loads(meta_p.read_text())
    cases = []
    for key in ("paths", "docs"):
        j = dict(good); del j[key]; cases.append(j)
        j = dict(good); j[key] = 7; cases.append(j)
    j = json.loads(json.dumps(good)); del j["docs"][0]["mtime"]; cases.append(j)
    j = json.loads(json.dumps(good)); j["docs"][0]["size_bytes"] = "big"; cases.append(j)
    j = json.loads(json.dumps(good)); j["paths"][0] = None; cases.append(j)
    j = dict(good); j["rows"] = -5; cases.append(j)
    j = dict(good); j["rows"] = "many"; cases.append(j)
    j = json.loads(json.dumps(good)); j["files"]["rows"] = "../../../etc/passwd"; cases.append(j)
    j = json.loads(json.dumps(good)); j["files"]["emb"] = 3; cases.append(j)
    for j in cases:
        meta_p.write_text(json.dumps(j))
        r = subprocess.run([WSBIN, "store-dump", str(d)], capture_output=True, text=True)
        assert r.returncode == 1 and r.stderr.startswith("Error:"), (j.keys(), r.returncode, r.stderr[:200])
        with pytest.raises(RuntimeError):
            Store.open(str(d))
    meta_p.write_text(json.dumps(good))                                 # and the intact record still loads
    assert subprocess.run([WSBIN, "store-dump", str(d)], capture_output=True).returncode == 0
    assert Store.open(str(d)).count_line_embeddings() == good["rows"]
