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
