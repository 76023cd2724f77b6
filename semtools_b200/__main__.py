"""`python -m semtools_b200 {search,workspace} ...` -- the flag surface of
reference src/bin/semtools.rs:29-132 for the two in-scope sub-commands.
The model directory (tokenizer.json, model.safetensors, config.json) comes from
$SEMTOOLS_B200_MODEL_DIR (the reference downloads MODEL_NAME from the hub instead)."""
import argparse
import os
import sys

from . import cmds
from .workspace import _rust_lines
from .model import MODEL_NAME, StaticModel


def build_parser():
    p = argparse.ArgumentParser(prog="semtools")
    sub = p.add_subparsers(dest="cmd", required=True)
    s = sub.add_parser("search", help="A CLI tool for fast semantic keyword search")
    s.add_argument("query")
    s.add_argument("files", nargs="*", help="Files to search, optional if using stdin")
    s.add_argument("-n", "--n-lines", "--context", dest="n_lines", type=int, default=3)
    s.add_argument("--top-k", dest="top_k", type=int, default=3)
    s.add_argument("-m", "--max-distance", "--threshold", dest="max_distance", type=float, default=None)
    s.add_argument("-i", "--ignore-case", dest="ignore_case", action="store_true")
    s.add_argument("-j", "--json", action="store_true")
    s.add_argument("-w", "--workspace", default=None)
    w = sub.add_parser("workspace", help="Manage semtools workspaces")
    w.add_argument("-j", "--json", action="store_true")
    ws = w.add_subparsers(dest="wcmd", required=True)
    u = ws.add_parser("use"); u.add_argument("name")
    st = ws.add_parser("status"); st.add_argument("name", nargs="?", default=None)
    pr = ws.add_parser("prune"); pr.add_argument("name", nargs="?", default=None)
    for x in (u, st, pr):
        x.add_argument("-j", "--json", action="store_true", dest="json_sub")
    return p


def main(argv=None) -> int:
    a = build_parser().parse_args(argv)
    if a.cmd == "workspace":
        js = a.json or getattr(a, "json_sub", False)
        if a.wcmd == "use":
            return cmds.workspace_use_cmd(a.name, js)
        if a.wcmd == "status":
            return cmds.workspace_status_cmd(js, a.name)
        return cmds.workspace_prune_cmd(js, a.name)
    model_dir = os.environ.get("SEMTOOLS_B200_MODEL_DIR")
    if not model_dir:
        sys.stderr.write(f"Error: set SEMTOOLS_B200_MODEL_DIR to a local copy of {MODEL_NAME} "
                         "(tokenizer.json, model.safetensors, config.json); hub download is unavailable offline\n")
        return 1
    model = StaticModel.from_pretrained(model_dir)
    stdin_tty = sys.stdin.isatty()
    # io::stdin().lock().lines() (cmds/search.rs:160-163): BufRead::lines splits on '\n' and drops one
    # '\r' before it; a bare '\r' is text.  Read bytes: text mode would translate newlines.
    lines = None
    if not (stdin_tty or a.files):
        lines = _rust_lines(sys.stdin.buffer.read().decode("utf-8"))       # same rules as str::lines()
    return cmds.search_cmd(a.query, a.files, a.n_lines, a.top_k, a.max_distance, a.ignore_case, a.json, a.workspace,
                           model, stdin_lines=lines, stdin_is_tty=stdin_tty, stdout_is_tty=sys.stdout.isatty())


if __name__ == "__main__":
    sys.exit(main())
