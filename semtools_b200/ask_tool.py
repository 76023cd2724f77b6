"""`ask` agent's search tool adapter (reference src/ask/tools.rs:34-94, 208-259) over this
package's search path.  The LLM front end itself (src/ask/*, OpenAI client) is out of scope;
this is the piece of it that sits on the search path: it runs the same search and wraps the
results in `<chunk file=… start=… end=… distance=…>` blocks.  No new compute.
"""
from __future__ import annotations

from . import capi
from .cmds import rust_display_f32, rust_display_f64, search_files
from .search import SearchConfig
from .workspace import Workspace, _rust_lines, search_with_workspace

NO_INPUT = "Error: No input provided. Either specify files as arguments or pipe input to stdin."


def format_search_results(results) -> str:                              # tools.rs:34-55
    out = []
    for r in results:
        out.append(f"<chunk file={r.filename} start={r.start} end={r.end} distance={rust_display_f64(r.distance)}>\n")
        out.extend(f"{line}\n" for line in r.lines)
        out.append("</chunk>\n")
    return "".join(out)


def format_ranked_lines(ranked_lines, n_lines: int) -> str:             # tools.rs:58-94
    out = []
    for rl in ranked_lines:
        m = int(rl.line_number)
        start = max(m - n_lines, 0)                                       # saturating_sub
        end = m + n_lines + 1                                             # NOT clamped in the header (:69-73)
        out.append(f"<chunk file={rl.path} start={start} end={end} distance={rust_display_f32(rl.distance)}>\n")
        try:
            with open(rl.path, encoding="utf-8") as f:
                lines = _rust_lines(f.read())
        except (OSError, UnicodeDecodeError):
            out.append("[Error: Could not read file content]")            # no newline, as the reference (:87)
        else:
            actual_end = min(end, len(lines))
            if start > actual_end:                                        # lines[start..end] panics in the reference
                raise IndexError(f"slice index starts at {start} but ends at {actual_end}")
            out.extend(f"{line}\n" for line in lines[start:actual_end])
        out.append("</chunk>\n")
    return "".join(out)


def search_tool(files, query: str, model, config: SearchConfig, files_searched: list, workspace_name=None) -> str:
    """SearchTool::search (tools.rs:208-259).  Raises RuntimeError(NO_INPUT) on no files."""
    if config.ignore_case:
        query = query.lower()
    if not files:
        raise RuntimeError(NO_INPUT)
    try:
        Workspace.active(workspace_name)
        in_ws = True
    except RuntimeError:
        in_ws = False
    if in_ws:
        if model.ctx is None:
            model.ctx = capi.Context(0)
        ranked = search_with_workspace(files, model.encode_single(query), lambda ls: model.encode_with_args(ls, 2048, 16384),
                                       config, workspace_name, ctx=model.ctx)
        for rl in ranked:
            if rl.path not in files_searched:
                files_searched.append(rl.path)
        return format_ranked_lines(ranked, config.n_lines)
    results = search_files(files, query, model, config)
    for r in results:
        if r.filename not in files_searched:
            files_searched.append(r.filename)
    return format_search_results(results)
