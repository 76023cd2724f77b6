"""ask search-tool adapter (reference src/ask/tools.rs:34-94,208-259): chunk formatting."""
import pytest

from semtools_b200 import ask_tool
from semtools_b200.search import RankedLine, SearchConfig, SearchResult


def test_format_search_results_chunks():
    rs = [SearchResult("a.txt", ["l0", "l1", "l2"], 0, 3, 1, 0.25),
          SearchResult("b.txt", ["x"], 4, 5, 4, 1.0)]
    assert ask_tool.format_search_results(rs) == (
        "<chunk file=a.txt start=0 end=3 distance=0.25>\nl0\nl1\nl2\n</chunk>\n"
        "<chunk file=b.txt start=4 end=5 distance=1>\nx\n</chunk>\n")          # f64 Display: 1.0 -> "1"
    assert ask_tool.format_search_results([]) == ""


def test_format_ranked_lines_reads_context_and_keeps_unclamped_end(tmp_path):
    p = tmp_path / "f.txt"
    p.write_text("a\nb\r\nc\nd\n")
    out = ask_tool.format_ranked_lines([RankedLine(str(p), 3, 0.5), RankedLine(str(p), 0, 0.1)], 2)
    assert out == (f"<chunk file={p} start=1 end=6 distance=0.5>\nb\nc\nd\n</chunk>\n"
                   f"<chunk file={p} start=0 end=3 distance=0.1>\na\nb\nc\n</chunk>\n")
    gone = tmp_path / "gone.txt"
    assert ask_tool.format_ranked_lines([RankedLine(str(gone), 0, 0.0)], 1) == \
        f"<chunk file={gone} start=0 end=2 distance=0>\n[Error: Could not read file content]</chunk>\n"
    with pytest.raises(IndexError):                                              # the reference panics here
        ask_tool.format_ranked_lines([RankedLine(str(p), 9, 0.0)], 1)


def test_search_tool_no_input():
    with pytest.raises(RuntimeError, match="No input provided"):
        ask_tool.search_tool([], "q", None, SearchConfig(3, 3, None, False), [])
