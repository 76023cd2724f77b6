"""semtools_b200 -- B200-native `search` hot path of run-llama/semtools.

The product is the C-ABI library `lib/libsemtools_b200.so` (hand-written CUDA
for sm_100a, see include/semtools_b200.h).  This package is the thin host-side
mirror used by the tests and bench: `capi` binds the C ABI with ctypes, `search`
mirrors the reference's `search_documents` / `Document` / `SearchConfig`
interface (reference src/search/mod.rs:18-120) on top of it.

There is no CPU fallback anywhere in this package: importing `capi` without the
built library, or creating a context without a B200, raises.
"""
from . import capi  # noqa: F401
from .search import (Document, SearchConfig, SearchResult, Searcher,  # noqa: F401
                     RankedLine)

__all__ = ["capi", "Document", "SearchConfig", "SearchResult", "Searcher", "RankedLine"]
