"""CPU oracle for the semtools `search` hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, bench.py's cpu_baseline / `--impl reference` legs and
__graft_entry__.smoke() may import this package.  The product package
(semtools_b200) never does.  Numeric parity for the model2vec / simsimd
arithmetic is UNPINNED (see semtools_oracle.c header and DESIGN.md).
"""
from .oracle import *  # noqa: F401,F403
