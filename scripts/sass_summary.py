#!/usr/bin/env python3
"""Per-kernel SASS mnemonic counts of the shipped library -> profiles/rNN_sass_summary.txt
    python scripts/sass_summary.py [lib.so] > profiles/r02_sass_summary.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "semtools_b200/lib/libsemtools_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, errors="ignore").stdout
pat = {"UTCHMMA (tcgen05.mma)": r"\bUTCHMMA", "LDTM (tcgen05.ld)": r"\bLDTM", "UBLKCP (cp.async.bulk)": r"\bUBLKCP",
       "UTCBAR (tcgen05.commit)": r"\bUTCBAR", "SYNCS (mbarrier)": r"\bSYNCS", "ACQBULK/PREEXIT (griddepcontrol, PDL)": r"\b(ACQBULK|PREEXIT)",
       "IDP.4A (dp4a)": r"\bIDP\.4A", "LDG.E.128": r"LDG\.E\.(NA\.)?128|LDG\.E\.128", "HMMA (legacy mma.sync)": r"\bHMMA",
       "DFMA (f64 re-rank)": r"\bDFMA", "ATOMG/RED (global atomics)": r"\b(ATOMG|RED)\b"}
cur, cnt = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); cnt[cur] = collections.Counter(); continue
    if cur:
        for k, r in pat.items():
            if re.search(r, line):
                cnt[cur][k] += 1
print("# per-kernel SASS mnemonic counts of", lib, "(cuobjdump -sass; all code objects sm_100a)\n")
tot = collections.Counter()
for fn, c in cnt.items():
    if not c:
        continue
    name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
    print(re.sub(r"\(.*", "", name)[:90])
    print("    " + ", ".join(f"{k}: {c[k]}" for k in pat if c[k]))
    tot.update(c)
print("\nTOTAL  " + ", ".join(f"{k}: {tot[k]}" for k in pat))
