"""Static SASS evidence: which of this library's kernels contain tcgen05 / TMEM / TMA / legacy-HMMA instructions.

    python tools/sass_evidence.py > profiles/r1_sass_evidence.md      (needs cuobjdump; no GPU)

Mnemonics per /opt/skills/guides/B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM,
cp.async.bulk.tensor -> UTMALDG/UTMASTG, mma.sync -> HMMA, ldmatrix -> LDSM, cp.async -> LDGSTS, red.global -> REDG."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multimodalrouting_b200", "csrc", "libmmr_b200.so")
COLS = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "LDSM", "LDGSTS", "REDG"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    cur, counts, total = None, collections.defaultdict(collections.Counter), collections.Counter()
    pat = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)")
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = pat.match(line)
        if m and cur:
            total[cur] += 1
            if m.group(1) in COLS:
                counts[cur][m.group(1)] += 1
    names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    rows = []
    for mangled, name in zip(counts, names):
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("mmr::", "")
        rows.append((name, total[mangled], counts[mangled]))
    rows.sort(key=lambda r: (-(r[2]["UTCHMMA"] > 0), -(r[2]["HMMA"] > 0), r[0]))
    print("# r1 — static SASS evidence (`cuobjdump -sass multimodalrouting_b200/csrc/libmmr_b200.so`, sm_100a)\n")
    print("Static instruction counts per kernel (not executed counts).  `UTCHMMA` = `tcgen05.mma`, `LDTM` = `tcgen05.ld`,")
    print("`UTMALDG`/`UTMASTG` = TMA tensor loads / stores, `UTCBAR` = `tcgen05.commit`, `SYNCS` = mbarrier ops, `HMMA` =")
    print("`mma.sync`, `LDSM` = `ldmatrix`, `LDGSTS` = `cp.async`, `REDG` = `red.global` (split-K accumulation).  Regenerate with")
    print("`python tools/sass_evidence.py`.\n")
    print("| kernel | SASS instrs | " + " | ".join(COLS) + " |")
    print("|---|---|" + "---|" * len(COLS))
    for name, n, c in rows:
        if not any(c[k] for k in ("UTCHMMA", "HMMA", "UTMALDG")):
            continue
        print(f"| `{name}` | {n} | " + " | ".join(str(c[k]) if c[k] else "" for k in COLS) + " |")


if __name__ == "__main__":
    sys.exit(main())
