"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals.

    python tools/summarize_launches.py gpurun_out/launches.csv "header line" > profiles/rN_launches.txt
"""
import csv
import sys


def main():
    path = sys.argv[1]
    header = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= iv:
            continue
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        unit = r[iu]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        rows.append((r[ik], us))
    tot = sum(u for _, u in rows)
    agg = {}
    for k, u in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += u
    if header:
        print(header)
    for k, (n, u) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:72]:72s} n={n:4d} us={u:10.1f} avg={u / n:8.1f} {100 * u / tot:5.1f}%")
    print(f"total us {tot:.2f}")


if __name__ == "__main__":
    main()
