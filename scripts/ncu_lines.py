#!/usr/bin/env python
"""Per-source-line hot spots of one kernel from an .ncu-rep (needs -lineinfo and --import-source on).

usage: python scripts/ncu_lines.py report.ncu-rep [top_n]
"""
import csv, subprocess, sys

def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = next(r for r in rows if r and r[0] == "Line No")
    ix = {}
    for i, n in enumerate(hdr):
        ix.setdefault(n, i)
    lines = [r for r in rows if len(r) == len(hdr) and r[0].isdigit()]
    num = lambda r, k: int(r[ix[k]]) if r[ix[k]].lstrip("-").isdigit() else 0
    ts = sum(num(r, "# Samples") for r in lines)
    ti = sum(num(r, "Instructions Executed") for r in lines)
    print(f"total samples {ts}  warp instructions {ti}")
    lines.sort(key=lambda r: -num(r, "# Samples"))
    print(f"{'line':>5} {'samp%':>6} {'inst%':>6} {'thr':>5}  source")
    for r in lines[:top]:
        print(f"{r[0]:>5} {100.0 * num(r, '# Samples') / max(ts, 1):6.2f} {100.0 * num(r, 'Instructions Executed') / max(ti, 1):6.2f} "
              f"{r[ix['Avg. Threads Executed']]:>5}  {r[1].strip()[:100]}")

if __name__ == "__main__":
    main()
