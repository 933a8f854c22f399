#!/usr/bin/env python
"""Per-source-line instruction counts and SIMD efficiency from an .ncu-rep (needs -lineinfo).
usage: python tools/ncu_lines.py prof.ncu-rep [kernel-substring] [top-n]"""
import csv
import subprocess
import sys
from collections import defaultdict


def main(path, kern="", top=45):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    cur_file, cur_fn, hdr = None, None, None
    agg = defaultdict(lambda: [0, 0, 0, ""])   # (file,line) -> inst, thread inst, samples, text
    tot = defaultdict(lambda: [0, 0, 0])
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            cur_fn = r[1]
        elif r[0] == "Line No":
            hdr = r
            ii, ti, si = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        elif hdr and r[0] != "" and kern in (cur_fn or ""):
            try:
                inst, tinst, smp = int(r[ii]), int(r[ti]), int(r[si])
            except ValueError:
                continue
            k = (cur_fn, cur_file, int(r[0]))
            a = agg[k]
            a[0] += inst; a[1] += tinst; a[2] += smp; a[3] = r[1].strip()[:90]
            t = tot[cur_fn]
            t[0] += inst; t[1] += tinst; t[2] += smp
    for fn, t in tot.items():
        print(f"== {fn[:90]}: warp-inst {t[0]:,} thread-inst {t[1]:,} avg lanes {t[1] / max(t[0], 1):.1f} samples {t[2]}")
        items = [(k, v) for k, v in agg.items() if k[0] == fn]
        items.sort(key=lambda kv: -kv[1][0])
        for (f, file, line), v in items[:top]:
            print(f"  {100 * v[0] / t[0]:5.1f}% inst  {100 * v[2] / max(t[2], 1):5.1f}% smp  lanes {v[1] / max(v[0], 1):5.1f}  {file}:{line:<4d} {v[3]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "", int(sys.argv[3]) if len(sys.argv) > 3 else 45)
