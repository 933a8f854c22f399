#!/usr/bin/env python
"""Summarises an ncu launch list (`ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]
--clock-control none --csv --log-file X.csv <command>`): per kernel launches, total time, share, DRAM bytes.
usage: python tools/launch_summary.py X.csv ["header line"]"""
import csv
import re
import sys
from collections import defaultdict


def main(path, header=""):
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if not l.startswith("==")) if r]
    hdr = rows[0]
    ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    iid = hdr.index("ID")
    t, n, d = defaultdict(float), defaultdict(set), defaultdict(float)
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        name = re.sub(r"\(.*$", "", r[ik]).replace("void ", "").replace("zdr::<unnamed>::", "").replace("(bool)", "").replace("(int)", "")[:64]
        val = float(r[iv].replace(",", ""))
        if r[im] == "gpu__time_duration.sum":
            scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
            t[name] += val * scale
            n[name].add(r[iid])
        elif r[im].startswith("dram__bytes"):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[iu], 1.0)
            d[name] += val * scale
    total = sum(t.values())
    if header:
        print(header)
    print(f"total kernel time {total:.1f} ms over {sum(len(v) for v in n.values())} launches" + (f", DRAM {sum(d.values()) / 1e9:.1f} GB" if d else ""))
    for k in sorted(t, key=lambda k: -t[k]):
        print(f"  {k:64s} n={len(n[k]):4d} {t[k]:10.2f} ms {100 * t[k] / total:5.1f}%" + (f" {d[k] / 1e9:9.2f} GB" if d else ""))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
