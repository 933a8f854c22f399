#!/usr/bin/env python
"""Per-source-line stall attribution from an .ncu-rep captured with --set full --import-source on (needs -lineinfo).
usage: python tools/ncu_stalls.py prof.ncu-rep <kernel-substring> [top-n]"""
import csv
import subprocess
import sys
from collections import defaultdict


def main(path, kern, top=40):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    cur_file = cur_fn = hdr = None
    agg, text, tot = defaultdict(lambda: defaultdict(float)), {}, defaultdict(lambda: defaultdict(float))
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            cur_fn = r[1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0] != "" and kern in (cur_fn or ""):
            try:
                int(r[0])
            except ValueError:
                continue
            k = (cur_fn, cur_file, int(r[0]))
            text[k] = r[1].strip()[:70]
            for name, val in zip(hdr, r):
                if name in ("Instructions Executed", "Thread Instructions Executed", "# Samples") or (name.startswith("stall_") and "Not Issued" not in name):
                    try:
                        v = float(val)
                    except ValueError:
                        continue
                    agg[k][name] += v
                    tot[cur_fn][name] += v
    for fn, t in tot.items():
        print("==", fn[:100], f"warp-inst {int(t['Instructions Executed']):,} lanes {t['Thread Instructions Executed'] / max(t['Instructions Executed'], 1):.1f} samples {int(t['# Samples'])}")
        st = {k: v for k, v in t.items() if k.startswith("stall_")}
        ssum = sum(st.values()) or 1
        print("  stall shares:", ", ".join(f"{k[6:]} {100 * v / ssum:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:10]))
        items = sorted(((k, v) for k, v in agg.items() if k[0] == fn), key=lambda kv: -kv[1]["# Samples"])[:top]
        for k, v in items:
            stl = sorted(((n[6:], x) for n, x in v.items() if n.startswith("stall_")), key=lambda kv: -kv[1])[:3]
            print(f"  {100 * v['# Samples'] / max(t['# Samples'], 1):5.1f}% smp {100 * v['Instructions Executed'] / max(t['Instructions Executed'], 1):5.1f}% inst "
                  f"lanes {v['Thread Instructions Executed'] / max(v['Instructions Executed'], 1):5.1f} {k[1]}:{k[2]:<4d} {text[k]:70s} | " + ", ".join(f"{n} {int(x)}" for n, x in stl))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
