#!/usr/bin/env python
"""Turns the files a `tools/gpu_final.sh <R>` session left in gpurun_out/ into the tracked summaries under profiles/:
launch-list summary of the default bench command, per-stage shares with DRAM bytes of one forward / one adjoint pass
(16 spp = one batch) + profiles/traffic.json (feeds bench.py's roofline.traffic), the --set full summary, bench lines.
usage: python tools/make_profiles.py r01v"""
import csv
import json
import os
import re
import shutil
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def clean(name):
    return re.sub(r"\(.*$", "", name).replace("void ", "").replace("zdr::<unnamed>::", "").replace("(bool)", "").replace("(int)", "")


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if not l.startswith("==")) if r]
    hdr = rows[0]
    ik, im, iv, iu, iid = [hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID")]
    by_id, order = {}, []
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        i = int(r[iid])
        if i not in by_id:
            by_id[i] = {"name": clean(r[ik]), "t": 0.0, "d": 0.0}
            order.append(i)
        v = float(r[iv].replace(",", ""))
        if r[im] == "gpu__time_duration.sum":
            by_id[i]["t"] += v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r[iu]]
        else:
            by_id[i]["d"] += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r[iu]]
    return [by_id[i] for i in order]


# 1. launch list of the default bench command
shutil.copy(os.path.join(G, f"{R}_launches_default.csv"), os.path.join(P, f"{R}_launches_default.csv"))
out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), os.path.join(P, f"{R}_launches_default.csv"),
                      "python bench.py --no-cpu-baseline (C3, 1024^2, 256 spp: stats pass + 2 warm-up + 3 timed + 3 e2e steps), "
                      "ncu --metrics gpu__time_duration.sum --clock-control none"], capture_output=True, text=True).stdout
open(os.path.join(P, f"{R}_launches_default_summary.txt"), "w").write(out)

# 2. one forward and one adjoint pass at 16 spp, with DRAM bytes
shutil.copy(os.path.join(G, f"{R}_launches_spp16.csv"), os.path.join(P, f"{R}_launches_wavefront_spp16.csv"))
seq = launches(os.path.join(P, f"{R}_launches_wavefront_spp16.csv"))
passes, cur = [], None
for k in seq:
    if k["name"].startswith("k_wf_generate"):
        cur = {"kind": "bwd" if "<1>" in k["name"] else "fwd", "k": []}
        passes.append(cur)
    if cur is not None and k["name"].startswith("k_wf_"):
        cur["k"].append(k)
merged = []                       # the batches of one render call (two lanes: two half-batches) count as one pass
for p_ in passes:
    if merged and merged[-1]["kind"] == p_["kind"]:
        merged[-1]["k"] += p_["k"]
    else:
        merged.append({"kind": p_["kind"], "k": list(p_["k"])})
passes = merged
plain = [p for p in passes if not any(re.match(r"k_wf_trace<\d, 1", k["name"]) for k in p["k"])]      # not the counting pass
fw, bw = [p for p in plain if p["kind"] == "fwd"][-1], [p for p in plain if p["kind"] == "bwd"][-1]
lines = []
for title, p in (("forward pass", fw), ("adjoint pass", bw)):
    tot, dr = sum(k["t"] for k in p["k"]), sum(k["d"] for k in p["k"])
    lines.append(f"{title} (1024^2, 16 spp, {sum(1 for k in p['k'] if k['name'].startswith('k_wf_generate'))} batch(es)): total {tot:.2f} ms under ncu, DRAM {dr / 1e9:.2f} GB, {len(p['k'])} launches")
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for k in p["k"]:
        a = agg[k["name"]]
        a[0] += 1; a[1] += k["t"]; a[2] += k["d"]
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"   {n:34s} n={a[0]:3d} {a[1]:8.3f} ms ({100 * a[1] / tot:4.1f}%) {a[2] / 1e9:8.2f} GB")
open(os.path.join(P, f"{R}_wavefront_stage_shares.txt"), "w").write("\n".join(lines) + "\n")
json.dump({"res": 1024, "spp": 16, "integrator": "path", "schedule": "wavefront",
           "forward_dram_bytes": sum(k["d"] for k in fw["k"]), "backward_dram_bytes": sum(k["d"] for k in bw["k"]),
           "forward_ms_under_ncu": sum(k["t"] for k in fw["k"]), "backward_ms_under_ncu": sum(k["t"] for k in bw["k"]),
           "source": f"profiles/{R}_launches_wavefront_spp16.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum)"},
          open(os.path.join(P, "traffic.json"), "w"), indent=1)

# 3. --set full summary, bench lines
rep = os.path.join(G, f"prof_{R}_wf.ncu-rep")
if os.path.exists(rep):
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    open(os.path.join(P, f"{R}_wf_ncu_summary.txt"), "w").write(txt)
for src, dst in ((f"{R}_bench.json", f"{R}_bench_full.json"), (f"{R}_bench_reference.json", f"{R}_bench_reference.json")):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))
print("\n".join(lines))
