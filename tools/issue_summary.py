#!/usr/bin/env python
"""Instruction-issue and DRAM summary of ONE bench step from an ncu launch list.

    ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum \\
        --clock-control none --csv --log-file X.csv python bench.py --config c3 --spp 16 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline
    python tools/issue_summary.py X.csv BENCH.json profiles/r02a_issue_c3.json [--text profiles/r02a_issue_c3.txt]

BENCH.json = the JSON line the SAME command printed when run without ncu (ray count of a step, config, schedule).
The step is the library's launches after the last L2-flush fill (bench.py writes a 256 MiB buffer before every timed
step); per kernel: launches, time under ncu, warp instructions, lanes per warp instruction, DRAM bytes.  bench.py reads
the JSON this writes for its `roofline` (bound = warp-instruction issue) and scales it by the live ray count."""
import csv
import json
import re
import sys
from collections import defaultdict


def clean(name):
    return re.sub(r"\(.*$", "", name).replace("void ", "").replace("zdr::<unnamed>::", "").replace("(bool)", "").replace("(int)", "")


def main():
    csv_path, bench_path, out_path = sys.argv[1:4]
    text_path = sys.argv[sys.argv.index("--text") + 1] if "--text" in sys.argv else None
    bench = json.loads([l for l in open(bench_path) if l.startswith("{")][-1])
    rows = [r for r in csv.reader(l for l in open(csv_path, errors="replace") if not l.startswith("==")) if r]
    hdr = rows[0]
    ik, im, iv, iu, iid = [hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID")]
    by_id, order = {}, []
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        i = int(r[iid])
        if i not in by_id:
            by_id[i] = {"raw": r[ik], "name": clean(r[ik]), "t": 0.0, "inst": 0.0, "tinst": 0.0, "dram": 0.0}
            order.append(i)
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        k = by_id[i]
        if r[im] == "gpu__time_duration.sum":
            k["t"] += v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
        elif r[im] == "smsp__inst_executed.sum":
            k["inst"] += v
        elif r[im] == "smsp__thread_inst_executed.sum":
            k["tinst"] += v
        elif r[im].startswith("dram__bytes"):
            k["dram"] += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[iu], 1.0)
    seq = [by_id[i] for i in order]
    last_flush = max((n for n, k in enumerate(seq) if "FillFunctor<unsigned char>" in k["raw"]), default=-1)
    step = [k for k in seq[last_flush + 1:] if k["name"].startswith("k_")]
    agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
    for k in step:
        a = agg[k["name"]]
        a[0] += 1; a[1] += k["t"]; a[2] += k["inst"]; a[3] += k["tinst"]; a[4] += k["dram"]
    tot = [sum(a[j] for a in agg.values()) for j in range(5)]
    rays = bench["config"]["rays_per_step"]
    rec = {"config": bench["config"].get("config"), "schedule": bench["config"].get("schedule", "auto"), "workload": bench["config"]["workload"],
           "rays_per_step": rays, "launches_per_step": tot[0], "ms_under_ncu": round(tot[1], 3),
           "warp_inst_per_step": tot[2], "thread_inst_per_step": tot[3], "lanes_per_warp_inst": round(tot[3] / max(tot[2], 1), 2),
           "dram_bytes_per_step": tot[4],
           "kernels": {n: {"launches": a[0], "ms_under_ncu": round(a[1], 3), "warp_inst": a[2], "lanes": round(a[3] / max(a[2], 1), 2),
                           "dram_bytes": a[4]} for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1])},
           "source": f"{csv_path} (ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_*.sum --clock-control none)"}
    json.dump(rec, open(out_path, "w"), indent=1)
    lines = [f"{rec['workload']} [schedule {rec['schedule']}]",
             f"one step: {tot[0]} launches, {tot[1]:.2f} ms under ncu, {tot[2] / 1e9:.3f} G warp instructions ({tot[2] / max(rays, 1):.1f} per ray, "
             f"{rec['lanes_per_warp_inst']} lanes each), DRAM {tot[4] / 1e9:.2f} GB ({tot[4] / max(rays, 1):.0f} B per ray), {rays} rays"]
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"   {n:40s} n={a[0]:4d} {a[1]:9.3f} ms ({100 * a[1] / max(tot[1], 1e-9):4.1f}%) {a[2] / 1e6:10.1f} M warp-inst  {a[3] / max(a[2], 1):5.1f} lanes {a[4] / 1e9:8.3f} GB")
    txt = "\n".join(lines) + "\n"
    if text_path:
        open(text_path, "w").write(txt)
    print(txt, end="")


if __name__ == "__main__":
    main()
