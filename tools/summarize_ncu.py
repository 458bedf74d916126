#!/usr/bin/env python
"""Turns gpurun_out/{launches.csv, prof.ncu-rep, bench.json} into small tracked summaries under profiles/.
usage: python tools/summarize_ncu.py <tag>      e.g. r01a"""
import collections
import csv
import json
import re
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1]
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)
lines = ["# ncu summary %s" % tag, ""]

bj = os.path.join(G, "bench.json")
if os.path.isfile(bj) and os.path.getsize(bj):
    b = json.loads(open(bj).read().strip().splitlines()[-1])
    json.dump(b, open(os.path.join(out, "%s_bench.json" % tag), "w"), indent=1)
    lines += ["bench: value %.0f %s, e2e %.0f, ms/step %.2f, launches/steps %d/%d, clocks %s" % (
        b["value"], b["unit"], b["e2e"]["value"], b["ms_per_step"], b["gpu_launches"], b["steps"], b["clocks"]), ""]
    if b.get("roofline"):
        lines += ["per-kernel CUDA-event timings (fine pass, 4096x192 rows): " + json.dumps(b["roofline"]["kernels"]), ""]

lc = os.path.join(G, "launches.csv")
if os.path.isfile(lc):
    rows = list(csv.reader(open(lc)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hi + 2:] if len(r) == len(hdr)]
    names = [r[ci["Kernel Name"]] for r in data]
    dur = [float(r[ci["Metric Value"]].replace(",", "")) for r in data]
    gi = [i for i, n in enumerate(names) if "gather_batch" in n]
    a, b_ = gi[1], gi[2]
    agg = collections.OrderedDict()
    for n, d in zip(names[a:b_], dur[a:b_]):
        k = n.split("(")[0]
        agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += d
    tot = sum(v[1] for v in agg.values())
    lines += ["## launch list of ONE steady-state training step (ncu --metrics gpu__time_duration.sum --clock-control none;",
              "cold-cache, serialised: compare SHARES)", "", "| kernel | launches | us | share |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("| %s | %d | %.1f | %.1f%% |" % (k, v[0], v[1] / 1e3, 100 * v[1] / tot))
    lines += ["| **total** | %d | %.1f | |" % (b_ - a, tot / 1e3), ""]
    with open(os.path.join(out, "%s_launches.csv" % tag), "w") as f:
        w = csv.writer(f); w.writerow(["kernel", "duration_ns"])
        for n, d in zip(names[a:b_], dur[a:b_]):
            w.writerow([n.split("(")[0], int(d)])

rep = os.path.join(G, "prof.ncu-rep")
if os.path.isfile(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
    lines += ["## ncu --set full (fine pass: 786 432 rows)", "", "| metric | " + " | ".join(r[hdr.index("Kernel Name")].split("(")[0] for r in rows[2:]) + " |",
              "|---|" + "---|" * len(rows[2:])]
    traffic = {}
    for wname in want[1:]:
        if wname in hdr:
            i = hdr.index(wname)
            lines.append("| %s [%s] | " % (wname, units[i]) + " | ".join(r[i] for r in rows[2:]) + " |")
    for r in rows[2:]:
        try:
            rd, wr = float(r[hdr.index("dram__bytes_read.sum")]), float(r[hdr.index("dram__bytes_write.sum")])
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
            ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
            name = r[hdr.index("Kernel Name")].split("(")[0].replace("tc::", "").replace("void ", "")
            traffic[re.sub(r"<.*>", "", name)] = rd * scale[ur] + wr * scale[uw]     # template arguments dropped
        except Exception:
            pass
    json.dump(traffic, open(os.path.join(out, "roofline_traffic.json"), "w"), indent=1)
    lines += ["", "dram traffic per launch (bytes): " + json.dumps(traffic), ""]
open(os.path.join(out, "%s_summary.md" % tag), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
