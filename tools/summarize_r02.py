#!/usr/bin/env python
"""gpurun_out/r02f_step_metrics.csv (ncu --metrics ... --profile-from-start off over tools/step_once.py: ONE training step in
bf16, then ONE in bf16x3) + gpurun_out/r02f_prof.ncu-rep (--set full of the MLP kernels) -> profiles/r02f_summary.md and
profiles/roofline_traffic.json.  usage: python tools/summarize_r02.py [tag]"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, OUT = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02f"
HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
lines = ["# ncu summary %s (B200; tools/gpu_profile_r02.sh)" % tag, ""]

rows = [r for r in csv.reader(open(os.path.join(G, "%s_step_metrics.csv" % tag))) if len(r) > 10]
hdr = rows[0]
ci = {h: i for i, h in enumerate(hdr)}
launches = collections.OrderedDict()
for r in rows[1:]:
    k = int(r[ci["ID"]])
    d = launches.setdefault(k, {"name": re.sub(r"\(.*", "", r[ci["Kernel Name"]]).replace("<unnamed>::", "").replace("void ", ""), "grid": r[ci["Grid Size"]],
                                "block": r[ci["Block Size"]]})
    d[r[ci["Metric Name"]]] = float(r[ci["Metric Value"]].replace(",", ""))
ls = list(launches.values())
# the two profiled steps: split at the second gather_batch launch
starts = [i for i, l in enumerate(ls) if l["name"].startswith("gather_batch")]
steps = {"bf16": ls[starts[0]:starts[1]], "bf16x3": ls[starts[1]:]} if len(starts) >= 2 else {"step": ls}
for mode, st in steps.items():
    tot = sum(l["gpu__time_duration.sum"] for l in st)
    agg = collections.OrderedDict()
    for l in st:
        a = agg.setdefault(l["name"], {"n": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0, "tensor": [], "warps": [], "regs": l["launch__registers_per_thread"]})
        a["n"] += 1; a["ns"] += l["gpu__time_duration.sum"]; a["rd"] += l["dram__bytes_read.sum"]; a["wr"] += l["dram__bytes_write.sum"]
        a["tensor"].append(l["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]); a["warps"].append(l["sm__warps_active.avg.pct_of_peak_sustained_active"])
    lines += ["## one %s training step, 4096 rays x (64+128): %d launches, %.1f us of kernel time" % (mode, len(st), tot / 1e3),
              "(per-launch device time is cold-cache and serialised under ncu: compare SHARES; DRAM bytes and pipe utilisation are per launch)", "",
              "| kernel | launches | us | share | DRAM read MB | DRAM write MB | DRAM GB/s | of %.0f GB/s | tensor pipe %% | warps active %% | regs |" % HBM,
              "|---|---|---|---|---|---|---|---|---|---|---|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        gbs = (a["rd"] + a["wr"]) / a["ns"]
        lines.append("| %s | %d | %.1f | %.1f%% | %.2f | %.2f | %.0f | %.2f | %s | %.0f | %d |" % (
            k, a["n"], a["ns"] / 1e3, 100 * a["ns"] / tot, a["rd"] / 1e6, a["wr"] / 1e6, gbs, gbs / HBM,
            "/".join("%.1f" % t for t in a["tensor"]) if max(a["tensor"]) > 0 else "-", sum(a["warps"]) / len(a["warps"]), a["regs"]))
    lines.append("")

rep = os.path.join(G, "%s_prof.ncu-rep" % tag)
traffic = {"source": "profiles/%s_summary.md (ncu --set full, fine pass 786 432 rows, same commit as the bench line)" % tag}
if os.path.isfile(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, units = rr[0], rr[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "sm__cycles_elapsed.max",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__grid_size"]
    data = rr[2:]
    names = [re.sub(r"\(.*", "", r[h.index("Kernel Name")]).replace("void ", "").replace("tc::", "") for r in data]
    lines += ["## ncu --set full, every MLP kernel launch of the two steps (coarse 262 144 rows, fine 786 432 rows)", "",
              "| metric | " + " | ".join("%s #%d" % (n, i) for i, n in enumerate(names)) + " |", "|---|" + "---|" * len(data)]
    for w in want:
        if w in h:
            i = h.index(w)
            lines.append("| %s [%s] | " % (w, units[i]) + " | ".join(r[i] for r in data) + " |")
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
    for n, r in zip(names, data):
        try:
            ir, iw = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
            tot = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
            key = re.sub(r"<.*>", "", n)
            traffic[key] = max(traffic.get(key, 0.0), tot)            # the fine-pass launch (largest)
        except Exception:
            pass
    lines += ["", "largest (fine-pass) DRAM traffic per launch, bytes: " + json.dumps({k: v for k, v in traffic.items() if k != "source"}), ""]
    json.dump(traffic, open(os.path.join(OUT, "roofline_traffic.json"), "w"), indent=1)
open(os.path.join(OUT, "%s_summary.md" % tag), "w").write("\n".join(lines) + "\n")
print("\n".join(lines)[:6000])
