#!/usr/bin/env python
"""Summarise one GPU round (gpurun_out/<tag>/) into profiles/<tag>_*.{csv,md}: the ncu launch list
(per-kernel count / mean time / share of the step) and the key metrics of the `--set full` capture."""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

tag = sys.argv[1]
src = os.path.join("gpurun_out", tag)
os.makedirs("profiles", exist_ok=True)
out = [f"# GPU round {tag}", ""]

gpu = os.path.join(src, "gpu.txt")
if os.path.exists(gpu):
    out += ["```", open(gpu).read().strip(), "```", ""]

lf = os.path.join(src, "launches.csv")
if os.path.exists(lf):
    rows = list(csv.reader(open(lf, errors="replace")))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = defaultdict(list)
    for r in rows[h + 1:]:
        if len(r) > mv:
            d[r[kn]].append(float(r[mv].replace(",", "")))
    tot = sum(sum(v) for v in d.values())
    out += ["## ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`, cold-cache, serialised)", "",
            "| kernel | launches | mean us | share of profiled time |", "|---|---|---|---|"]
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| `{k[:90]}` | {len(v)} | {sum(v) / len(v) / 1e3:.2f} | {100 * sum(v) / tot:.1f}% |")
    out.append("")
    with open(os.path.join("profiles", f"{tag}_launches.csv"), "w") as f:
        f.write("kernel,launches,mean_us,share\n")
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"\"{k}\",{len(v)},{sum(v) / len(v) / 1e3:.3f},{sum(v) / tot:.4f}\n")

rep = os.path.join(src, "prof.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg"]
    idx = [hdr.index(w) for w in want if w in hdr]
    out += ["## ncu `--set full --clock-control none` capture of the top kernel", ""]
    for r in rows[2:]:
        out.append("| metric | value | unit |")
        out.append("|---|---|---|")
        for i in idx:
            out.append(f"| {hdr[i]} | {r[i][:100]} | {units[i]} |")
        out.append("")
    with open(os.path.join("profiles", f"{tag}_ncu_raw.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])

for name in sorted(os.listdir(src)):
    if name.startswith("bench") and name.endswith(".json"):
        txt = open(os.path.join(src, name)).read().strip()
        if txt:
            out += [f"## {name}", "", "```json", txt, "```", ""]
open(os.path.join("profiles", f"{tag}_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:40]))
