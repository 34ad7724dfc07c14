#!/usr/bin/env python
"""Summarise an ncu report (--set full) as the JSON kept under profiles/: per kernel the launch shape, time,
DRAM bytes, L2 hit rate, issue / occupancy figures, plus (with --import-source on) the top source lines by
stall samples of one kernel.
usage: ncu_summary.py report.ncu-rep out.json "capture description" "reading" [kernel-substring-for-lines]"""
import csv, json, subprocess, sys, re

rep, out, capture, reading = sys.argv[1:5]
ksub = sys.argv[5] if len(sys.argv) > 5 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
kernels = []
for r in rows[2:]:
    k = {}
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            k[w + (" [%s]" % units[i] if units[i] else "")] = r[i]
    kernels.append(k)
doc = {"capture": capture, "reading": reading, "kernels": kernels}
if ksub:
    src = subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_lines.py"), rep, ksub, "", "400"],
                         capture_output=True, text=True).stdout
    lines = []
    for l in src.splitlines():
        m = re.match(r"\s*([\d.]+)% instr\s+([\d.]+)% smp lanes\s+([\d.]+)\s+(\S+)\s+(.*)", l)
        if m:
            lines.append({"pct_warp_instr": float(m[1]), "pct_stall_samples": float(m[2]), "active_lanes": float(m[3]),
                          "where": m[4], "source": m[5][:110]})
        elif l.startswith("total"):
            doc["lines_total"] = l.strip()
    lines.sort(key=lambda d: -d["pct_stall_samples"])
    doc["top_lines_by_stall_samples(" + ksub + ")"] = lines[:25]
json.dump(doc, open(out, "w"), indent=1)
print("wrote", out, len(kernels), "kernels")
