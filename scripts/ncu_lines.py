#!/usr/bin/env python
"""Per-source-line instruction totals from an ncu report (needs --import-source on, -lineinfo).
usage: ncu_lines.py report.ncu-rep 'k_resolve_smem<(int)3>' [file-substring] [top N]"""
import csv, subprocess, sys, collections

rep, func = sys.argv[1], sys.argv[2]
fsub = sys.argv[3] if len(sys.argv) > 3 else ""
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = collections.OrderedDict()
cur_file = cur_func = None
hdr = None
line = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path": cur_file = r[1]; continue
    if r[0] == "Function Name": cur_func = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if func not in (cur_func or "") or fsub not in (cur_file or ""):
        continue
    if r[0]:
        line = (cur_file.split("/")[-1], int(r[0]), r[1].strip())
        agg.setdefault(line, [0, 0, 0])
        continue
    if line is None or r[2] == "...":
        continue
    ie, te, smp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    try:
        agg[line][0] += int(r[ie]); agg[line][1] += int(r[te]); agg[line][2] += int(r[smp])
    except ValueError:
        pass
tot = sum(v[0] for v in agg.values()) or 1
tots = sum(v[2] for v in agg.values()) or 1
print(f"total warp instr {tot}  samples {tots}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    lanes = v[1] / v[0] if v[0] else 0
    print(f"{100*v[0]/tot:5.1f}% instr {100*v[2]/tots:5.1f}% smp lanes {lanes:5.1f}  {k[0]}:{k[1]}  {k[2][:90]}")
