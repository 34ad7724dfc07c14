#!/usr/bin/env python
"""SASS census of the shipped library: per kernel the code size and the counts of the instruction classes that matter on this
path (cuobjdump -sass, no GPU needed). usage: sass_summary.py [libafq.so] [out.json]"""
import collections, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "alevin_fry_b200", "libafq.so")
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "r2_sass_summary.json")
sys.path.insert(0, ROOT)
import bench
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", txt)), capture_output=True, text=True).stdout.splitlines()
kernels, cur, k = [], None, -1
classes = [("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("ATOMS", r"\bATOMS"), ("ATOMG/RED", r"\b(ATOMG|RED)\b"), ("LDS", r"\bLDS"), ("LDG", r"\bLDG"),
           ("STG", r"\bSTG"), ("BAR", r"\bBAR\b"), ("SHFL", r"\bSHFL"), ("VOTE", r"\bVOTE"), ("POPC", r"\bPOPC"), ("LDL/STL(local memory)", r"\b(LDL|STL)\b")]
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        k += 1
        cur = collections.OrderedDict(kernel=names[k], sass_bytes=0, instructions=0)
        for c, _ in classes: cur[c] = 0
        kernels.append(cur)
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur is not None:
        cur["instructions"] += 1
        cur["sass_bytes"] = int(m[1], 16) + 16
        for c, rx in classes:
            if re.search(rx, m[2]): cur[c] += 1
kernels.sort(key=lambda d: -d["sass_bytes"])
json.dump({"what": "cuobjdump -sass alevin_fry_b200/libafq.so (sm_100a), per kernel: code size and counts of the instruction classes that matter on this "
           "path; UBLKCP = cp.async.bulk (bulk copy global -> shared), SYNCS = mbarrier operations", "build": bench.build_hash(), "kernels": kernels},
          open(out, "w"), indent=1)
print("wrote", out, len(kernels), "kernels; build", bench.build_hash())
for d in kernels[:12]: print(d["kernel"][:60], d["sass_bytes"], d["UBLKCP"], d["LDL/STL(local memory)"])
