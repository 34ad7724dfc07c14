#!/usr/bin/env python
"""Turn the captures of scripts/gpu_profiles_r2.sh (gpurun_out/<tag>_*) into the committed evidence under profiles/:
ncu --set full summaries, launch lists + per-kernel shares, and ncu_traffic.json (DRAM bytes of one full-size step, tied to the
sha256 of csrc/ it was captured on — bench.py reports roofline.traffic only when that matches the sources it runs).
usage: profiles_r2.py <tag> <build hash of the captured tree>"""
import csv, json, os, re, subprocess, sys, collections, shutil

tag, build = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.environ.get("PROFILES_OUT", os.path.join(ROOT, "profiles"))
os.makedirs(P, exist_ok=True)


def launches(fn):
    rows = list(csv.reader(l for l in open(fn, errors="replace") if l.startswith('"')))
    h = rows[0]
    ik, im, iv = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    out = []
    for r in rows[1:]:
        try:
            out.append((r[ik], r[im], float(r[iv].replace(",", ""))))
        except (ValueError, IndexError):
            pass
    return out


for cfg in ("c2", "c3"):
    fn = os.path.join(G, f"{tag}_launches_{cfg}.csv")
    if not os.path.exists(fn):
        continue
    shutil.copy(fn, os.path.join(P, f"{tag}_ncu_launches_bench_{cfg}_20kcells.csv"))
    tot = collections.OrderedDict()
    for k, m, v in launches(fn):
        if m == "gpu__time_duration.sum":
            k = re.sub(r"\(.*", "", k)
            tot.setdefault(k, [0, 0.0]); tot[k][0] += 1; tot[k][1] += v
    s = sum(v[1] for v in tot.values()) or 1
    with open(os.path.join(P, f"{tag}_ncu_launch_shares_{cfg}.csv"), "w") as f:
        f.write("kernel,launches,total_us,share\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{v[0]},{v[1] / 1000:.1f},{v[1] / s:.4f}\n")

FULL = [("full_resolve", "k_resolve_smem<(int)3>", "bench.py --steps 1 --warmup 1 --cells 10000 (C2, cr-like)", "k_resolve_smem<0..5> as shipped at the end of round 2 (design unchanged since round 1)"),
        ("full_pug", "k_pug_build<(int)0>", "bench.py --config C3 --steps 1 --warmup 1 --cells 10000 (parsimony)", "the split parsimony path as shipped: k_pug_build<0..3> (bulk-copy loads), k_pug_cover2 / _g<4> / _g<8> / _w, k_pug_count"),
        ("full_em_c4", "k_em_cells<(int)1>", "bench.py --config C4 --steps 1 --warmup 1 --cells 8000 (USA cr-like-em)", "stage B (k_pug_back) and stage C (k_em_cells) of the EM resolutions on the split path, C4"),
        ("full_em_c5", "k_em_cells<(int)0>", "bench.py --config C5 --steps 1 --warmup 1 --cells 10000 (parsimony-em)", "stage B (k_pug_back) and stage C (k_em_cells) of the EM resolutions on the split path, C5")]
for name, ksub, cmd, reading in FULL:
    rep = os.path.join(G, f"{tag}_{name}.ncu-rep")
    if os.path.exists(rep):
        subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep, os.path.join(P, f"{tag}_ncu_{name}.json"),
                        "ncu --set full --clock-control none --import-source on, " + cmd + " (scripts/gpu_profiles_r2.sh), csrc build " + build, reading, ksub], check=False)

traffic = {"build": build, "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none over the per-cell "
           "resolve family of ONE full-size device-resident step (the 4th of the run; scripts/gpu_profiles_r2.sh); bytes per step = sum over the launches", "per_launch": {}}
for cfg, key in (("c2", "C2"), ("c3", "C3")):
    fn = os.path.join(G, f"{tag}_traffic_{cfg}.csv")
    if not os.path.exists(fn):
        continue
    order = []
    cur = None
    for k, m, v in launches(fn):
        if m == "dram__bytes_read.sum":
            cur = {"kernel": re.sub(r"\(.*", "", k), "dram_read_bytes": 0.0, "dram_write_bytes": 0.0}
            order.append(cur)
        if cur is None:
            continue
        if m == "dram__bytes_read.sum": cur["dram_read_bytes"] = v
        elif m == "dram__bytes_write.sum": cur["dram_write_bytes"] = v
        elif m == "gpu__time_duration.sum": cur["us(cold, serialised)"] = v / 1000.0
        elif m == "smsp__inst_executed.sum": cur["warp_instructions"] = v
    # the capture holds every launch of the run: split it into steps (every step's first launch of the family is
    # k_resolve_smem<5>) and keep the 4th — the timed device-resident full-size step behind the 3 warm-ups
    steps_ = []
    for d in order:
        if d["kernel"].endswith("k_resolve_smem<5>") or not steps_:
            steps_.append([])
        steps_[-1].append(d)
    print(key, "steps in the capture:", [len(x) for x in steps_][:12])
    order = steps_[3] if len(steps_) > 3 else []
    traffic[key] = sum(d["dram_read_bytes"] + d["dram_write_bytes"] for d in order)
    traffic.setdefault("warp_instructions", {})[key] = sum(d.get("warp_instructions", 0.0) for d in order)
    traffic.setdefault("cells_per_step", {})[key] = {"C2": 100000, "C3": 125000}[key]
    traffic["per_launch"][key] = order
for key, order in traffic["per_launch"].items():
    names = [d["kernel"] for d in order]
    print(key, len(names), "launches;", "DUPLICATES (window is not one step)" if len(set(names)) != len(names) and key == "C2" else "", sorted(collections.Counter(names).items()))
json.dump(traffic, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
print({k: v for k, v in traffic.items() if k in ("build", "C2", "C3")})
