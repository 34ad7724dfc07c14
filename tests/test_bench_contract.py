"""bench.py's JSON-line contract, as far as it can be checked without a GPU: the reference arm (which runs the CPU port and
needs no device) at a tiny size, the `config` object both arms share, and the build hash that ties profiles/ncu_traffic.json
to the CUDA sources."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "exactly ONE JSON line on stdout: " + p.stdout[:500]
    return json.loads(lines[0])


_line = {}


def reference_line():
    if not _line:
        _line.update(run_bench("--impl", "reference", "--cells", "200", "--steps", "2", "--warmup", "1"))
    return _line


def test_reference_arm_line_has_the_contract_keys():
    j = reference_line()
    assert j["impl"] == "reference" and j["unit"] == "cells/s" and j["higher_is_better"] is True and j["scaling"] == "weak"
    assert j["n_gpus"] == 1 and j["steps"] == 2 and j["vs_baseline"] is None and j["data"] == "synthetic" and j["dtype"] == "u32"
    assert j["value"] > 0 and j["ms_per_step"] > 0
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "sample" in cb
    assert cb["one_worker"]["cores"] == 1 and cb["one_worker"]["value"] > 0        # 1 worker beside all cores
    e = j["e2e"]
    assert e["value"] == j["value"] and e["unit"] == j["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    # (with --cells the other configurations are not measured; the default run adds other_configs C3 / C4 / C5)
    assert "other_configs" not in j


def test_both_arms_build_the_same_config_object():
    sys.path.insert(0, ROOT)
    import bench
    import synth
    for name in ("C2", "C3", "C4", "C5"):
        spec = synth.config_spec(name)
        res = bench.CONFIGS[name][1]
        a = bench.workload_config(synth, name, spec, 500, res, 1)
        b = bench.workload_config(synth, name, spec, 500, res, 1)
        assert a == b and "model" not in a and a["workload"].startswith(name) and a["resolution"] == res
        assert a["cells_per_gpu"] == 500 and a["records_per_gpu"] > 0 and a["refs_per_gpu"] >= a["records_per_gpu"]
    j = reference_line()
    spec = synth.config_spec("C2")
    assert j["config"] == bench.workload_config(synth, "C2", spec, 200, "cr-like", 1)


def test_traffic_file_is_tied_to_a_build_hash():
    sys.path.insert(0, ROOT)
    import bench
    h = bench.build_hash()
    assert re.fullmatch(r"[0-9a-f]{16}", h)
    tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    assert re.fullmatch(r"[0-9a-f]{16}", tj["build"]) and tj["C2"] > 0 and tj["C3"] > 0
    assert tj["warp_instructions"]["C2"] > 0 and tj["cells_per_step"]["C2"] == 100000
    # (bench.py reports roofline.traffic / roofline.issue only when tj["build"] == build_hash(); a capture of another build
    # is reported as null with a note, never silently)
