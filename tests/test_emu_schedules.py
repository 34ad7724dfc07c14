"""The emulated kernels must give the oracle's answer under other fiber schedules too: reverse round robin,
and pseudo-random schedules that also pre-empt a thread right before an atomic (so the lost-race paths of the
open-address tables, lists and the union-find are executed). tests/emu/cuda_emu.h reads AFQ_EMU_SCHED once per
process, hence the subprocess."""
import os
import subprocess
import sys

import pytest

import emu_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("sched", ["1", "5"])
def test_emulated_kernels_under_other_schedules(sched):
    emu_lib.build()
    env = dict(os.environ, AFQ_EMU_SCHED=sched)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "emu_asan_cases.py"), "quick"], env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) >= 9 and all(l.endswith("OK") for l in lines), r.stdout
