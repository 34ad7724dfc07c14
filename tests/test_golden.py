"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py): the oracle and the
kernels' logic under emulation must reproduce them bit for bit on CPU; the CUDA path on a B200 (integer
resolutions bit-exact, EM within the north-star tolerance 1e-5 relative)."""
import pytest

import emu_lib
import golden_lib
import oracle_lib
from alevin_fry_b200 import Quantifier

CASES = golden_lib.cases()


def test_golden_vectors_exist():
    assert len(CASES) >= 20


@pytest.mark.parametrize("name,res", CASES)
def test_oracle_reproduces_golden(name, res):
    opts, t2g, batch, want = golden_lib.load(name, res)
    golden_lib.assert_matches(oracle_lib.oracle_quant(opts, t2g, batch, n_threads=2), want, True, f"{name}/{res}")


@pytest.mark.parametrize("name,res", [c for c in CASES if c[0] in ("c2_mini", "dense_umi_components")])
def test_emulated_kernels_reproduce_golden(name, res):
    opts, t2g, batch, want = golden_lib.load(name, res)
    golden_lib.assert_matches(emu_lib.emu_quant(opts, t2g, batch), want, True, f"{name}/{res}")


@pytest.mark.gpu
@pytest.mark.parametrize("name,res", CASES)
def test_cuda_reproduces_golden(name, res):
    opts, t2g, batch, want = golden_lib.load(name, res)
    with Quantifier(opts, t2g) as q:
        got = q.quantify_batch(batch)
        assert q.launch_count > 0
    golden_lib.assert_matches(got, want, not res.endswith("-em"), f"{name}/{res}")
