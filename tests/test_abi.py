"""CPU-side checks of the C-ABI boundary: the CUDA library loads, exports every symbol
include/afq.h declares, and fails loudly (no fallback) when no GPU is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from alevin_fry_b200 import _abi, QuantOpts, Quantifier, AfqError, CellBatch

ROOT = _abi.REPO_ROOT


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "afq.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(afq_[a-z_0-9]+)\s*\(", hdr)))


def test_header_symbols_all_exported_and_bound():
    names = declared_symbols()
    assert "afq_create" in names and "afq_quant_device" in names
    l = _abi.lib()
    for n in names:
        assert hasattr(l, n), f"libafq.so does not export {n}"
        assert n in _abi.SYMBOLS, f"{n} missing from the ctypes binding table"
    assert l.afq_abi_version() == 3


def test_struct_sizes_match_header_layout():
    assert C.sizeof(_abi.AfqConfig) == 56
    assert C.sizeof(_abi.AfqBatch) == 88
    assert C.sizeof(_abi.AfqResult) == 80
    assert C.sizeof(_abi.AfqDeviceOut) == 80


def test_quantopts_validation():
    with pytest.raises(ValueError):
        QuantOpts(resolution="full").to_c()
    assert QuantOpts(resolution="CR-LIKE").to_c().resolution == _abi.RES_CR_LIKE  # case-insensitive
    with pytest.raises(ValueError):
        QuantOpts(sa_model="nope").to_c()


def test_batch_validation_and_slicing():
    b = CellBatch.from_cells([[(1, [0, 1])], [], [(2, [3]), (3, [4, 5, 6])]])
    assert (b.n_cells, b.n_records, b.n_refs_total) == (3, 3, 6)
    s = b.slice_cells(2, 3)
    assert s.n_records == 2 and s.refs.tolist() == [3, 4, 5, 6] and s.rec_ref_offsets.tolist() == [0, 1, 4]
    with pytest.raises(ValueError):
        CellBatch(np.array([0, 2]), np.array([1]), np.array([0, 1]), np.array([0]))


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(AfqError) as ei:
        Quantifier(QuantOpts(num_gene_ids=4, num_rows=4), np.arange(4, dtype=np.uint32))
    assert ei.value.code == _abi.AFQ_ERR_NO_DEVICE
    assert "no CPU fallback" in str(ei.value)


def test_product_never_references_the_oracle():
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "alevin_fry_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"oracle_lib|libafq_oracle|afq_oracle_|oracle/", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, f"product files reference the oracle: {bad}"


def test_host_header_symbols_exported_and_struct_layouts_match(tmp_path):
    # include/afq_host.h: every afqh_* function is exported by libafq_host.so, and the ctypes mirrors in alevin_fry_b200/host.py
    # have the sizes the C compiler gives the header's structs
    import subprocess
    from alevin_fry_b200 import host
    hdr = open(os.path.join(ROOT, "include", "afq_host.h")).read()
    hdr_nc = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(afqh_[a-z_0-9]+)\s*\(", hdr_nc)))
    assert "afqh_quantify" in names and "afqh_infer" in names and "afqh_host_stage_bench" in names
    l = host.lib()
    for n in names:
        assert hasattr(l, n), f"libafq_host.so does not export {n}"
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "%s"\nint main(void) { printf("%%zu %%zu %%zu %%zu", sizeof(afqh_quant_opts), sizeof(afqh_infer_opts), '
                   'sizeof(afqh_rad_info), sizeof(afqh_stage_info)); return 0; }\n' % os.path.join(ROOT, "include", "afq_host.h"))
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-o", str(exe), str(src)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(host._Opts), C.sizeof(host._InferOpts), C.sizeof(host.RadInfo), C.sizeof(host.StageInfo)]
