"""CPU tests of the boundary: the C-ABI library loads without a GPU, exports every symbol that
include/pngloss_b200.h declares, and refuses to compute without a device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import pngloss_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pngloss_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(pngloss_b200_\w+|optimize_with_rows|optimize_with_stride|optimize_image|"
                       r"optimizeForAverageFilter)\s*\(", text)
    return sorted(set(names))


@pytest.fixture(scope="module", autouse=True)
def _built():
    if not os.path.exists(pngloss_b200.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(pngloss_b200.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(pngloss_b200.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_library_has_no_oracle_or_torch_dependency():
    import subprocess
    out = subprocess.run(["ldd", pngloss_b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "torch" not in out and "pngloss_ref" not in out


def test_no_cpu_fallback_without_device():
    """On a GPU-less host the product must fail loudly instead of computing on the CPU."""
    if pngloss_b200.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(pngloss_b200.PnglossError):
        pngloss_b200.Context(0)
    img = np.full((2, 2, 4), 7, np.uint8)
    rf = np.zeros(2, np.uint8)
    rc = pngloss_b200.optimize_with_rows(img, rf, False, 20, 2)
    assert rc == pngloss_b200.DEVICE_ERROR
    assert (img == 7).all() and not rf.any()


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the CPU arm the driver runs beside the GPU arm) on the smallest config: one JSON
    line with the keys the contract names, on this GPU-less host."""
    import json
    import subprocess
    import sys
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        pytest.skip("oracle not built")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "1",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["cores"] >= 1
    assert line["config"]["baseline_config"] == "1" and line["config"]["strength"] == 19
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
