"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/mxgpu.h
declares, fails loudly without a device (no CPU fallback), and its host-only helpers work."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "mxgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mxg_[A-Za-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from matrixextra_b200 import _lib
    lib = _lib.load()
    declared = _header_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mxgpu.h but not exported"
    assert sorted(_lib.exported_names()) == declared, "ctypes table and header disagree"


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from matrixextra_b200 import _lib, rcpp_exports as rx
    with pytest.raises(_lib.MxgError) as ei:
        rx.matmul_csr_dvec_numeric([0, 1], [0], [1.0], [2.0], 1)
    assert ei.value.code == _lib.MXG_ERR_CUDA
    assert _lib.device_count() == 0


def test_product_never_imports_the_oracle():
    code = ("import sys; import matrixextra_b200, matrixextra_b200.rcpp_exports, matrixextra_b200.device, "
            "matrixextra_b200.sharded; print(any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules))")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, check=True)
    assert out.stdout.strip() == "False"
    pkg = os.path.join(ROOT, "matrixextra_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "cpu_oracle" not in src and "libmxoracle" not in src and "libmxref" not in src, f


def test_options_roundtrip_and_errors():
    from matrixextra_b200 import _lib
    old = _lib.get_option("piece")
    _lib.set_option("piece", 777)
    assert _lib.get_option("piece") == 777
    _lib.set_option("piece", old)
    with pytest.raises(_lib.MxgError) as ei:
        _lib.set_option("no_such_option", 1)
    assert ei.value.code == _lib.MXG_ERR_ARG and "no_such_option" in ei.value.message


def test_row_partition_balances_entries():
    from matrixextra_b200.device import row_partition
    rng = np.random.default_rng(0)
    lens = np.floor(5 * (1 - rng.random(10000)) ** (-1 / 1.5)).astype(np.int64)  # power law
    p = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    for parts in (1, 2, 3, 4, 8):
        b = row_partition(p, parts)
        assert b[0] == 0 and b[-1] == 10000 and (np.diff(b) >= 0).all()
        per = np.diff(p[b])
        assert per.sum() == p[-1]
        assert per.max() <= p[-1] / parts + lens.max()  # within one row of perfect balance
    # degenerate inputs: empty matrix, more parts than rows, all entries in one row
    assert row_partition(np.zeros(1, np.int32), 4).tolist() == [0, 0, 0, 0, 0]
    b = row_partition(np.array([0, 0, 0, 0], np.int32), 2)
    assert b[0] == 0 and b[-1] == 3
    b = row_partition(np.array([0, 100, 100, 100], np.int32), 3)
    assert b[0] == 0 and b[-1] == 3 and (np.diff(b) >= 0).all()


def test_header_is_plain_c_and_cxx():
    """include/mxgpu.h is the drop-in boundary: it must compile as C99 (cgo / .Call-style consumers) and as C++11 (the
    Rcpp glue) with warnings as errors, with nothing but <stddef.h> / <stdint.h> behind it."""
    import subprocess
    hdr = os.path.join(ROOT, "include", "mxgpu.h")
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr],
                   check=True, env=env)
    subprocess.run(["/usr/bin/g++", "-std=c++11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c++", hdr],
                   check=True, env=env)
    text = open(hdr).read()
    assert "torch" not in text.replace("torch symmetric memory", "").replace("torch.distributed._symmetric_memory", "")
