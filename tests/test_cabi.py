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


def test_new_entry_points_fail_loudly_without_a_device():
    """mxg_set_devices, the handle products with host operands and the operand cache have no CPU path either."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from matrixextra_b200 import _lib, rcpp_exports as rx
    with pytest.raises(_lib.MxgError) as ei:
        _lib.call("mxg_set_devices", 2)
    assert ei.value.code == _lib.MXG_ERR_CUDA
    with pytest.raises(_lib.MxgError) as ei:
        rx.as_gpu_csr([0, 1], [0], [1.0], 1)
    assert ei.value.code == _lib.MXG_ERR_CUDA
    _lib.set_option("cache_mb", 64)
    try:
        with pytest.raises(_lib.MxgError) as ei:
            rx.tcrossprod_csr_dense_numeric([0, 1], [0], [1.0], np.ones((2, 1), order="F"))
        assert ei.value.code == _lib.MXG_ERR_CUDA
    finally:
        _lib.set_option("cache_mb", 0)
    # the cache bookkeeping itself needs no device
    hits, misses, nbytes, entries = C.c_ulonglong(), C.c_ulonglong(), C.c_size_t(), C.c_int()
    _lib.call("mxg_cache_stats", C.byref(hits), C.byref(misses), C.byref(nbytes), C.byref(entries))
    assert nbytes.value == 0 and entries.value == 0
    n = C.c_int(0)
    _lib.call("mxg_get_devices", C.byref(n))
    assert n.value == 1
    # no page-locked result memory either: the glue's allocator hook then falls back to an ordinary block
    ptr = C.c_void_p()
    with pytest.raises(_lib.MxgError) as ei:
        _lib.call("mxg_host_alloc", 64 << 20, C.byref(ptr))
    assert ei.value.code == _lib.MXG_ERR_CUDA and not ptr.value
    assert rx._pooled_empty((1 << 20, 8), np.float64) is None


def test_r_side_artefacts_are_files_and_apply():
    """rglue/: the glue sources, the R methods file and the patch for the two reference files that change."""
    for f in ("matmul_gpu_glue.cpp", "rowops_gpu_glue.cpp", "handle_gpu_glue.cpp", "matmul_gpu_methods.R", "mxgpu.patch"):
        assert os.path.getsize(os.path.join(ROOT, "rglue", f)) > 500, f
    r = open(os.path.join(ROOT, "rglue", "matmul_gpu_methods.R")).read()
    for sig in ('setMethod("crossprod", signature(x="RsparseMatrix", y="matrix")', 'signature(x="CsparseMatrix", y="matrix")',
                'signature(x="matrix", y="RsparseMatrix")', 'setClass("gpuRsparse"', "as.gpu.csr <- function", "mxgpu.settings <- function"):
        assert sig in r, sig
    # every glue export the R file calls exists in the glue sources
    glue = "".join(open(os.path.join(ROOT, "rglue", f)).read() for f in ("matmul_gpu_glue.cpp", "handle_gpu_glue.cpp"))
    for fn in ("crossprod_csr_dense_numeric", "crossprod_csr_dense_float32", "as_gpu_csr", "gpu_csr_free", "gpu_csr_tcrossprod_dense_numeric",
               "gpu_csr_tcrossprod_dense_float32", "gpu_csr_dense_tcrossprod_numeric", "gpu_csr_dense_tcrossprod_float32",
               "gpu_csr_crossprod_dense_numeric", "gpu_csr_crossprod_dense_float32", "gpu_csr_dvec_numeric", "mxgpu_configure"):
        assert fn + "(" in r and fn + "(" in glue, fn
    patch = open(os.path.join(ROOT, "rglue", "mxgpu.patch")).read()
    assert "-DMATRIXEXTRA_USE_MXGPU" in patch and "#ifndef MATRIXEXTRA_USE_MXGPU" in patch and "-lmxgpu" in patch
    ref = "/root/reference"
    if os.path.isdir(os.path.join(ref, "src")):  # build container only: the patch applies to the reference tree as it is
        import shutil
        import tempfile
        with tempfile.TemporaryDirectory() as tmp:
            shutil.copytree(os.path.join(ref, "src"), os.path.join(tmp, "src"))
            subprocess.run(["patch", "-p1", "-s", "-i", os.path.join(ROOT, "rglue", "mxgpu.patch")], cwd=tmp, check=True)
            fenced = open(os.path.join(tmp, "src", "matmul.cpp")).read()
            a, b = fenced.index("#ifndef MATRIXEXTRA_USE_MXGPU"), fenced.index("#endif /* MATRIXEXTRA_USE_MXGPU */")
            inside = fenced[a:b]
            for name in ("gemm_csr_drm_as_drm", "gemm_csr_drm_as_dcm", "matmul_dense_csc_numeric", "tcrossprod_csr_dense_float32",
                         "matmul_csr_dvec_float32"):
                assert name in inside, name
            assert "matmul_colvec_by_scolvecascsr" not in inside and "matmul_colvec_by_scolvecascsr" in fenced[b:]
