"""Development probe: per-call wall time of the warm (handle) product when consecutive calls differ in shape or in the
host_colsplit option — does the device memory pool / page-locked arena make a call pay for its predecessor's buffers?"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
from matrixextra_b200 import _lib, rcpp_exports as rx  # noqa: E402
from matrixextra_b200._lib import MXG_KEEP_F32  # noqa: E402
from matrixextra_b200.device import DeviceCSR  # noqa: E402

wl = bench.WORKLOADS["cfg3"]
A = DeviceCSR.synth(wl["m"], wl["K"], wl["nnz"], wl["row_model"], wl["col_model"], seed=wl["seed"], keep=MXG_KEEP_F32 | 2)
p, j, x = A.to_host()
A.free()
K, m = wl["K"], wl["m"]
h = rx.as_gpu_csr(p, j, x, K, keep_float64=False, keep_float32=True)
rng = np.random.default_rng(1)
X = {n: np.asfortranarray(rng.standard_normal((n, K)).astype(np.float32)) for n in (64, 32, 48)}


def pinned(a):
    import torch
    t = torch.empty(a.size, dtype=torch.float32).pin_memory().numpy().reshape(a.shape, order="F")
    t[...] = a
    return t


XP = {64: pinned(X[64])}
OUT = {64: pinned(np.zeros((64, m), dtype=np.float32, order="F"))}


def run(tag, n, split, reps=4, pin=False):
    _lib.set_option("host_colsplit", split)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = rx.gpu_csr_dense_tcrossprod_float32(XP[n] if pin else X[n], h, out=OUT[n] if pin else None)
        ts.append(round((time.perf_counter() - t0) * 1e3, 2))
        del r
    print(json.dumps({"phase": tag, "n": n, "host_colsplit": split, "page_locked": pin, "ms_per_call": ts}), flush=True)


run("first calls", 64, 1)
run("same shape, one piece", 64, 0)
run("back to halves", 64, 1)
run("one piece again", 64, 0)
run("narrower", 32, 0)
run("wider again", 64, 0)
run("n = 48", 48, 0)
run("n = 64", 64, 0)
for rep in range(3):
    run("page-locked operand and result", 64, 1, 6, pin=True)
    run("page-locked operand and result", 64, 0, 6, pin=True)
rx.gpu_csr_free(h)
