#!/usr/bin/env python
"""How much of the fused all-gather step is the SM-side cost of the extra stores?  The bcast product with all
destinations LOCAL (no NVLink): 1, 2, 4, 8 result buffers on one GPU (cfg3, fp32 k=64, row-major)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import WORKLOADS, _time_ms  # noqa: E402
from matrixextra_b200._lib import MXG_F32, MXG_KEEP_F32, MXG_ROWS_CONTIGUOUS  # noqa: E402
from matrixextra_b200.device import DeviceCSR  # noqa: E402

torch.cuda.set_device(0)
wl = WORKLOADS["cfg3"]
m, K, n = wl["m"], wl["K"], 64
A = DeviceCSR.synth(m, K, wl["nnz"], 1, 1, seed=1003, keep=MXG_KEEP_F32)
B = torch.randn(K, n, device="cuda")
outs = [torch.empty(m, n, device="cuda") for _ in range(8)]
for nd in (1, 2, 4, 8):
    ptrs = [o.data_ptr() for o in outs[:nd]]
    ms = _time_ms(lambda: A.spmm_bcast(B, ptrs, n, MXG_F32, MXG_ROWS_CONTIGUOUS), 10, 3)
    print(json.dumps(dict(case="bcast_all_local", destinations=nd, ms=ms, extra_store_GB=(nd - 1) * m * n * 4 / 1e9)), flush=True)
