"""Real peers: the fused product + all-gather over NVLink (peer stores through cudaIpc mappings, NVLS multicast stores,
SpMV) must leave on EVERY rank exactly the bytes of `product into the local block` + NCCL all-gather.  One process per
GPU under torch.distributed.run (tools/check_multi_gpu.py is the per-rank program); needs >= 2 GPUs, skipped otherwise
(the 1-GPU proxy of the same code path is tests/test_peer_bcast.py)."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _device_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_fused_allgather_bit_identical_to_nccl(world):
    if _device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "check_multi_gpu.py")]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert res.returncode == 0 and lines, res.stdout[-2000:] + res.stderr[-4000:]
    report = json.loads(lines[-1])
    assert report["world"] == world and report["all_ranks_ok"], report
    checked = report["rank0"]
    for key in ("f32_n64_rows", "f32_n64_cols", "f64_n24_rows", "f64_n24_cols", "f32_n64_rows_push", "f32_n64_cols_push",
                "f64_n24_rows_push", "f64_n24_cols_push", "f32_n64_cols_pipelined", "f64_n24_cols_pipelined", "spmv"):
        assert checked.get(key) is True, report
    if not checked.get("mcast_unavailable"):
        assert checked.get("f32_n64_rows_mcast") is True and checked.get("f64_n24_rows_mcast") is True, report
