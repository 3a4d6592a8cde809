"""One process per GPU over NCCL (needs >= 2 GPUs; skipped on a single-GPU box): the slab run must
equal the single-handle run bit for bit.  The single-GPU equivalent is tests/test_gpu_slabs.py."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("transport", ["nccl", "p2p", "rods"])
def test_nccl_slabs_equal_single_handle(crb_lib, transport):
    """transport nccl: six-phase protocol, buffers over ncclSend/ncclRecv; p2p: crb_slab_step, the export kernels
    store into the neighbours' arenas (CUDA IPC over NVLink), device-side flags, no host synchronisation; rods:
    bacterial rods with divisions through the six-phase protocol over NCCL"""
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = min(n, 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", {"nccl": "29533", "p2p": "29534", "rods": "29535"}[transport], str(ROOT / "tools" / "check_nccl_slabs.py"),
           transport]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "NCCL_SLABS_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
