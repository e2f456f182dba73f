"""Multi-GPU: the in-library NCCL path (myrio_comm_init, myrio_score_topx_sharded, myrio_cluster_sharded
with the device-side exchange) on 2 ranks, one per GPU, against the oracle.  Needs >= 2 visible GPUs
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_nccl.py`); skipped on a single-GPU box, where the same
protocol is covered by the emulated-shard tests (test_gpu_parity.py, test_gpu_configs.py) and the gloo tests."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_nccl_two_ranks():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip(f"{n} GPU visible: the NCCL path needs 2")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "NCCL_WORKER_OK" in r.stdout
