"""N>1 on real GPUs: the batch sharded over the ranks of one box; the gathered observation tensor -- NCCL all-gather
and every variant of the fused NVLink push (CUDA IPC / symmetric memory + NVLS multicast, sync / pipelined, with rank
skew, lag-1 consumers and CUDA-graph replays) -- equals the single-GPU result, on 2 ranks and on every GPU of the box.
Skipped on a one-GPU box (run with `gpurun --gpus 2` / `--gpus 8`; logs of both under profiles/)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_worker(world):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=int(os.environ.get('CARLB_MGPU_TIMEOUT', '500')), cwd=ROOT)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert p.stdout.count("MGPU_OK") == world


def test_two_rank_gather_equals_single_gpu():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run_worker(2)


def test_all_rank_gather_equals_single_gpu():
    import torch

    n = torch.cuda.device_count()
    if n <= 2:
        pytest.skip("needs more than 2 GPUs (the 2-rank case is covered above)")
    _run_worker(min(n, 8))
