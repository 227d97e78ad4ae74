"""Real multi-GPU parity (SURVEY section 4 tier T4): spawns one process per GPU with torchrun and runs
tests/mp_gpu_worker.py - the cross-process transports (CUDA-IPC peer memory, NCCL all-gather) and the
`ShardedSwarmsDB` front-end with its default exchange, each compared with single-queue ground truth.
Skips on boxes with fewer than two GPUs (run it with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`)."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _gpus() -> int:
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_transports_and_frontend_match_single_queue(world):
    n = _gpus()
    if n < world:
        pytest.skip(f"needs {world} GPUs, this box has {n}")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(ROOT / "tests" / "mp_gpu_worker.py")]
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    p = subprocess.run(cmd, cwd=str(ROOT), env=env, capture_output=True, text=True, timeout=600)
    tail = (p.stdout[-3000:] + "\n--- stderr ---\n" + p.stderr[-3000:])
    assert p.returncode == 0 and f"MP_GPU_OK world={world}" in p.stdout, tail
