"""GPU, >= 2 devices: the multi-GPU paths, each run as a self-spawned one-process-per-GPU job (torchrun) so that
`pytest -m gpu` on a multi-GPU box covers them (skipped on a 1-GPU box; the gloo world-2 tests cover the host
logic there):
  * tests/dist_fused_check.py: the peer-memory gradient exchange + Adam kernel equals NCCL all-reduce +
    torch.optim.Adam, replicas bit-identical, 300 skewed rounds;
  * bench.py --check-only: replicas bit-identical after real steps and the first step's loss equal to the
    single-GPU loss of the global batch (data-parallel == large batch).
"""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(n, script_args, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port())] + script_args
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)


def _world():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    return min(n, 8)


def test_fused_exchange_and_adam_across_gpus():
    n = _world()
    r = _torchrun(n, [os.path.join(ROOT, "tests", "dist_fused_check.py")])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "dist_fused_check ok" in r.stdout


@pytest.mark.parametrize("scaling", ["weak", "strong"])
def test_bench_data_parallel_self_check(scaling):
    n = _world()
    r = _torchrun(n, [os.path.join(ROOT, "bench.py"), "--gpus", str(n), "--check-only", "--scaling", scaling,
                      "--steps", "3", "--warmup", "3", "--no-cpu-baseline"])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "dp self-check ok" in r.stdout + r.stderr
