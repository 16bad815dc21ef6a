"""The product's multi-GPU path over NCCL: ShardedOptimizer(GeneralizedICP()) on 2 GPUs of one box,
one process per GPU under torchrun.  Skipped below 2 GPUs (the CPU suite covers the host logic with gloo,
tests/test_sharding_gloo.py)."""
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


def test_sharded_gicp_over_nccl_matches_single_gpu_and_the_oracle_run():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tools", "multigpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-3000:])
    lines = [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == world
    for r in lines:
        assert r["raw_batch_identical"] and r["align_matches_oracle_run"] and r["ranks_agree"], r
        assert r["local_units"] in (17, 18) and r["local_target_sets"] <= 3       # 35 units, 5 target sets, 2 ranks
