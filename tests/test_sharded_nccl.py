"""The one-process-per-GPU path (pynbodyext.gravity.sharded: shard upload, ONE NCCL all-gather, target shards) on real
hardware: a 2-rank torchrun job whose ranks each check their shard against the CPU oracle. Needs >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_nccl_job_matches_the_oracle():
    import pynbodyext._rust as r
    if r._load().pnbx_device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "workers", "nccl_sharded_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "NCCL_SHARDED_OK world=2" in out.stdout
