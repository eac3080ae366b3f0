"""N > 1 on real GPUs: NCCL all-gather of keyframe descriptor blocks + cross-shard matching (BASELINE config 5), two
ranks under torch.distributed.run.  Skipped on a box with one GPU (the gloo twin of the host logic is
tests/test_abi_and_host.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_nccl_cross_shard_matching_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29631", os.path.join(ROOT, "tests", "multi_gpu_worker.py")], env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("identical to the oracle") == 2, r.stdout[-2000:]
