"""The sharded product path on real GPUs: tools/sharded_check.py under torchrun on 2 GPUs (skipped on a 1-GPU box; the
gloo world-size-2 test of tests/test_sharding_gloo.py covers the host logic everywhere)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_equals_unsharded_on_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "sharded_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SHARDED_CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
