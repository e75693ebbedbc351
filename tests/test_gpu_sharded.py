"""GPU (>= 2 devices): the sharded filter equals the single-GPU filter."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["placed", "p2p", "pull", "nccl"])
def test_sharded_filter_equals_single_gpu(mode):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = min(n, 4)
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(here, "sharded_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, FS2_DIST=mode))
    assert r.returncode == 0 and "SHARDED_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    assert "migrated=0 " not in r.stdout          # the run really moved particles between GPUs
