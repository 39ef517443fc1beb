"""NCCL path (needs >= 2 GPUs on the box; skipped on the single-GPU tier): one process per GPU,
decomposed blocks with halo exchange vs. the single-GPU run, bit for bit (tools/check_dist.py)."""
import os
import subprocess
import sys

import pytest

from tests.util import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_decomposition_bit_identical(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), os.path.join(ROOT, "tools", "check_dist.py")]
    p = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900, text=True)
    assert p.returncode == 0 and "DIST_CHECK PASS" in p.stdout, p.stdout[-3000:]
