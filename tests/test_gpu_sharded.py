"""Row-sharded multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): runs scripts/sharded_check.py
under torchrun and checks every reported error against the 1e-10 bar."""
import os
import re
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_row_sharded_expv_phiv_kiops_two_gpus():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    if torch.cuda.device_count() < 2:
        pytest.skip("row sharding needs at least 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "scripts", "sharded_check.py"), "parity", "dense", "reorth"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    errs = [float(x) for x in re.findall(r"relerr ([0-9.e+-]+)", res.stdout)]
    assert len(errs) >= 17 and max(errs) < 1e-10, res.stdout
    for a, b in re.findall(r"stats (\([^)]*\)) vs (\([^)]*\))", res.stdout):
        assert a == b
