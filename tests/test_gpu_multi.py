"""-m gpu, needs >= 2 GPUs (skipped otherwise): the data-parallel sweep through the product code -- NCCL all-reduce of the
per-bond gradient inside trmps_bond_step, replicated split -- on every GPU of the box.  tools/dp_check.py asserts, after one
full training step, that the weights are BIT-IDENTICAL on all ranks (one ULP of divergence in the all-reduced gradient would
desynchronise the replicated splits) and equal, to FP32 round-off, to a single-rank run on the concatenated batch."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_data_parallel_step_is_bit_identical_on_all_ranks():
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "dp_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(res.stdout[-2000:])
    assert res.returncode == 0, res.stderr[-3000:]
    assert "ranks bit-identical: True" in res.stdout
    assert ("world=%d" % n) in res.stdout
