"""GPU, >= 2 devices: the data-parallel protected step over NCCL (one process per GPU) against the oracle's model of
the reference's nn.DataParallel (experiments/base.py:24-39): torch.chunk shards, per-replica BatchNorm, summed
gradients, losses over the whole batch.  Skipped on a single-GPU box (run with `gpurun --gpus 2`)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["eager", "graph"])
def test_two_rank_step_vs_sharded_oracle(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29500 + (os.getpid() % 400)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "nccl_step_worker.py"), mode, "32"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
    assert r.returncode == 0 and lines, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.loads(lines[-1][7:])
    assert res["ok"] and res["replicas_identical"], res
    assert res["worst_metric_rel"] <= 2e-2, res                       # bf16 budget on all seven metrics, both steps
    assert res["rank_spread"] == 0.0, res                             # get_metrics() returns the global value on every rank
    assert res["param_max_abs"] <= 2.0e-3, res                        # two Adam steps of lr 2e-4 on either side
