"""One eager protected IPR-DCGAN step between cudaProfilerStart/Stop, for `ncu --profile-from-start off`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ipr_gan_b200.trainer import ProtectedDCGANTrainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda", 0)
tr = ProtectedDCGANTrainer(B, dev, use_graph=False)
g = torch.Generator().manual_seed(1234)
tr.set_inputs(torch.randn(B, 3, 32, 32, generator=g).clamp(-1, 1), torch.randn(B, 128, generator=g))
for _ in range(3):
    tr._step()
torch.cuda.synchronize()
# algorithmic operand / result bytes of every tensor-core GEMM launch of the step (for roofline.traffic)
import json
from ipr_gan_b200 import dense
algo = {"bytes": 0.0}
_run, _wrun = dense.Plan.run, dense.WGradPlan.run


def run(self, a, b_packed, *args, **kw):
    out, st = _run(self, a, b_packed, *args, **kw)
    algo["bytes"] += a.numel() * 2 + b_packed.numel() * 2 + out.numel() * out.element_size()
    return out, st


def wrun(self, y, x, grad, *args, **kw):
    algo["bytes"] += y.numel() * 2 + x.numel() * 2 + grad.numel() * 4
    return _wrun(self, y, x, grad, *args, **kw)


dense.Plan.run, dense.WGradPlan.run = run, wrun
torch.cuda.profiler.start()
tr._step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"batch": B, "gemm_algorithmic_bytes": algo["bytes"]}, open(os.path.join(ROOT, "gpurun_out", "r2_step_algorithmic.json"), "w"))
print("done")
