"""BASELINE config 5: verification sweep over 10 000 trigger samples (SSIM + pHash p-value + sign-bit BER),
on-device, vs the CPU oracle on a bounded sample.  Prints one JSON line."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ipr_gan_b200  # noqa: E402

ipr_gan_b200.enable_dropin()
import models  # noqa: E402
from configs import presets  # noqa: E402
from ipr_gan_b200 import verify  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:                                   # python -m torch.distributed.run --nproc-per-node N scripts/verify_bench.py
    torch.distributed.init_process_group("nccl", device_id=dev)
torch.manual_seed(1234)
model = models.DCGAN(presets.dcgan_model(), device=[dev])
model = models.BlackBoxWrapper(model, presets.dcgan_blackbox())
model = models.WhiteBoxWrapper(model, presets.dcgan_whitebox())
G = model.G
verify.verification_sweep(G, model.fn_inp, model.fn_out, 1000, batch=500, sign_model=model.loss_model)
torch.cuda.synchronize()
if world > 1:
    torch.distributed.barrier()
t0 = time.perf_counter()
res = verify.verification_sweep(G, model.fn_inp, model.fn_out, N, batch=500, sign_model=model.loss_model)
torch.cuda.synchronize()
if world > 1:
    torch.distributed.barrier()
dt = time.perf_counter() - t0
if rank != 0:
    os._exit(0)
# CPU oracle on a bounded sample of the same verification arithmetic (crops -> SSIM + pHash p-value)
from oracle import ipr_oracle as orc  # noqa: E402
wx = torch.rand(500, 3, 16, 16)
wy = (wx + 0.05 * torch.randn_like(wx)).clamp(0, 1)
t1 = time.perf_counter()
orc.ssim_per_sample(wx, wy)
orc.matching_prob(wx, wy)
cpu_dt = time.perf_counter() - t1
print(json.dumps({"n_gpus": world, "workload": "verification sweep, %d trigger samples (G fwd x2 + paste + crop + SSIM + pHash p + BER)" % N,
                  "samples_per_s": N / dt, "seconds": dt, "result": {k: v for k, v in res.items() if k != "per_sample"},
                  "cpu_oracle_verification_only_samples_per_s": 500 / cpu_dt, "cpu_threads": torch.get_num_threads()}))
if world > 1:
    sys.stdout.flush()
    os._exit(0)
