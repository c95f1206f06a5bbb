"""Where the verification sweep's time goes (CUDA events around each stage of one 2500-sample batch)."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import ipr_gan_b200
ipr_gan_b200.enable_dropin()
import models
from configs import presets
from ipr_gan_b200 import ops, verify
dev = torch.device("cuda", 0)
torch.manual_seed(1234)
model = models.DCGAN(presets.dcgan_model(), device=[dev])
model = models.BlackBoxWrapper(model, presets.dcgan_blackbox())
model = models.WhiteBoxWrapper(model, presets.dcgan_whitebox())
G, fn_inp, fn_out = model.G, model.fn_inp, model.fn_out
G.eval()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2500
gen = torch.Generator().manual_seed(1)
crop_bg = torch.zeros(1, 1, 16, 16, device=dev)
names, evs = [], []
def mark(n):
    e = torch.cuda.Event(enable_timing=True); e.record(); names.append(n); evs.append(e)
with torch.no_grad():
    for rep in range(3):
        names, evs = [], []
        t0 = time.perf_counter()
        mark("start")
        z = torch.randn(B, 128, generator=gen).to(dev, non_blocking=True); mark("z host randn + H2D")
        x = G(z); mark("G(z)")
        xwm = G(fn_inp(z)); mark("fn_inp + G(zwm)")
        ywm = fn_out(x); mark("fn_out paste")
        wm_x = ops.crop_patch(xwm, crop_bg, "tl", 16, postproc=True)
        wm_y = ops.crop_patch(ywm, crop_bg, "tl", 16, postproc=True); mark("crop + postproc x2 (fused)")
        q = ops.ssim_per_sample(wm_x, wm_y); mark("ssim per sample")
        p, r = ops.matching_prob(wm_x, wm_y); mark("pHash p-value (bicubic + hash x2 + p)")
        s = torch.stack([q.double().sum(), p.double().sum(), (p < 0.01).double().sum()]); mark("sums")
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
print("batch %d: wall %.3f ms" % (B, wall * 1e3))
for i in range(1, len(evs)):
    print("  %-40s %8.1f us" % (names[i], evs[i - 1].elapsed_time(evs[i]) * 1e3))
