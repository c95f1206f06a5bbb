"""Per-block comparison of the native engine with the bf16-matched oracle (diagnostic; run on the GPU box)."""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import ipr_gan_b200; ipr_gan_b200.enable_dropin()
import networks
from ipr_gan_b200 import seqnet
from oracle import seq_oracle as so
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_seqnet as T

def rel(a, b):
    a, b = a.detach().cpu().float(), b.detach().cpu().float()
    return float((a - b).norm() / (b.norm() + 1e-12))

for name, shape in (("SRResNet", (3, 3, 24, 24)), ("Discriminator96", (16, 3, 96, 96)), ("Resnet9Blocks", (1, 3, 64, 64))):
    torch.manual_seed(2)
    net = getattr(networks, name)(); T._randomise(net)
    ref = copy.deepcopy(net)
    x = torch.rand(*shape) * 2 - 1
    seqnet.CAPTURE = []
    out = net.cuda()(x.cuda())
    cap, seqnet.CAPTURE = seqnet.CAPTURE, None
    simcap = []
    bl = seqnet.lower(ref)
    sim = so.forward_sim_bf16(ref, x, bl, capture=simcap)
    print(name, "final", rel(out, sim))
    for i, ((y0, z), (sy, sz)) in enumerate(zip(cap, simcap)):
        b = bl[i]
        print("  block %2d k%d s%d %s act%d res%s  conv %.2e  out %.2e" % (i, b.k, b.stride, type(b.norm).__name__[:5], b.act, b.residual,
              rel(y0[..., :b.cout].permute(0, 3, 1, 2), sy), rel(z.permute(0, 3, 1, 2)[:, :sz.shape[1]], sz)))
