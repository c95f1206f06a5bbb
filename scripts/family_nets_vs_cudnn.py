"""Forward + backward of the four SRGAN / CycleGAN networks at their BASELINE shapes: the native engine (bf16 tcgen05)
vs the SAME module trees evaluated by PyTorch operators on this GPU (cuDNN, fp32 and TF32) -- the honest comparison for
configs 3 / 4 (eager on both sides; the step-level numbers of bench.py replay CUDA graphs)."""
import os, sys, copy, torch
import torch.nn as nn
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import ipr_gan_b200
ipr_gan_b200.enable_dropin()
import networks
dev = torch.device("cuda", 0)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for name, shape in (("SRResNet", (16, 3, 24, 24)), ("Discriminator96", (16, 3, 96, 96)),
                    ("Resnet9Blocks", (1, 3, 128, 128)), ("ConvDiscriminator", (1, 3, 128, 128))):
    torch.manual_seed(0)
    net = getattr(networks, name)().to(dev)
    ref = copy.deepcopy(net)
    x = torch.rand(*shape, device=dev)

    def native():
        xx = x.clone().requires_grad_(True)
        y = net(xx)
        y.backward(torch.ones_like(y))

    def torch_ops():
        xx = x.clone().requires_grad_(True)
        y = nn.Sequential.forward(ref, xx)
        y.backward(torch.ones_like(y))

    t_native = timed(native)
    res = {}
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        res[tf32] = timed(torch_ops)
    print("%-18s %-16s native %.2f ms | PyTorch ops fp32 %.2f ms, TF32 %.2f ms" % (name, shape, t_native, res[False], res[True]))
