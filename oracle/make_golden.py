"""oracle/make_golden.py -- TEST INFRASTRUCTURE ONLY.

Generates ``tests/golden/*.npz`` by running the UNMODIFIED reference (``/root/reference`` through
``oracle/ref_bridge.py``) on seeded synthetic inputs, in the build container.  The reference has
no golden vectors of its own; these files are the pinned outputs of the
reference's Python that the oracle restatement and the CUDA path are both checked against.

    python -m oracle.make_golden          # from the repo root

Third-party arithmetic inside these vectors (pytorch-msssim SSIM, pdqhash bits) comes from the
restatements under oracle/shims (parity unpinned for those two).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_bridge  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
MARK = os.path.join(ROOT, "ipr_gan_b200", "assets", "watermark_a.png")
SEED = 1234


def t2n(t):
    return t.detach().cpu().numpy()


def gen_triggers(ref):
    tools, Config = ref["tools"], ref["configs"].Config
    out = {}
    torch.manual_seed(SEED)
    x = torch.randn(4, 3, 32, 32).clamp(-1, 1)
    x[0, 0, 0, 0] = -0.0  # signed-zero edge case
    out["x"] = t2n(x)
    for tag, opaque, norm, pos, size in (("op_tl", True, True, "tl", 16), ("al_br", False, True, "br", 16),
                                         ("al_tr_raw", False, False, "tr", 12), ("op_bl", True, False, "bl", 20)):
        cfg = Config({"size": size, "opaque": opaque, "watermark": MARK, "position": pos, "type": "PasteWatermark"})
        m = tools.PasteWatermark(cfg, normalized=norm)
        out[f"paste_{tag}_fg"] = t2n(m.fg)
        out[f"paste_{tag}_bg"] = t2n(m.bg)
        out[f"paste_{tag}_y"] = t2n(m(x))
        out[f"paste_{tag}_crop"] = t2n(m.apply_mask(x))
    torch.manual_seed(SEED + 1)
    cfg = Config({"size": 12, "position": "br", "type": "RandomNoisePatch"})
    m = tools.RandomNoisePatch(cfg, normalized=False)
    x24 = torch.rand(3, 3, 24, 24)
    out["noise_x"] = t2n(x24)
    out["noise_fg"] = t2n(m.fg)
    out["noise_y"] = t2n(m(x24))
    out["noise_crop"] = t2n(m.apply_mask(x24))
    torch.manual_seed(SEED + 2)
    cfg = Config({"n_bit": 10, "constant": -10.0, "z_dim": 128, "type": "RandomBitMask"})
    m = tools.RandomBitMask(cfg)
    z = torch.randn(6, 128)
    out["z"] = t2n(z)
    out["bitmask_mask"] = t2n(m.mask)
    out["bitmask_y"] = t2n(m(z))
    out["tdist_y"] = t2n(tools.TransformDist(Config({"type": "TransformDist"}))(z))
    torch.manual_seed(SEED + 3)
    m = tools.TransformVar(Config({"type": "TransformVar"}))
    out["tvar_a"], out["tvar_w"], out["tvar_y"] = t2n(m.a), t2n(m.w), t2n(m(z))
    np.savez_compressed(os.path.join(GOLD, "triggers.npz"), **out)


def gen_ssim(ref):
    tools = ref["tools"]
    import pytorch_msssim  # the shim, as the reference sees it
    out = {}
    torch.manual_seed(SEED + 10)
    for tag, shape, norm in (("a", (4, 3, 32, 32), True), ("b", (2, 3, 40, 52), False), ("c", (1, 3, 96, 96), False)):
        x = torch.rand(*shape)
        y = (x + 0.25 * torch.randn(*shape)).clamp(0, 1)
        if norm:
            x, y = x * 2 - 1, y * 2 - 1
        x.requires_grad_(True)
        loss = tools.ssim(normalized=norm)(x, y)
        loss.backward()
        out[f"{tag}_x"], out[f"{tag}_y"] = t2n(x), t2n(y)
        out[f"{tag}_loss"], out[f"{tag}_grad"] = t2n(loss), t2n(x.grad)
        out[f"{tag}_norm"] = np.array(norm)
    wx = torch.rand(8, 3, 16, 16)
    wy = (wx + 0.1 * torch.randn(8, 3, 16, 16)).clamp(0, 1)
    wy[0] = wx[0]
    out["ps_x"], out["ps_y"] = t2n(wx), t2n(wy)
    out["ps_ssim"] = t2n(pytorch_msssim.ssim(wx, wy, data_range=1, size_average=False))
    np.savez_compressed(os.path.join(GOLD, "ssim.npz"), **out)


def gen_sign(ref):
    tools, networks, Config = ref["tools"], ref["networks"], ref["configs"].Config
    out = {}
    torch.manual_seed(SEED + 20)
    G = networks.ConvGenerator32()
    cfg = Config({"gamma_0": 0.1, "string": "EXAMPLE A", "target": "G"})
    sm = tools.SignLossModel(G, cfg)
    names = [n.replace(".", "_") for n, m in G.named_modules() if isinstance(m, torch.nn.BatchNorm2d)]
    out["names"] = np.array(names)
    bns = [m for m in G.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    with torch.no_grad():
        for i, m in enumerate(bns):
            noise = torch.randn_like(m.weight) * 0.2
            m.weight.copy_(noise)           # mixed signs, some |gamma| < gamma_0
            if i == 0:
                m.weight[:3] = 0.0           # gamma == 0 counts as a bit error
    loss = sm(G)
    loss.backward()
    for n, m in zip(names, bns):
        out[f"sign_{n}"] = t2n(getattr(sm, n))
        out[f"gamma_{n}"] = t2n(m.weight)
        out[f"grad_{n}"] = t2n(m.weight.grad)
    out["loss"] = t2n(loss)
    out["ber"] = t2n(sm.compute_ber(G))
    np.savez_compressed(os.path.join(GOLD, "sign.npz"), **out)


def gen_phash(ref):
    tools = ref["tools"]
    out = {}
    torch.manual_seed(SEED + 30)
    base = torch.rand(12, 3, 16, 16)
    smooth = torch.nn.functional.interpolate(torch.rand(12, 3, 4, 4), size=16, mode="bilinear")
    wx = (0.5 * base + 0.5 * smooth).clamp(0, 1)
    wy = (wx + 0.05 * torch.randn_like(wx)).clamp(0, 1)
    wy[0] = wx[0]
    wy[1] = torch.rand(3, 16, 16)
    wx[2] = 1.0          # saturated patch (bicubic of a constant)
    wy[2] = 0.0
    wx[3, :, :8] = 1.0   # hard edge -> bicubic overshoot, uint8 wrap-around
    wx[3, :, 8:] = 0.0
    out["x16"], out["y16"] = t2n(wx), t2n(wy)
    out["p16"] = t2n(tools.compute_matching_prob(wx, wy))
    bx = torch.rand(5, 3, 48, 40)
    by = (bx + 0.1 * torch.randn_like(bx)).clamp(0, 1)
    out["x48"], out["y48"] = t2n(bx), t2n(by)
    out["p48"] = t2n(tools.compute_matching_prob(bx, by))
    from tools import phash_pvalue
    up = torch.nn.functional.interpolate(wx, size=(32, 32), mode="bicubic", align_corners=False)
    out["x16_up"] = t2n(up)
    out["x16_hash"] = phash_pvalue.compute_hash(up).astype(np.uint8)
    out["x48_hash"] = phash_pvalue.compute_hash(bx).astype(np.uint8)
    from scipy.stats import binom
    r = np.arange(257)
    out["ptable"] = np.array([1 - binom(n=256, p=0.5).cdf(k - 1) for k in r]).astype(np.float32)
    np.savez_compressed(os.path.join(GOLD, "phash.npz"), **out)


def gen_dcgan_step(ref):
    """Three full IPR-DCGAN steps (black box + white box) at B=8 through the reference's own
    models.DCGAN -> BlackBoxWrapper -> WhiteBoxWrapper; weights come from the seeded default init so the
    fixture only stores inputs' seed, per-step metrics and small output slices."""
    models, Config = ref["models"], ref["configs"].Config
    torch.manual_seed(SEED)
    mcfg = Config({"G": "ConvGenerator32", "D": "SNDiscriminator32", "opt": "Adam",
                   "opt_param": {"lr": 2.0e-4, "betas": [0.5, 0.999]}, "type": "DCGAN"})
    model = models.DCGAN(mcfg, device=[torch.device("cpu")])
    bbox = Config({"fn_inp": {"type": "TransformDist"},
                   "fn_out": {"size": 16, "opaque": True, "type": "PasteWatermark", "watermark": MARK},
                   "lambda": 1.0, "loss_fn": "ssim", "normalized": True, "input_var": "latent",
                   "output_var": "generated", "target": "G"})
    model = models.BlackBoxWrapper(model, bbox)
    wbox = Config({"gamma_0": 0.1, "string": "EXAMPLE A", "target": "G"})
    model = models.WhiteBoxWrapper(model, wbox)
    out = {"seed": np.array(SEED), "batch": np.array(8)}
    g = torch.Generator().manual_seed(SEED)
    keys = None
    rows = []
    for step in range(3):
        real = torch.randn(8, 3, 32, 32, generator=g).clamp(-1, 1)
        z = torch.randn(8, 128, generator=g)
        model.update_d({"real_sample": real, "latent": z})
        model.update_g({"fake_sample": model.fake_sample})
        met = model.get_metrics()
        keys = sorted(met)
        rows.append([met[k] for k in keys])
        if step == 0:
            out["fake0"] = t2n(model.fake_sample[:2])
            out["Gxwm0"] = t2n(model.Gxwm[:2])
            out["ywm0"] = t2n(model.ywm[:2])
            out["xwm0"] = t2n(model.xwm[:2])
    out["metric_keys"] = np.array(keys)
    out["metrics"] = np.array(rows, dtype=np.float64)
    sdG = model.G.state_dict()
    sdD = model.D.state_dict()
    out["G_checksum"] = np.array([float(v.double().sum()) for v in sdG.values()])
    out["G_abs_checksum"] = np.array([float(v.double().abs().sum()) for v in sdG.values()])
    out["D_checksum"] = np.array([float(v.double().sum()) for v in sdD.values()])
    out["D_abs_checksum"] = np.array([float(v.double().abs().sum()) for v in sdD.values()])
    out["G_keys"] = np.array(list(sdG.keys()))
    out["D_keys"] = np.array(list(sdD.keys()))
    out["ber"] = t2n(model.loss_model.compute_ber(model.G))
    out["state_keys"] = np.array(list(model.state_dict().keys()))
    np.savez_compressed(os.path.join(GOLD, "dcgan_step.npz"), **out)


def _wrap(models, Config, model, bbox, wbox):
    model = models.BlackBoxWrapper(model, Config(bbox))
    return models.WhiteBoxWrapper(model, Config(wbox))


def gen_srgan_cyclegan_steps(ref):
    """One protected pre-training step + one protected GAN step of IPR-SRGAN (24 -> 96, batch 2; VGG-19 with seeded random
    weights, the pretrained file needs the network) and one protected IPR-CycleGAN step (Resnet9Blocks, 64 x 64, batch 1)
    through the reference's own models / wrappers, driven as experiments/image_super_resolution.py:84-113 and
    experiments/image_translation.py:90-112 do."""
    import torchvision
    models, networks, Config = ref["models"], ref["networks"], ref["configs"].Config
    import networks.vgg as ref_vgg
    ref_vgg.vgg19 = lambda pretrained=True: torchvision.models.vgg19(weights=None)
    out = {}
    opt = {"lr": 1.0e-4, "betas": [0.9, 0.999]}
    torch.manual_seed(SEED)
    sr = models.SRGAN(Config({"G": "SRResNet", "D": "Discriminator96", "V": "VGG19Feature", "opt": "Adam", "opt_param": opt,
                              "type": "SRGAN"}), device=[torch.device("cpu")])
    sr = _wrap(models, Config, sr,
               {"fn_inp": {"type": "RandomNoisePatch", "size": 12}, "fn_out": {"size": 48, "opaque": True, "type": "PasteWatermark",
                                                                              "watermark": MARK},
                "lambda": 1.0, "loss_fn": "ssim", "normalized": False, "input_var": "low_res", "output_var": "super_res",
                "target": "G"},
               {"gamma_0": 0.1, "string": "EXAMPLE A", "target": "G"})
    g = torch.Generator().manual_seed(SEED)
    lr, hr = torch.rand(2, 3, 24, 24, generator=g), torch.rand(2, 3, 96, 96, generator=g)
    sr.update_g({"low_res": lr, "high_res": hr, "pretrain": True, "inhibit_bbox": True})
    m0 = sr.get_metrics()
    sr.update_g({"low_res": lr, "high_res": hr, "pretrain": False})
    sr.update_d({"high_res": sr.high_res, "super_res": sr.super_res})
    m1 = sr.get_metrics()
    out["sr_pre_keys"], out["sr_pre"] = np.array(sorted(m0)), np.array([m0[k] for k in sorted(m0)])
    out["sr_gan_keys"], out["sr_gan"] = np.array(sorted(m1)), np.array([m1[k] for k in sorted(m1)])
    out["sr_super_res"] = t2n(sr.super_res[:1, :, :8, :8])
    out["sr_state_keys"] = np.array(list(sr.state_dict().keys()))
    out["sr_sign_keys"] = np.array(list(sr.state_dict()["sign"].keys())[:3])

    torch.manual_seed(SEED)
    cg = models.CycleGAN(Config({"G": "Resnet9Blocks", "D": "ConvDiscriminator", "lambda_A": 10.0, "lambda_B": 10.0,
                                 "lambda_idt": 0.5, "opt": "Adam", "opt_param": {"lr": 2.0e-4, "betas": [0.5, 0.999]},
                                 "pool_size": 50, "epoch": 200, "type": "CycleGAN"}), device=[torch.device("cpu")])
    cg = _wrap(models, Config, cg,
               {"fn_inp": {"type": "RandomNoisePatch", "size": 32}, "fn_out": {"size": 32, "opaque": True, "type": "PasteWatermark",
                                                                              "watermark": MARK},
                "lambda": 1.0, "loss_fn": "ssim", "normalized": True, "input_var": "real_B", "output_var": "fake_A",
                "target": "GB"},
               {"gamma_0": 0.1, "string": "EXAMPLE A", "target": "GB"})
    a, b = torch.rand(1, 3, 64, 64, generator=g) * 2 - 1, torch.rand(1, 3, 64, 64, generator=g) * 2 - 1
    cg.update_g({"real_A": a, "real_B": b})
    cg.update_d({"real_A": cg.real_A, "real_B": cg.real_B, "fake_A": cg.fake_A.detach(), "fake_B": cg.fake_B.detach()})
    m2 = cg.get_metrics()
    out["cg_keys"], out["cg"] = np.array(sorted(m2)), np.array([m2[k] for k in sorted(m2)])
    out["cg_fake_A"] = t2n(cg.fake_A[:1, :, :8, :8])
    out["cg_state_keys"] = np.array(list(cg.state_dict().keys()))
    out["cg_ber"] = t2n(cg.loss_model.compute_ber(cg.GB))
    np.savez_compressed(os.path.join(GOLD, "srgan_cyclegan_steps.npz"), **out)


def gen_srgan_cyclegan_baseline_shapes(ref):
    """The same protected steps at the BASELINE.json shapes: config 3 = SRGAN 24 -> 96 at batch 16 (noise patch 12,
    watermark 48), config 4 = CycleGAN Resnet9Blocks at 128 x 128, batch 1, patch / watermark 64
    (configs/SRGAN/complete/srgan-imagenet-a.yaml, configs/CycleGAN/complete/cyclegan-city-a.yaml)."""
    import torchvision
    models, Config = ref["models"], ref["configs"].Config
    import networks.vgg as ref_vgg
    ref_vgg.vgg19 = lambda pretrained=True: torchvision.models.vgg19(weights=None)
    out = {}
    torch.manual_seed(SEED)
    sr = models.SRGAN(Config({"G": "SRResNet", "D": "Discriminator96", "V": "VGG19Feature", "opt": "Adam",
                              "opt_param": {"lr": 1.0e-4, "betas": [0.9, 0.999]}, "type": "SRGAN"}), device=[torch.device("cpu")])
    sr = _wrap(models, Config, sr,
               {"fn_inp": {"type": "RandomNoisePatch", "size": 12}, "fn_out": {"size": 48, "opaque": True, "type": "PasteWatermark",
                                                                              "watermark": MARK},
                "lambda": 1.0, "loss_fn": "ssim", "normalized": False, "input_var": "low_res", "output_var": "super_res",
                "target": "G"},
               {"gamma_0": 0.1, "string": "EXAMPLE A", "target": "G"})
    g = torch.Generator().manual_seed(SEED + 1)
    lr, hr = torch.rand(16, 3, 24, 24, generator=g), torch.rand(16, 3, 96, 96, generator=g)
    sr.update_g({"low_res": lr, "high_res": hr, "pretrain": False})
    sr.update_d({"high_res": sr.high_res, "super_res": sr.super_res})
    m = sr.get_metrics()
    out["sr_keys"], out["sr"] = np.array(sorted(m)), np.array([m[k] for k in sorted(m)])
    out["sr_super_res"] = t2n(sr.super_res[:2, :, 40:56, 40:56])
    torch.manual_seed(SEED)
    cg = models.CycleGAN(Config({"G": "Resnet9Blocks", "D": "ConvDiscriminator", "lambda_A": 10.0, "lambda_B": 10.0,
                                 "lambda_idt": 0.5, "opt": "Adam", "opt_param": {"lr": 2.0e-4, "betas": [0.5, 0.999]},
                                 "pool_size": 50, "epoch": 200, "type": "CycleGAN"}), device=[torch.device("cpu")])
    cg = _wrap(models, Config, cg,
               {"fn_inp": {"type": "RandomNoisePatch", "size": 64}, "fn_out": {"size": 64, "opaque": True, "type": "PasteWatermark",
                                                                              "watermark": MARK},
                "lambda": 1.0, "loss_fn": "ssim", "normalized": True, "input_var": "real_B", "output_var": "fake_A",
                "target": "GB"},
               {"gamma_0": 0.1, "string": "EXAMPLE A", "target": "GB"})
    a, b = torch.rand(1, 3, 128, 128, generator=g) * 2 - 1, torch.rand(1, 3, 128, 128, generator=g) * 2 - 1
    cg.update_g({"real_A": a, "real_B": b})
    cg.update_d({"real_A": cg.real_A, "real_B": cg.real_B, "fake_A": cg.fake_A.detach(), "fake_B": cg.fake_B.detach()})
    m2 = cg.get_metrics()
    out["cg_keys"], out["cg"] = np.array(sorted(m2)), np.array([m2[k] for k in sorted(m2)])
    out["cg_fake_A"] = t2n(cg.fake_A[:1, :, 56:72, 56:72])
    np.savez_compressed(os.path.join(GOLD, "srgan_cyclegan_baseline_shapes.npz"), **out)


def main():
    if "--baseline-shapes-only" in sys.argv:
        torch.set_num_threads(1)
        with ref_bridge.reference_modules() as ref:
            gen_srgan_cyclegan_baseline_shapes(ref)
        return
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(1)  # summation order of CPU reductions must not depend on the thread count
    with ref_bridge.reference_modules() as ref:
        gen_triggers(ref)
        gen_ssim(ref)
        gen_sign(ref)
        gen_phash(ref)
        gen_dcgan_step(ref)
        gen_srgan_cyclegan_steps(ref)
        gen_srgan_cyclegan_baseline_shapes(ref)
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
