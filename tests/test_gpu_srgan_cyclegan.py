"""GPU: protected IPR-SRGAN (config 3 shape: 24 -> 96) and IPR-CycleGAN (config 4: Resnet9Blocks, InstanceNorm sign
loss) steps through the drop-in models / wrappers against metrics recorded from the UNMODIFIED reference
(tests/golden/srgan_cyclegan_steps.npz, oracle/make_golden.py).  Generators and discriminators run natively
(ipr_gan_b200/seqnet.py: tcgen05 convolutions with bf16 operands); triggers, SSIM watermark loss and sign loss run on the
sm_100a kernels; the frozen VGG feature extractor is outside the accelerated path (PyTorch).
Tolerance 2e-2 (north_star, bf16 convolutions) on every metric."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
SEED = 1234


def _cfg(d):
    from configs import Config
    return Config(d)


def _close(got, keys, want, tol=2e-2):
    assert sorted(got) == [str(k) for k in keys]
    for k, w in zip(keys, want):
        g = got[str(k)]
        assert abs(g - w) <= tol * max(1.0, abs(w)), (str(k), g, w)


def _rel(a, b):
    """relative Frobenius error of an image patch (24-37 bf16 blocks deep: 3e-2, see tests/test_gpu_seqnet.py)"""
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-12))


def test_srgan_protected_steps(golden, watermark_path):
    import models
    os.environ["IPR_VGG_RANDOM_INIT"] = "1"
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = golden("srgan_cyclegan_steps")
    dev = torch.device("cuda", 0)
    torch.manual_seed(SEED)
    sr = models.SRGAN(_cfg({"G": "SRResNet", "D": "Discriminator96", "V": "VGG19Feature", "opt": "Adam",
                            "opt_param": {"lr": 1.0e-4, "betas": [0.9, 0.999]}, "type": "SRGAN"}), device=[dev])
    sr = models.BlackBoxWrapper(sr, _cfg({"fn_inp": {"type": "RandomNoisePatch", "size": 12},
                                          "fn_out": {"size": 48, "opaque": True, "type": "PasteWatermark",
                                                     "watermark": watermark_path},
                                          "lambda": 1.0, "loss_fn": "ssim", "normalized": False, "input_var": "low_res",
                                          "output_var": "super_res", "target": "G"}))
    sr = models.WhiteBoxWrapper(sr, _cfg({"gamma_0": 0.1, "string": "EXAMPLE A", "target": "G"}))
    gen = torch.Generator().manual_seed(SEED)
    lr, hr = torch.rand(2, 3, 24, 24, generator=gen), torch.rand(2, 3, 96, 96, generator=gen)
    sr.update_g({"low_res": lr, "high_res": hr, "pretrain": True, "inhibit_bbox": True})
    _close(sr.get_metrics(), g["sr_pre_keys"], g["sr_pre"])
    sr.update_g({"low_res": lr, "high_res": hr, "pretrain": False})
    sr.update_d({"high_res": sr.high_res, "super_res": sr.super_res})
    _close(sr.get_metrics(), g["sr_gan_keys"], g["sr_gan"])
    assert _rel(sr.super_res[:1, :, :8, :8].detach().cpu().numpy(), g["sr_super_res"]) < 3e-2
    assert list(sr.state_dict().keys()) == [str(k) for k in g["sr_state_keys"]]
    assert list(sr.state_dict()["sign"].keys())[:3] == [str(k) for k in g["sr_sign_keys"]]
    assert sr.loss_model.compute_ber_counts(sr.G) == (0, 33 * 64)          # 2 112 signature bits in 33 BatchNorm layers


def test_cyclegan_protected_step(golden, watermark_path):
    import models
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = golden("srgan_cyclegan_steps")
    dev = torch.device("cuda", 0)
    torch.manual_seed(SEED)
    cg = models.CycleGAN(_cfg({"G": "Resnet9Blocks", "D": "ConvDiscriminator", "lambda_A": 10.0, "lambda_B": 10.0,
                               "lambda_idt": 0.5, "opt": "Adam", "opt_param": {"lr": 2.0e-4, "betas": [0.5, 0.999]},
                               "pool_size": 50, "epoch": 200, "type": "CycleGAN"}), device=[dev])
    cg = models.BlackBoxWrapper(cg, _cfg({"fn_inp": {"type": "RandomNoisePatch", "size": 32},
                                          "fn_out": {"size": 32, "opaque": True, "type": "PasteWatermark",
                                                     "watermark": watermark_path},
                                          "lambda": 1.0, "loss_fn": "ssim", "normalized": True, "input_var": "real_B",
                                          "output_var": "fake_A", "target": "GB"}))
    cg = models.WhiteBoxWrapper(cg, _cfg({"gamma_0": 0.1, "string": "EXAMPLE A", "target": "GB"}))
    gen = torch.Generator().manual_seed(SEED)
    for _ in range(2):                                   # the SRGAN fixture consumed these draws first
        torch.rand(2, 3, 24, 24, generator=gen) if _ == 0 else torch.rand(2, 3, 96, 96, generator=gen)
    a, b = torch.rand(1, 3, 64, 64, generator=gen) * 2 - 1, torch.rand(1, 3, 64, 64, generator=gen) * 2 - 1
    cg.update_g({"real_A": a, "real_B": b})
    cg.update_d({"real_A": cg.real_A, "real_B": cg.real_B, "fake_A": cg.fake_A.detach(), "fake_B": cg.fake_B.detach()})
    _close(cg.get_metrics(), g["cg_keys"], g["cg"])
    assert _rel(cg.fake_A[:1, :, :8, :8].detach().cpu().numpy(), g["cg_fake_A"]) < 3e-2
    assert list(cg.state_dict().keys()) == [str(k) for k in g["cg_state_keys"]]
    assert cg.loss_model.compute_ber_counts(cg.GB) == (0, 5248)            # 23 InstanceNorm layers of Resnet9Blocks


def test_srgan_step_at_config3_shape(golden, watermark_path):
    """BASELINE config 3: 16 x 3 x 24 x 24 -> 96 x 96, noise patch 12, watermark 48 (reference metrics: oracle/make_golden.py
    gen_srgan_cyclegan_baseline_shapes)."""
    import models
    os.environ["IPR_VGG_RANDOM_INIT"] = "1"
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = golden("srgan_cyclegan_baseline_shapes")
    dev = torch.device("cuda", 0)
    torch.manual_seed(SEED)
    sr = models.SRGAN(_cfg({"G": "SRResNet", "D": "Discriminator96", "V": "VGG19Feature", "opt": "Adam",
                            "opt_param": {"lr": 1.0e-4, "betas": [0.9, 0.999]}, "type": "SRGAN"}), device=[dev])
    sr = models.BlackBoxWrapper(sr, _cfg({"fn_inp": {"type": "RandomNoisePatch", "size": 12},
                                          "fn_out": {"size": 48, "opaque": True, "type": "PasteWatermark",
                                                     "watermark": watermark_path},
                                          "lambda": 1.0, "loss_fn": "ssim", "normalized": False, "input_var": "low_res",
                                          "output_var": "super_res", "target": "G"}))
    sr = models.WhiteBoxWrapper(sr, _cfg({"gamma_0": 0.1, "string": "EXAMPLE A", "target": "G"}))
    gen = torch.Generator().manual_seed(SEED + 1)
    lr, hr = torch.rand(16, 3, 24, 24, generator=gen), torch.rand(16, 3, 96, 96, generator=gen)
    sr.update_g({"low_res": lr, "high_res": hr, "pretrain": False})
    sr.update_d({"high_res": sr.high_res, "super_res": sr.super_res})
    _close(sr.get_metrics(), g["sr_keys"], g["sr"])
    assert _rel(sr.super_res[:2, :, 40:56, 40:56].detach().cpu().numpy(), g["sr_super_res"]) < 3e-2
    # the trigger pair is bit-exact whatever the networks do: noise patch pasted on the input, watermark on the output
    assert torch.equal(sr.xwm[:, :, :12, :12].cpu(), sr.fn_inp.module.fg.cpu().expand(16, -1, -1, -1))
    assert torch.equal(sr.ywm[:, :, 48:, :].cpu(), sr.super_res.detach()[:, :, 48:, :].cpu())


def test_cyclegan_step_at_config4_shape(golden, watermark_path):
    """BASELINE config 4: Resnet9Blocks at 128 x 128, batch 1, noise patch / watermark 64, InstanceNorm sign loss."""
    import models
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = golden("srgan_cyclegan_baseline_shapes")
    dev = torch.device("cuda", 0)
    torch.manual_seed(SEED)
    cg = models.CycleGAN(_cfg({"G": "Resnet9Blocks", "D": "ConvDiscriminator", "lambda_A": 10.0, "lambda_B": 10.0,
                               "lambda_idt": 0.5, "opt": "Adam", "opt_param": {"lr": 2.0e-4, "betas": [0.5, 0.999]},
                               "pool_size": 50, "epoch": 200, "type": "CycleGAN"}), device=[dev])
    cg = models.BlackBoxWrapper(cg, _cfg({"fn_inp": {"type": "RandomNoisePatch", "size": 64},
                                          "fn_out": {"size": 64, "opaque": True, "type": "PasteWatermark",
                                                     "watermark": watermark_path},
                                          "lambda": 1.0, "loss_fn": "ssim", "normalized": True, "input_var": "real_B",
                                          "output_var": "fake_A", "target": "GB"}))
    cg = models.WhiteBoxWrapper(cg, _cfg({"gamma_0": 0.1, "string": "EXAMPLE A", "target": "GB"}))
    gen = torch.Generator().manual_seed(SEED + 1)
    torch.rand(16, 3, 24, 24, generator=gen), torch.rand(16, 3, 96, 96, generator=gen)      # draws of the SRGAN fixture
    a, b = torch.rand(1, 3, 128, 128, generator=gen) * 2 - 1, torch.rand(1, 3, 128, 128, generator=gen) * 2 - 1
    cg.update_g({"real_A": a, "real_B": b})
    cg.update_d({"real_A": cg.real_A, "real_B": cg.real_B, "fake_A": cg.fake_A.detach(), "fake_B": cg.fake_B.detach()})
    _close(cg.get_metrics(), g["cg_keys"], g["cg"])
    assert _rel(cg.fake_A[:1, :, 56:72, 56:72].detach().cpu().numpy(), g["cg_fake_A"]) < 3e-2
    assert cg.loss_model.compute_ber_counts(cg.GB) == (0, 5248)


def _run_family(kind, use_graph, steps=3):
    """Metrics of `steps` protected steps on fresh host batches + the parameters afterwards (ipr_gan_b200/trainer.py)."""
    from ipr_gan_b200 import trainer as T
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(5)
    if kind == "srgan":
        tr = T.ProtectedSRGANTrainer(4, dev, use_graph=use_graph)
        draw = lambda: (torch.rand(4, 3, 24, 24, generator=g), torch.rand(4, 3, 96, 96, generator=g))
        nets = lambda m: (m.G, m.D)
        static = (tr.low_res, tr.high_res)
    else:
        tr = T.ProtectedCycleGANTrainer(dev, 64, use_graph=use_graph)
        draw = lambda: (torch.rand(1, 3, 64, 64, generator=g) * 2 - 1, torch.rand(1, 3, 64, 64, generator=g) * 2 - 1)
        nets = lambda m: (m.GA, m.GB, m.DA, m.DB)
        static = (tr.real_A, tr.real_B)
    for dst, src in zip(static, draw()):
        dst.copy_(src)
    tr.capture(warmup=2)
    if use_graph:
        assert (tr.graph if kind == "srgan" else tr.graph_g) is not None
    out = [tr.step_from_host(*draw()) for _ in range(steps)]
    torch.cuda.synchronize()
    params = [p.detach().clone() for n in nets(tr.model) for p in n.parameters()]
    return out, params


@pytest.mark.parametrize("kind", ["srgan", "cyclegan"])
def test_family_graph_replay_equals_eager(kind):
    """bench.py --workload srgan|cyclegan times CUDA-graph replays (one graph for the SRGAN step; generator-update and
    discriminator-update graphs around the eager image-pool exchange for CycleGAN): the replays must do what the eager
    reference-API calls do.  The native networks are deterministic; the frozen VGG (cuDNN) in SRGAN's content loss is
    held to 1e-4."""
    det = torch.backends.cudnn.deterministic
    torch.backends.cudnn.deterministic = True          # the frozen VGG (cuDNN): same algorithms eagerly and under capture
    try:
        m_e, p_e = _run_family(kind, False)
        m_g, p_g = _run_family(kind, True)
    finally:
        torch.backends.cudnn.deterministic = det
    # CycleGAN runs on this library only: replay == eager to rounding.  SRGAN's content loss goes through the frozen
    # PyTorch VGG: cuDNN picks its algorithms anew under capture, the two trajectories then separate by ~5e-4 in the
    # discriminator terms within three Adam steps.
    tol, ptol = (5e-3, 1e-3) if kind == "srgan" else (1e-4, 1e-5)
    for i, (a, b) in enumerate(zip(m_e, m_g)):
        assert sorted(a) == sorted(b)
        for k in a:
            assert abs(a[k] - b[k]) <= tol * max(1.0, abs(a[k])), (i, k, a[k], b[k])
    assert len({tuple(sorted(m.items())) for m in m_g}) == len(m_g)        # the replays consumed the new inputs
    for a, b in zip(p_e, p_g):
        assert float((a - b).abs().max()) <= ptol * max(1.0, float(a.abs().max()))
