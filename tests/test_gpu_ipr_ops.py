"""GPU parity tests proper: every IPR kernel, called through the C ABI (ipr_gan_b200.ops -> ctypes ->
libipr_b200.so), against (a) the committed vectors produced by the unmodified reference and (b) the
CPU oracle on fresh seeded inputs, plus size-independent properties at BASELINE sizes.
Bit-exact for copy / integer / index work; rel 1e-4 for floating point (north_star)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
T = torch.from_numpy


def dev(a):
    return (T(a) if isinstance(a, np.ndarray) else a).cuda()


def bits(t):
    return t.detach().cpu().numpy().view(np.uint32)


# ----------------------------------------------------------------------------- triggers
def test_paste_and_crop_bit_exact_vs_reference(golden, watermark_path):
    import tools
    from configs import Config
    g = golden("triggers")
    x = dev(g["x"])
    for tag, opaque, norm, pos, size in (("op_tl", True, True, "tl", 16), ("al_br", False, True, "br", 16),
                                         ("al_tr_raw", False, False, "tr", 12), ("op_bl", True, False, "bl", 20)):
        cfg = Config({"size": size, "opaque": opaque, "watermark": watermark_path, "position": pos})
        m = tools.PasteWatermark(cfg, normalized=norm).cuda()
        assert np.array_equal(m.fg.cpu().numpy(), g[f"paste_{tag}_fg"])
        assert np.array_equal(m.bg.cpu().numpy(), g[f"paste_{tag}_bg"])
        assert np.array_equal(bits(m(x)), g[f"paste_{tag}_y"].view(np.uint32)), tag
        assert np.array_equal(bits(m.apply_mask(x)), g[f"paste_{tag}_crop"].view(np.uint32)), tag
        assert np.array_equal(bits(m(T(g["x"]))), g[f"paste_{tag}_y"].view(np.uint32))   # host input is staged


def test_noise_patch_and_unaligned_shapes(golden, oracle):
    import tools
    from configs import Config
    g = golden("triggers")
    torch.manual_seed(1235)
    m = tools.RandomNoisePatch(Config({"size": 12, "position": "br"}), normalized=False).cuda()
    assert np.array_equal(m.fg.cpu().numpy(), g["noise_fg"])
    x = dev(g["noise_x"])
    assert np.array_equal(bits(m(x)), g["noise_y"].view(np.uint32))
    assert np.array_equal(bits(m.apply_mask(x)), g["noise_crop"].view(np.uint32))
    # ragged: odd width (scalar path), window not multiple of 4, every corner
    from ipr_gan_b200 import ops
    torch.manual_seed(5)
    xr = torch.randn(3, 3, 19, 23)
    fg, bg = torch.rand(1, 3, 7, 7), (torch.rand(1, 1, 7, 7) > 0.5).float()
    for pos in ("tl", "tr", "bl", "br"):
        want = oracle.paste_patch(xr, fg, bg, pos, 7)
        assert np.array_equal(bits(ops.paste_patch(xr.cuda(), fg.cuda(), bg.cuda(), pos, 7)), want.numpy().view(np.uint32))
        want = oracle.crop_patch(xr, bg, pos, 7)
        assert np.array_equal(bits(ops.crop_patch(xr.cuda(), bg.cuda(), pos, 7)), want.numpy().view(np.uint32))


def test_paste_special_values(oracle):
    from ipr_gan_b200 import ops
    x = torch.zeros(1, 3, 32, 32)
    x[0, 0, 0, :4] = torch.tensor([-0.0, float("inf"), float("nan"), -1e-45])
    fg = torch.rand(1, 3, 16, 16)
    for bgv in (0.0, 1.0):
        bg = torch.full((1, 1, 16, 16), bgv)
        want = oracle.paste_patch(x, fg, bg, "tl", 16).numpy().view(np.uint32)
        got = bits(ops.paste_patch(x.cuda(), fg.cuda(), bg.cuda(), "tl", 16))
        nan = np.isnan(want.view(np.float32))
        assert np.array_equal(got[~nan], want[~nan]) and np.isnan(got.view(np.float32)[nan]).all()


def test_paste_roundtrip_at_baseline_size(watermark_path):
    """B=512 (C2): outside the window the batch is untouched, inside it equals fg; crop(paste(x)) == fg."""
    import tools
    from configs import Config
    m = tools.PasteWatermark(Config({"size": 16, "opaque": True, "watermark": watermark_path}), normalized=True).cuda()
    x = torch.randn(512, 3, 32, 32, device="cuda").clamp(-1, 1)
    y = m(x)
    assert torch.equal(y[..., :16, :16], m.fg.expand(512, -1, -1, -1))
    assert torch.equal(y[..., 16:, :], x[..., 16:, :]) and torch.equal(y[..., :16, 16:], x[..., :16, 16:])
    assert torch.equal(m.apply_mask(y), m.fg.expand(512, -1, -1, -1))
    assert torch.equal(m(y), y)                                    # idempotent


def test_latent_triggers(golden, oracle):
    import tools
    from configs import Config
    from ipr_gan_b200 import ops
    g = golden("triggers")
    z = dev(g["z"])
    torch.manual_seed(1236)
    m = tools.RandomBitMask(Config({"n_bit": 10, "constant": -10.0, "z_dim": 128})).cuda()
    assert np.array_equal(m.mask.cpu().numpy(), g["bitmask_mask"])
    assert np.array_equal(bits(m(z)), g["bitmask_y"].view(np.uint32))
    td = tools.TransformDist(Config({})).cuda()(z).cpu().numpy()
    assert np.allclose(td, g["tdist_y"], rtol=1e-4, atol=1e-6)          # erf: libdevice vs CPU libm, tol 1e-4
    torch.manual_seed(1237)
    tv = tools.TransformVar(Config({})).cuda()
    assert np.array_equal(tv.a.cpu().numpy(), g["tvar_a"]) and np.array_equal(tv.w.cpu().numpy(), g["tvar_w"])
    assert np.array_equal(bits(tv(z)), g["tvar_y"].view(np.uint32))
    # baseline size (C2) property: exactly n positions overwritten with c
    zz = torch.randn(512, 128, device="cuda")
    out = m(zz)
    changed = (out != zz)
    assert int(changed.sum()) == 512 * 10 and bool((out[changed] == -10.0).all())
    # fused trigger pair == separate calls
    x = torch.randn(64, 3, 32, 32, device="cuda")
    fg, bg = torch.rand(1, 3, 16, 16, device="cuda"), torch.zeros(1, 1, 16, 16, device="cuda")
    xwm, ywm = ops.trigger_pair(x, fg, bg, "tl", 16, zz[:64].contiguous())
    assert torch.equal(xwm, ops.transform_dist(zz[:64].contiguous())) and torch.equal(ywm, ops.paste_patch(x, fg, bg, "tl", 16))


# ----------------------------------------------------------------------------- SSIM
def _check_ssim(x, y, norm, want_loss, want_grad):
    from ipr_gan_b200 import ops
    loss, dx = ops.ssim_loss_fwd_bwd(x.cuda(), y.cuda(), norm)
    assert abs(loss.item() - float(want_loss)) <= 1e-4 * max(1.0, abs(float(want_loss)))
    scale = np.abs(want_grad).max()
    err = np.abs(dx.cpu().numpy() - want_grad).max()
    assert err <= 1e-4 * scale, (err, scale)        # rel 1e-4 of the gradient scale (north_star fp32 tolerance)


def test_ssim_vs_reference_vectors(golden):
    g = golden("ssim")
    for tag in "abc":
        _check_ssim(T(g[f"{tag}_x"]), T(g[f"{tag}_y"]), bool(g[f"{tag}_norm"]), g[f"{tag}_loss"], g[f"{tag}_grad"])


@pytest.mark.parametrize("shape,norm", [((64, 3, 32, 32), True), ((16, 3, 96, 96), False), ((1, 3, 128, 128), True),
                                        ((5, 3, 11, 11), False), ((3, 1, 45, 70), False), ((7, 3, 33, 64), True)])
def test_ssim_vs_oracle(oracle, shape, norm):
    torch.manual_seed(sum(shape))
    x = torch.rand(*shape)
    y = (x + 0.2 * torch.randn(*shape)).clamp(0, 1)
    if norm:
        x, y = x * 2 - 1, y * 2 - 1
    loss64, grad64 = oracle.ssim_loss_grad_closed_form(x, y, norm)
    _check_ssim(x, y, norm, float(loss64), grad64.float().numpy())
    xr = x.clone().requires_grad_(True)
    ref = oracle.ssim_loss(xr, y, norm)
    ref.backward()
    _check_ssim(x, y, norm, ref.item(), xr.grad.numpy())


def test_ssim_properties_at_baseline_size():
    from ipr_gan_b200 import ops
    x = torch.rand(512, 3, 32, 32, device="cuda")
    loss, dx = ops.ssim_loss_fwd_bwd(x, x.clone(), False)
    assert abs(loss.item()) < 1e-6                                   # ssim(x, x) = 1
    assert float(dx.abs().max()) < 1e-6
    y = torch.rand(512, 3, 32, 32, device="cuda")
    l1, _ = ops.ssim_loss_fwd_bwd(x, y, False, need_grad=False)
    l2, _ = ops.ssim_loss_fwd_bwd(y, x, False, need_grad=False)
    assert abs(l1.item() - l2.item()) < 1e-6                         # symmetry
    la, da = ops.ssim_loss_fwd_bwd(x, y, False)
    lb, db = ops.ssim_loss_fwd_bwd(x, y, False)
    assert torch.equal(la, lb) and torch.equal(da, db)               # run-to-run deterministic
    _, d3 = ops.ssim_loss_fwd_bwd(x, y, False, grad_scale=3.0)
    assert torch.allclose(d3, 3.0 * da, rtol=1e-6, atol=0)           # linear in grad_scale
    # mean of the per-sample values == the batch loss
    ps = ops.ssim_per_sample(x, y)
    assert abs((1 - ps.mean()).item() - la.item()) < 1e-5


def test_ssim_autograd_and_per_sample(golden, oracle):
    import pytorch_msssim
    import tools
    g = golden("ssim")
    ps = pytorch_msssim.ssim(T(g["ps_x"]), T(g["ps_y"]), data_range=1, size_average=False)   # host tensors, as eval does
    assert ps.device.type == "cpu" and np.allclose(ps.numpy(), g["ps_ssim"], rtol=1e-4, atol=1e-6)
    x = dev(g["a_x"]).requires_grad_(True)
    loss = tools.ssim(normalized=True)(x, dev(g["a_y"]))
    (2.5 * loss).backward()
    assert np.abs(x.grad.cpu().numpy() - 2.5 * g["a_grad"]).max() <= 1e-4 * 2.5 * np.abs(g["a_grad"]).max()
    # verification shape (C5): 10 000 crops of 16x16
    torch.manual_seed(9)
    wx = torch.rand(10000, 3, 16, 16)
    wy = (wx + 0.1 * torch.randn_like(wx)).clamp(0, 1)
    want = oracle.ssim_per_sample(wx, wy)
    from ipr_gan_b200 import ops
    got = ops.ssim_per_sample(wx.cuda(), wy.cuda()).cpu()
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-6)


# ----------------------------------------------------------------------------- signature
def test_sign_loss_and_ber(golden, oracle):
    from ipr_gan_b200 import ops
    g = golden("sign")
    names = [str(n) for n in g["names"]]
    gammas = [dev(g[f"gamma_{n}"]) for n in names]
    signs = [dev(g[f"sign_{n}"]) for n in names]
    loss, grads = ops.sign_loss_fwd_bwd(gammas, signs, 0.1)
    assert abs(loss.item() - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    for n, gr in zip(names, grads):
        assert np.array_equal(gr.cpu().numpy(), g[f"grad_{n}"]), n        # -sign/C where active: exact
    counts = ops.sign_ber_counts(gammas, signs).tolist()
    wrong, total = oracle.bit_error_rate([T(g[f"gamma_{n}"]) for n in names], [T(g[f"sign_{n}"]) for n in names])
    assert counts == [wrong, total]                                       # integer, bit-exact
    assert np.float32(counts[0]) / np.float32(counts[1]) == g["ber"]
    # accumulate mode adds into existing gradient buffers
    base = [torch.ones_like(x) for x in gammas]
    ops.sign_loss_fwd_bwd(gammas, signs, 0.1, grad_scale=2.0, grads=base, accumulate=True)
    for n, b in zip(names, base):
        assert np.array_equal(b.cpu().numpy(), 1 + 2 * g[f"grad_{n}"])


def test_sign_model_module(oracle):
    """tools.SignLossModel on the product generator: signs, gamma rewrite, loss 0 / BER 0 at init, known answers."""
    import networks
    import tools
    from configs import Config
    torch.manual_seed(0)
    G = networks.ConvGenerator32().cuda()
    sm = tools.SignLossModel(G, Config({"gamma_0": 0.1, "string": "EXAMPLE A"})).cuda()
    want = oracle.signature_signs("EXAMPLE A", [256, 128, 64])
    for (safe, m), s in zip(oracle.norm_layers(G), want):
        assert torch.equal(getattr(sm, safe).cpu(), s) and bool((m.weight.sign().cpu() == s).all())
    assert list(sm.state_dict().keys()) == ["convs_0_1", "convs_1_1", "convs_2_1"]
    assert sm(G).item() == 0.0 and sm.compute_ber(G).item() == 0.0
    with torch.no_grad():
        for (_, m), s in zip(oracle.norm_layers(G), want):
            m.weight.copy_(0.05 * s.cuda())
    loss = sm(G)
    assert abs(loss.item() - 0.15) < 1e-6                                 # 3 layers x (0.1 - 0.05)
    loss.backward()
    for (_, m), s in zip(oracle.norm_layers(G), want):
        assert torch.equal(m.weight.grad.cpu(), -s / s.numel())
    # flip 10 % of the signs (sign_flip.py:59-75): BER == flipped fraction exactly
    with torch.no_grad():
        G.convs[0][1].weight[:44] *= -1
    assert sm.compute_ber_counts(G) == (44, 448)
    # many layers (> table size is chunked): 70 layers of 64
    gam = [torch.randn(64, device="cuda") for _ in range(70)]
    sg = [torch.sign(torch.randn(64, device="cuda")) for _ in range(70)]
    from ipr_gan_b200 import ops
    loss, _ = ops.sign_loss_fwd_bwd(gam, sg, 0.1)
    ref = oracle.sign_loss([x.cpu() for x in gam], [x.cpu() for x in sg], 0.1)
    assert abs(loss.item() - ref.item()) <= 1e-4 * abs(ref.item())
    assert ops.sign_ber_counts(gam, sg).tolist() == list(oracle.bit_error_rate([x.cpu() for x in gam], [x.cpu() for x in sg]))


# ----------------------------------------------------------------------------- verification
def test_bicubic_bit_exact():
    from ipr_gan_b200 import ops
    torch.manual_seed(3)
    for (h, w, ho, wo) in ((16, 16, 32, 32), (24, 20, 38, 32), (11, 13, 32, 37), (31, 17, 58, 32), (12, 12, 32, 32)):
        x = torch.rand(4, 3, h, w)
        x[0] = 1.0
        x[1, :, : h // 2] = 1.0
        x[1, :, h // 2:] = 0.0
        ref = torch.nn.functional.interpolate(x, size=(ho, wo), mode="bicubic", align_corners=False)
        got = ops.bicubic_resize(x.cuda(), ho, wo).cpu()
        assert np.array_equal(bits(got), bits(ref)), (h, w, ho, wo)


def test_pdq_hash_and_pvalue_vs_reference_vectors(golden, oracle):
    import tools
    from ipr_gan_b200 import ops
    g = golden("phash")
    assert np.array_equal(ops.pvalue_table_host(), g["ptable"])
    hb = ops.unpack_hash_bits(ops.pdq_hash(dev(g["x16_up"])))
    assert np.array_equal(hb, g["x16_hash"])                              # hash bits, bit-exact
    hb = ops.unpack_hash_bits(ops.pdq_hash(dev(g["x48"])))
    assert np.array_equal(hb, g["x48_hash"])
    for tag in ("16", "48"):
        p = tools.compute_matching_prob(T(g[f"x{tag}"]), T(g[f"y{tag}"]))   # host tensors in, host tensor out
        assert p.device.type == "cpu" and p.dtype == torch.float32
        assert np.array_equal(p.numpy(), g[f"p{tag}"]), tag                # p-values bit-exact
    _, r = ops.matching_prob(dev(g["x16"]), dev(g["y16"]))
    _, want_r = oracle.matching_prob(T(g["x16"]), T(g["y16"]))
    assert np.array_equal(r.cpu().numpy(), want_r)


@pytest.mark.parametrize("size", [32, 48, 64, 33])
def test_pdq_vs_oracle_sizes(oracle, size):
    from ipr_gan_b200 import ops
    torch.manual_seed(size)
    x = torch.rand(40, 3, size, size)
    x[0] = 0.5
    x[1] = torch.nn.functional.interpolate(torch.rand(1, 3, 4, 4), size=size, mode="bilinear")[0]
    got, coeffs = ops.pdq_hash(x.cuda(), want_coeffs=True)
    want = oracle.pdq().compute_batch(oracle.to_rgb_u8(x))
    _, wc = oracle.pdq().compute_with_coeffs(oracle.to_rgb_u8(x)[3])
    assert np.array_equal(coeffs[3].cpu().numpy().view(np.uint32), wc.view(np.uint32))   # DCT block bit-exact
    assert np.array_equal(ops.unpack_hash_bits(got), want)


def test_verification_sweep_shape(oracle):
    """C5 shape, reduced to 2 000 pairs so the CPU oracle finishes in seconds: r and p bit-exact, hash(x)^hash(x) = 0."""
    from ipr_gan_b200 import ops
    torch.manual_seed(11)
    base = torch.nn.functional.interpolate(torch.rand(2000, 3, 5, 5), size=16, mode="bilinear")
    wx = (0.7 * base + 0.3 * torch.rand(2000, 3, 16, 16)).clamp(0, 1)
    wy = (wx + 0.08 * torch.randn_like(wx)).clamp(0, 1)
    wy[:10] = torch.rand(10, 3, 16, 16)
    p, r = ops.matching_prob(wx.cuda(), wy.cuda())
    wp, wr = oracle.matching_prob(wx, wy)
    assert np.array_equal(r.cpu().numpy(), wr) and np.array_equal(p.cpu().numpy(), wp.numpy())
    p0, r0 = ops.matching_prob(wx.cuda(), wx.cuda())
    assert bool((r0 == 256).all()) and bool((p0 == 0).all())


def test_crop_postproc_fused_bit_exact():
    """The evaluation loop's ``postproc(apply_mask(x))`` (experiments/image_generation.py:141-149, 208-209) as one launch:
    crop, clamp(-1, 1), (x + 1) / 2 -- bit-identical to the three torch operations, special values included."""
    from ipr_gan_b200 import ops
    torch.manual_seed(3)
    x = torch.randn(9, 3, 32, 32) * 1.5
    x[0, 0, 0, :6] = torch.tensor([float("nan"), float("inf"), -float("inf"), -0.0, 1.0, -1.0])
    for pos in ("tl", "tr", "bl", "br"):
        for s in (16, 5):
            bg = torch.zeros(1, 1, s, s)
            want = (ops.crop_patch(x.cuda(), bg.cuda(), pos, s).cpu().clamp(-1, 1) + 1.0) / 2.0
            got = ops.crop_patch(x.cuda(), bg.cuda(), pos, s, postproc=True).cpu()
            assert torch.equal(torch.nan_to_num(got, nan=7.0), torch.nan_to_num(want, nan=7.0))
            assert torch.equal(got.isnan(), want.isnan())
