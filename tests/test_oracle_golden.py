"""CPU: the restated oracle (oracle/ipr_oracle.py) against the vectors produced by the UNMODIFIED
reference (tests/golden/*.npz, made by oracle/make_golden.py).  Integer / copy arithmetic must be
bit-exact; floating point within the tolerance written at each check."""
import numpy as np
import torch

T = torch.from_numpy


def test_paste_and_crop_bit_exact(golden, oracle, watermark_path):
    g = golden("triggers")
    x = T(g["x"])
    for tag, opaque, norm, pos, size in (("op_tl", True, True, "tl", 16), ("al_br", False, True, "br", 16),
                                         ("al_tr_raw", False, False, "tr", 12), ("op_bl", True, False, "bl", 20)):
        fg, bg = oracle.load_watermark(watermark_path, size, opaque, norm)
        assert np.array_equal(fg.numpy(), g[f"paste_{tag}_fg"])
        assert np.array_equal(bg.numpy(), g[f"paste_{tag}_bg"])
        y = oracle.paste_patch(x, fg, bg, pos, size)
        assert np.array_equal(y.numpy().view(np.uint32), g[f"paste_{tag}_y"].view(np.uint32)), tag
        c = oracle.crop_patch(x, bg, pos, size)
        assert np.array_equal(c.numpy().view(np.uint32), g[f"paste_{tag}_crop"].view(np.uint32)), tag


def test_noise_patch_bit_exact(golden, oracle):
    g = golden("triggers")
    torch.manual_seed(1235)
    fg, bg = oracle.draw_noise_patch(12, False)
    assert np.array_equal(fg.numpy(), g["noise_fg"])
    x = T(g["noise_x"])
    assert np.array_equal(oracle.paste_patch(x, fg, bg, "br", 12).numpy(), g["noise_y"])
    assert np.array_equal(oracle.crop_patch(x, bg, "br", 12).numpy(), g["noise_crop"])


def test_latent_triggers(golden, oracle):
    g = golden("triggers")
    z = T(g["z"])
    torch.manual_seed(1236)
    mask = oracle.draw_bitmask(128, 10)
    assert np.array_equal(mask.numpy(), g["bitmask_mask"])
    assert np.array_equal(oracle.bitmask_scatter(z, mask, -10.0).numpy(), g["bitmask_y"])
    assert np.array_equal(oracle.transform_dist(z).numpy(), g["tdist_y"])
    torch.manual_seed(1237)
    a, w = oracle.draw_transform_var()
    assert np.array_equal(a.numpy(), g["tvar_a"]) and np.array_equal(w.numpy(), g["tvar_w"])
    assert np.array_equal(oracle.transform_var(z, a, w).numpy(), g["tvar_y"])
    # known answer from SURVEY.md 8a row A1: z = 0 -> sqrt(2 pi) / 2
    assert abs(float(oracle.transform_dist(torch.zeros(1))) - 1.2533141) < 1e-6


def test_ssim_loss_and_grad(golden, oracle):
    g = golden("ssim")
    for tag in "abc":
        x = T(g[f"{tag}_x"]).requires_grad_(True)
        y = T(g[f"{tag}_y"])
        norm = bool(g[f"{tag}_norm"])
        loss = oracle.ssim_loss(x, y, norm)
        loss.backward()
        assert np.allclose(loss.item(), g[f"{tag}_loss"], rtol=1e-6, atol=1e-7)
        assert np.allclose(x.grad.numpy(), g[f"{tag}_grad"], rtol=1e-5, atol=1e-9)
        # closed-form float64 gradient (independent of autograd); tolerance rel 1e-4 of the gradient scale
        l64, g64 = oracle.ssim_loss_grad_closed_form(x.detach(), y, norm)
        assert abs(float(l64) - float(g[f"{tag}_loss"])) < 1e-5
        scale = np.abs(g[f"{tag}_grad"]).max()
        assert np.abs(g64.numpy() - g[f"{tag}_grad"]).max() < 1e-4 * scale
    ps = oracle.ssim_per_sample(T(g["ps_x"]), T(g["ps_y"]))
    assert np.allclose(ps.numpy(), g["ps_ssim"], rtol=1e-6, atol=1e-7)
    assert abs(ps[0].item() - 1.0) < 1e-6      # identical pair -> SSIM 1


def test_ssim_taps_match_kernel_constants(oracle):
    taps = oracle.msssim.gauss_taps()
    expect = [float.fromhex(h) for h in ("0x1.0d957p-10", "0x1.f1fe02p-8", "0x1.26eb18p-5", "0x1.bff0fep-4",
                                          "0x1.b43c3ep-3", "0x1.10656p-2")]
    expect = expect + expect[-2::-1]
    assert [float(t) for t in taps] == expect    # same bit patterns as kTap[] in csrc/ssim.cu


def test_signature(golden, oracle):
    g = golden("sign")
    names = [str(n) for n in g["names"]]
    signs = oracle.signature_signs("EXAMPLE A", [len(g[f"sign_{n}"]) for n in names])
    gammas = []
    for n, s in zip(names, signs):
        assert np.array_equal(s.numpy(), g[f"sign_{n}"])
        gammas.append(T(g[f"gamma_{n}"]).clone().requires_grad_(True))
    loss = oracle.sign_loss(gammas, signs, 0.1)
    loss.backward()
    assert np.allclose(loss.item(), g["loss"], rtol=1e-6)
    for n, gm in zip(names, gammas):
        assert np.array_equal(gm.grad.numpy(), g[f"grad_{n}"])
    wrong, total = oracle.bit_error_rate([gm.detach() for gm in gammas], signs)
    assert total == 448 and np.float32(wrong) / np.float32(total) == g["ber"]
    # known answers: 80-bit period, first signs of layer 0
    bits = oracle.signature_bits("EXAMPLE A")
    assert len(bits) == 80 and "".join(map(str, bits[:16])) == "0100010101011000"
    assert signs[0][:8].tolist() == [-1, 1, -1, -1, -1, 1, -1, 1]


def test_phash_pvalue(golden, oracle):
    g = golden("phash")
    assert np.array_equal(oracle.pvalue_table(), g["ptable"])
    for tag in ("16", "48"):
        p, r = oracle.matching_prob(T(g[f"x{tag}"]), T(g[f"y{tag}"]))
        assert np.array_equal(p.numpy(), g[f"p{tag}"]), tag
    assert g["p16"][0] == 0.0                      # identical pair: r = 256 -> p = 1 - cdf(255) = 0
    hb = oracle.pdq().compute_batch(oracle.to_rgb_u8(T(g["x16_up"])))
    assert np.array_equal(hb, g["x16_hash"])
    assert int(hb[1].sum()) == 128                 # generic image: exactly half of the bits set
    # p-value table known answers
    t = g["ptable"]
    assert abs(t[128] - 0.52490955) < 1e-7 and abs(t[150] - 3.5406e-3) < 1e-6 and t[256] == 0.0


def test_bicubic_restatement_bit_exact(oracle):
    """The operation order the CUDA bicubic kernel follows, restated in numpy, against torch's CPU kernel."""
    from tests.bicubic_ref import bicubic_numpy
    torch.manual_seed(3)
    for (h, w, ho, wo) in ((16, 16, 32, 32), (24, 20, 38, 32), (11, 13, 32, 37), (31, 17, 58, 32)):
        x = torch.rand(3, 3, h, w)
        x[0] = 1.0
        ref = torch.nn.functional.interpolate(x, size=(ho, wo), mode="bicubic", align_corners=False).numpy()
        assert np.array_equal(bicubic_numpy(x.numpy(), ho, wo), ref), (h, w, ho, wo)


def test_dcgan_step_matches_reference_run(golden, oracle, watermark_path):
    g = golden("dcgan_step")
    torch.set_num_threads(1)
    torch.manual_seed(int(g["seed"]))
    G, D = oracle.make_generator(), oracle.make_discriminator()
    fg, bg = oracle.load_watermark(watermark_path, 16, True, True)
    step = oracle.DCGANStepOracle(G, D, oracle.transform_dist, lambda y: oracle.paste_patch(y, fg, bg, "tl", 16))
    gen = torch.Generator().manual_seed(int(g["seed"]))
    keys = [str(k) for k in g["metric_keys"]]
    for i in range(3):
        real = torch.randn(8, 3, 32, 32, generator=gen).clamp(-1, 1)
        z = torch.randn(8, 128, generator=gen)
        step.step(real, z)
        m = step.metrics()
        got = np.array([m[k] for k in keys])
        assert np.allclose(got, g["metrics"][i], rtol=1e-5, atol=1e-6), (i, dict(zip(keys, got - g["metrics"][i])))
        if i == 0:
            assert np.allclose(step.fake[:2].detach().numpy(), g["fake0"], rtol=1e-5, atol=1e-6)
            assert np.allclose(step.Gxwm[:2].detach().numpy(), g["Gxwm0"], rtol=1e-5, atol=1e-6)
            assert np.array_equal(step.ywm[:2].numpy(), oracle.paste_patch(step.fake[:2].detach(), fg, bg, "tl", 16).numpy())
    cs = np.array([float(v.double().sum()) for v in G.state_dict().values()])
    assert np.allclose(cs, g["G_checksum"], rtol=1e-4, atol=1e-3)
