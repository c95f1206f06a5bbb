"""GPU: the on-device verification sweep against the oracle's CPU evaluation of the same generator outputs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_verification_sweep_matches_oracle(watermark_path):
    import models
    import tools
    from configs import presets
    from ipr_gan_b200 import ops, verify
    from oracle import ipr_oracle as orc
    torch.manual_seed(3)
    dev = torch.device("cuda", 0)
    model = models.DCGAN(presets.dcgan_model(), device=[dev])
    model = models.BlackBoxWrapper(model, presets.dcgan_blackbox())
    model = models.WhiteBoxWrapper(model, presets.dcgan_whitebox())
    G = model.G                      # the Replica: the signature buffers are named from this module tree
    res = verify.verification_sweep(G, model.fn_inp, model.fn_out, n_samples=96, batch=40, p_thres=0.01,
                                    sign_model=model.loss_model, return_per_sample=True)
    assert res["N"] == 96 and res["WBOX_counts"] == (0, 448)
    # oracle on the SAME crops (the generator itself is covered by test_gpu_dcgan.py): regenerate them on the GPU
    gen = torch.Generator().manual_seed(1234)
    G.eval()
    qs, ps, rs = [], [], []
    with torch.no_grad():
        for b in (40, 40, 16):
            z = torch.randn(b, 128, generator=gen).to(dev)
            x = G(z)
            xwm, ywm = G(model.fn_inp(z)), model.fn_out(x)
            cx = ((xwm[..., :16, :16].clamp(-1, 1) + 1) / 2).cpu()
            cy = ((ywm[..., :16, :16].clamp(-1, 1) + 1) / 2).cpu()
            qs.append(orc.ssim_per_sample(cx, cy))
            p, r = orc.matching_prob(cx, cy)
            ps.append(p), rs.append(torch.from_numpy(r))
    q, p, r = torch.cat(qs), torch.cat(ps), torch.cat(rs)
    per = res["per_sample"]
    assert np.array_equal(per["r"].cpu().numpy(), r.numpy())                 # Hamming agreement count: bit-exact
    assert np.array_equal(per["p"].cpu().numpy(), p.numpy())                 # p-values: bit-exact
    assert torch.allclose(per["q"].cpu(), q, rtol=1e-4, atol=1e-6)           # per-sample SSIM: 1e-4
    assert res["MATCH"] == int((p < 0.01).sum())
    assert abs(res["P"] - float(p.double().mean())) < 1e-9 and abs(res["Q_WM"] - float(q.double().mean())) < 1e-5
    # sign flip of 10 % of the signature (sign_flip.py:59-75): BER counts exactly the flipped bits
    with torch.no_grad():
        G.module.convs[0][1].weight[:44] *= -1
    res2 = verify.verification_sweep(G, model.fn_inp, model.fn_out, n_samples=8, batch=8, sign_model=model.loss_model)
    assert res2["WBOX_counts"] == (44, 448)
