"""CPU: checkpoint / wire-format compatibility.  A checkpoint with the reference's structure
(models/base.py:24-38; DataParallel ``module.`` prefixes; legacy spectral-norm ``weight_orig/_u/_v``; Adam state) written
from the oracle's reference-architecture networks loads into the drop-in stack, and the drop-in stack's own checkpoint
has exactly the reference's keys (golden list recorded from the unmodified reference)."""
import io

import torch


def _reference_style_checkpoint():
    from oracle import ipr_oracle as orc
    torch.manual_seed(11)
    G, D = orc.make_generator(), orc.make_discriminator()
    optG = torch.optim.Adam(G.parameters(), lr=2e-4, betas=(0.5, 0.999))
    optD = torch.optim.Adam(D.parameters(), lr=2e-4, betas=(0.5, 0.999))
    x = torch.randn(4, 3, 32, 32)
    D(G(torch.randn(4, 128))).mean().backward()
    D(x).mean().backward()
    optG.step(), optD.step()
    pre = lambda sd: {"module." + k: v for k, v in sd.items()}
    sign = {n: s for n, s in zip(("module_convs_0_1", "module_convs_1_1", "module_convs_2_1"),
                                 orc.signature_signs("EXAMPLE A", [256, 128, 64]))}
    ck = {"G": pre(G.state_dict()), "D": pre(D.state_dict()), "optG": optG.state_dict(), "optD": optD.state_dict(),
          "fn_inp": {}, "fn_out": {"module.bg": torch.zeros(1, 1, 16, 16), "module.fg": torch.rand(1, 3, 16, 16)},
          "sign": sign, "step": 1000}
    buf = io.BytesIO()
    torch.save(ck, buf)
    buf.seek(0)
    return torch.load(buf, map_location="cpu"), G, D


def test_reference_style_checkpoint_loads_into_dropin(golden):
    import models
    from configs import presets
    ck, G, D = _reference_style_checkpoint()
    model = models.DCGAN(presets.dcgan_model(), device=[torch.device("cpu")])
    model = models.BlackBoxWrapper(model, presets.dcgan_blackbox())
    model = models.WhiteBoxWrapper(model, presets.dcgan_whitebox())
    model.load_state_dict(ck, strict=True)                      # asserts every entry is present (models/base.py:26-28)
    for (n, p), (_, q) in zip(model.G.module.named_parameters(), G.named_parameters()):
        assert torch.equal(p, q), n
    for (n, b), (_, c) in zip(model.D.module.named_buffers(), D.named_buffers()):
        assert torch.equal(b, c), n                             # weight_u / weight_v power-iteration state
    assert torch.equal(model.fn_out.module.fg, ck["fn_out"]["module.fg"])
    assert model.optG.state_dict()["param_groups"][0]["betas"] == (0.5, 0.999)
    # and back: the drop-in stack's checkpoint has the reference's keys, in the reference's order
    out = model.state_dict()
    g = golden("dcgan_step")
    assert list(out.keys()) == [str(k) for k in g["state_keys"]]
    assert list(out["G"].keys()) == [str(k) for k in g["G_keys"]]
    assert list(out["D"].keys()) == [str(k) for k in g["D_keys"]]
    assert list(out["sign"].keys()) == ["module_convs_0_1", "module_convs_1_1", "module_convs_2_1"]
    assert set(out["optG"].keys()) == {"state", "param_groups"}


def test_wrapper_attribute_fallthrough():
    """models/base.py:52-58: unknown attributes of a Wrapper are None, known ones fall through to the wrapped model."""
    import models
    from configs import presets
    model = models.DCGAN(presets.dcgan_model(), device=[torch.device("cpu")])
    w = models.BlackBoxWrapper(model, presets.dcgan_blackbox())
    assert w.optG is model.optG and w.no_such_attribute is None and hasattr(w, "anything_at_all")
    assert w._modules is model._modules and "fn_inp" in model._modules
