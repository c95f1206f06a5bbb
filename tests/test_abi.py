"""CPU: the C-ABI library loads, exports every symbol include/ipr_b200.h declares, the ctypes table
matches the header, and the product path refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "ipr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ipr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from ipr_gan_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from ipr_gan_b200 import build
        build.build()
    handle = ctypes.CDLL(_lib.LIB_PATH)
    names = _header_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(handle, name), "header declares %s but the library does not export it" % name
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"
    lib = _lib.lib()
    assert lib.ipr_version() >= 100
    assert lib.ipr_strerror(0) == b"ok"
    assert b"NULL" in lib.ipr_strerror(-1)


def test_host_side_argument_checks_need_no_gpu():
    """Negative return codes are produced on the host before any launch."""
    from ipr_gan_b200 import _lib
    lib = _lib.lib()
    assert lib.ipr_transform_dist_f32(None, None, 16, None) == -1
    assert lib.ipr_paste_patch_f32(None, None, None, None, 1, 3, 32, 32, 16, 0, 0, None) == -1
    assert lib.ipr_ssim_workspace_bytes(512, 3, 32, 32) == 512 * 3 * 4
    assert lib.ipr_ssim_workspace_bytes(1, 3, 96, 96) == 3 * 9 * 4
    buf = (ctypes.c_float * 1024)()
    lib.ipr_pdq_dct_matrix_host(buf)
    from oracle import ipr_oracle
    import numpy as np
    assert np.array_equal(np.frombuffer(buf, dtype=np.float32), ipr_oracle.pdq().dct_matrix().ravel())


def test_no_cpu_fallback():
    from ipr_gan_b200 import ops
    x = torch.rand(2, 3, 32, 32)
    with pytest.raises(ops.IprError):
        ops.ssim_loss_fwd_bwd(x, x, False)
    with pytest.raises(ops.IprError):
        ops.transform_dist(torch.randn(4, 128))
    if not torch.cuda.is_available():
        import tools
        with pytest.raises(ops.IprError):
            tools.compute_matching_prob(torch.rand(2, 3, 16, 16), torch.rand(2, 3, 16, 16))
    # the networks are state containers on the CPU: their forward has no PyTorch-op path either
    import networks
    with pytest.raises(ops.IprError):
        networks.ConvGenerator32()(torch.randn(2, 128))
    with pytest.raises(ops.IprError):
        networks.SNDiscriminator32()(x)


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    """A missing libipr_b200.so is an error at the first call, never a silent fallback."""
    from ipr_gan_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libipr_b200.so"))
    with pytest.raises(Exception) as e:
        _lib.lib()
    assert "libipr_b200" in str(e.value)


def test_dropin_surface():
    import configs
    import models
    import networks
    import tools
    for name in ("PasteWatermark", "RandomNoisePatch", "RandomBitMask", "TransformDist", "TransformVar",
                 "SignLossModel", "compute_matching_prob", "l1", "mse", "ms_ssim", "ssim"):
        assert hasattr(tools, name), name
    for name in ("Model", "DCGAN", "BlackBoxWrapper", "WhiteBoxWrapper"):
        assert hasattr(models, name), name
    for name in ("ConvGenerator32", "ConvGenerator64", "SNDiscriminator32", "SNDiscriminator64"):
        assert hasattr(networks, name), name
    cfg = configs.Config({"a": {"b": 1}, "c": 2})
    assert cfg.a.b == 1 and cfg["c"] == 2 and cfg.get("zz", 5) == 5 and cfg.to_dict() == {"a": {"b": 1}, "c": 2}


def test_state_dict_keys_match_reference_format(golden):
    """Checkpoint format: network / wrapper state keys equal the reference's (golden from the reference run)."""
    import models
    from configs import presets
    g = golden("dcgan_step")
    model = models.DCGAN(presets.dcgan_model(), device=[torch.device("cpu")])
    assert list(model.G.state_dict().keys()) == ["module." + k if not k.startswith("module.") else k
                                                 for k in map(str, g["G_keys"])]
    assert list(model.D.state_dict().keys()) == [str(k) for k in g["D_keys"]]
