"""CPU (build container only -- needs the reference checkout): the reference's OWN entry package ``experiments``
imports and builds its ``ImageGeneration`` experiment on top of the drop-in packages
(experiments/image_generation.py:1-84, experiments/base.py:10-39), the way train.py / eval.py / sign_flip.py do.

What can run without a GPU: imports (incl. the pass-through names ``networks.InceptionActivations``, ``models.VAE``),
construction of DCGAN -> BlackBoxWrapper -> WhiteBoxWrapper from the reference's YAML, the checkpoint dictionary, and
a strict load of it into a second experiment (eval.py:31-40).  The step itself needs CUDA (tests/test_gpu_dcgan.py)."""
import os
import subprocess
import sys

import pytest

REF = os.environ.get("IPR_REFERENCE_ROOT", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "experiments", "image_generation.py")),
                                reason="reference checkout not present (GPU box)")

_SCRIPT = r'''
import os, sys, types
ROOT, REF, TMP = sys.argv[1:4]
sys.path.insert(0, ROOT)
import ipr_gan_b200
ipr_gan_b200.enable_dropin(reference_root=REF)
sys.path.insert(2, REF)                      # as if launched from the reference directory, drop-ins first
try:
    import skimage                            # image_super_resolution.py:4; not installed in this image
except ImportError:
    sk, skm = types.ModuleType("skimage"), types.ModuleType("skimage.metrics")
    skm.peak_signal_noise_ratio = skm.structural_similarity = None
    sk.metrics = skm
    sys.modules["skimage"], sys.modules["skimage.metrics"] = sk, skm
import torch
import experiments, models, networks, tools, datasets, configs
assert experiments.__file__.startswith(REF) and datasets.__file__.startswith(REF)
for m in (models, networks, tools, configs):
    assert m.__file__.startswith(os.path.join(ROOT, "ipr_gan_b200", "dropin")), m.__file__
# names outside the accelerated path resolve to the reference's own modules
assert networks.InceptionActivations.__module__ == "networks.inception"
assert models.VAE.__module__ == "models.vae" and issubclass(models.VAE, models.Model)
assert networks.Encoder32.__module__ == "networks.encoder" and networks.Decoder32.__module__ == "networks.decoder"
assert callable(tools.ms_ssim(normalized=True))

cfg = configs.Config.parse(os.path.join(REF, "configs", "DCGAN", "complete", "dcgan-cifar10-a.yaml"))
cfg.log.path = os.path.join(TMP, "log")
cfg.resource.gpu = False
cfg.hparam.bsz = 4
cfg.protection.bbox.fn_out.watermark = os.path.join(ROOT, "ipr_gan_b200", "assets", "watermark_a.png")

class FakeLoader(object):
    def __len__(self): return 8
    def __next__(self): return torch.rand(4, 3, 32, 32) * 2 - 1, torch.zeros(4)
datasets.cifar10 = lambda **kw: FakeLoader()

exp = experiments.ImageGeneration(cfg)
assert exp.bbox and exp.wbox
assert type(exp.model).__name__ == "WhiteBoxWrapper" and type(exp.model.model).__name__ == "BlackBoxWrapper"
sd = exp.model.state_dict()
assert list(sd.keys()) == ["G", "D", "optG", "optD", "fn_inp", "fn_out", "sign"], list(sd.keys())
assert "module.convs.0.1.weight" in sd["G"] and "module.net.0.0.weight_orig" in sd["D"]
assert sorted(sd["sign"].keys()) == ["module_convs_0_1", "module_convs_1_1", "module_convs_2_1"]
sd["step"] = 7
exp2 = experiments.ImageGeneration(cfg)
exp2.load_state_dict(sd, strict=True)
assert exp2.init_step == 8
for a, b in zip(exp.model.G.parameters(), exp2.model.G.parameters()):
    assert torch.equal(a, b)
# the wrapper falls through to the inner model and answers None for unknown names (models/base.py:52-58)
assert exp.model.fake_sample is None and exp.model.fn_inp is not None
# several devices in ONE process are refused with the torchrun instruction
try:
    models.Replica(torch.nn.Linear(2, 2), device_ids=[0, 1])
    raise SystemExit("Replica accepted two devices")
except RuntimeError as e:
    assert "torch.distributed.run" in str(e)
# the step itself computes on CUDA only: no CPU fallback
try:
    exp.train()
    raise SystemExit("train() ran on the CPU")
except Exception as e:
    assert "CUDA" in str(e) or "cuda" in str(e), e
print("OK")
'''


def test_reference_experiments_run_over_dropin(tmp_path):
    env = dict(os.environ)
    env.pop("IPR_REFERENCE_ROOT", None)
    r = subprocess.run([sys.executable, "-c", _SCRIPT, ROOT, REF, str(tmp_path)], capture_output=True, text=True,
                       env=env, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout[-2000:] + r.stderr[-4000:]
