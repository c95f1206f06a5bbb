"""Code-built configs with the reference's YAML schema (keys as in
configs/DCGAN/complete/dcgan-cifar10-a.yaml) for synthetic-data runs, tests and bench.py."""
import os

from configs import Config

_ASSETS = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "assets")
WATERMARK_A = os.path.join(_ASSETS, "watermark_a.png")


def dcgan_model(size=32):
    return Config({"G": "ConvGenerator%d" % size, "D": "SNDiscriminator%d" % size, "opt": "Adam",
                   "opt_param": {"lr": 2.0e-4, "betas": [0.5, 0.999]}, "type": "DCGAN"})


def dcgan_blackbox(fn_inp="TransformDist", wm_size=16, opaque=True, lam=1.0, watermark=WATERMARK_A):
    inp = {"type": fn_inp}
    if fn_inp == "RandomBitMask":
        inp.update({"n_bit": 10, "constant": -10.0, "z_dim": 128})
    return Config({"fn_inp": inp,
                   "fn_out": {"size": wm_size, "opaque": opaque, "type": "PasteWatermark", "watermark": watermark},
                   "lambda": lam, "loss_fn": "ssim",
                   # set by experiments/image_generation.py:63-67 before the wrapper is built
                   "normalized": True, "input_var": "latent", "output_var": "generated", "target": "G"})


def dcgan_whitebox(gamma_0=0.1, string="EXAMPLE A"):
    return Config({"gamma_0": gamma_0, "string": string, "target": "G"})
