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


# ---- BASELINE config 3: IPR-SRGAN 24 -> 96 (configs/SRGAN/complete/*.yaml: noise-patch trigger, pasted watermark, SSIM)
def srgan_model():
    return Config({"G": "SRResNet", "D": "Discriminator96", "V": "VGG19Feature", "opt": "Adam",
                   "opt_param": {"lr": 1.0e-4, "betas": [0.9, 0.999]}, "type": "SRGAN"})


def srgan_blackbox(watermark=WATERMARK_A):
    return Config({"fn_inp": {"type": "RandomNoisePatch", "size": 12},
                   "fn_out": {"size": 48, "opaque": True, "type": "PasteWatermark", "watermark": watermark},
                   "lambda": 1.0, "loss_fn": "ssim", "normalized": False, "input_var": "low_res",
                   "output_var": "super_res", "target": "G"})


# ---- BASELINE config 4: IPR-CycleGAN (configs/CycleGAN/complete/*.yaml: protects GB, InstanceNorm sign loss)
def cyclegan_model():
    return Config({"G": "Resnet9Blocks", "D": "ConvDiscriminator", "lambda_A": 10.0, "lambda_B": 10.0,
                   "lambda_idt": 0.5, "opt": "Adam", "opt_param": {"lr": 2.0e-4, "betas": [0.5, 0.999]},
                   "pool_size": 50, "epoch": 200, "type": "CycleGAN"})


def cyclegan_blackbox(watermark=WATERMARK_A):
    return Config({"fn_inp": {"type": "RandomNoisePatch", "size": 64},
                   "fn_out": {"size": 64, "opaque": True, "type": "PasteWatermark", "watermark": watermark},
                   "lambda": 1.0, "loss_fn": "ssim", "normalized": True, "input_var": "real_B",
                   "output_var": "fake_A", "target": "GB"})


def whitebox(target, gamma_0=0.1, string="EXAMPLE A"):
    return Config({"gamma_0": gamma_0, "string": string, "target": target})
