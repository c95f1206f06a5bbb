"""Watermark verification: ``compute_matching_prob(img1, img2, min_size=32) -> FloatTensor (B,)`` on
the CPU, the reference's signature (tools/phash_pvalue.py:19-38).  The reference loops over images
in Python and hashes on the CPU; here the batch is staged to the GPU once and the bicubic
up-sampling, uint8 conversion, PDQ hash, Hamming distance and p-value lookup run as batched
sm_100a kernels (csrc/pdq.cu).
"""
import torch

from ipr_gan_b200 import ops


def _device():
    if not torch.cuda.is_available():
        raise ops.IprError("compute_matching_prob needs a CUDA device: no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def compute_hash(img_tensor):
    """(B,3,H,W) in [0,1] -> (B,256) bool array like the reference's helper (tools/phash_pvalue.py:7-17)."""
    x = img_tensor.detach().to(_device(), torch.float32)
    return ops.unpack_hash_bits(ops.pdq_hash(x)).astype(bool)


def compute_matching_prob(img1, img2, min_size=32):
    dev = img1.device if img1.is_cuda else _device()
    x = img1.detach().to(dev, torch.float32, non_blocking=True)
    y = img2.detach().to(dev, torch.float32, non_blocking=True)
    p, _r = ops.matching_prob(x, y, min_size)
    return p.cpu()


def compute_matching_prob_device(img1, img2, min_size=32):
    """Same computation without the host round trip: -> (p, r) CUDA tensors."""
    return ops.matching_prob(img1, img2, min_size)
