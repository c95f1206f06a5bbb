"""Black-box trigger modules with the reference's constructor / forward / reset / apply_mask surface,
computing on sm_100a kernels through the C ABI.

Reference interfaces mirrored (file:line in the reference tree):
  PasteWatermark    tools/paste_watermark.py:6-61      buffers fg (1,3,s,s), bg (1,1,s,s)
  RandomNoisePatch  tools/random_noise_patch.py:6-54   buffers fg, bg; reset() re-draws from the CPU RNG
  RandomBitMask     tools/random_bitmask.py:4-30       buffer _mask (1,n) int64; .mask property
  TransformDist     tools/transform_dist.py:5-13
  TransformVar      tools/transform_var.py:5-17        buffers w, a (1,128)
Outputs are new tensors computed under no_grad, on the module's device (a CPU input is staged to
that device first, which is what the reference's DataParallel wrapper does).
"""
import torch
import torch.nn as nn

from ipr_gan_b200 import ops

_POSITIONS = ("tl", "tr", "bl", "br")


def _device_of(module):
    for b in module.buffers():
        return b.device
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


def _stage(x, device):
    if device.type != "cuda":
        raise ops.IprError("trigger modules compute on CUDA only; call .to('cuda') on the module first")
    return x.detach().to(device=device, dtype=torch.float32, non_blocking=True)


class _PatchTrigger(nn.Module):
    """Shared forward/apply_mask of the two image-space triggers (same arithmetic in the reference)."""

    def _setup(self, config, kwargs):
        self.config = config
        self.normalized = kwargs.get("normalized", False)
        self.position = config.get("position", "tl")
        assert self.position in _POSITIONS, "invalid position"

    def _set_patch(self, fg, bg):
        size = self.config.size
        device = self.fg.device if hasattr(self, "fg") else fg.device
        if self.normalized:
            fg = (fg - 0.5) / 0.5
        self.register_buffer("bg", bg.reshape(1, 1, size, size).float().to(device))
        self.register_buffer("fg", fg.reshape(1, 3, size, size).float().to(device))

    def forward(self, x):
        with torch.no_grad():
            x = _stage(x, self.fg.device)
            return ops.paste_patch(x, self.fg, self.bg, self.position, self.config.size)

    def apply_mask(self, x):
        with torch.no_grad():
            x = _stage(x, self.fg.device)
            return ops.crop_patch(x, self.bg, self.position, self.config.size)


class PasteWatermark(_PatchTrigger):
    def __init__(self, config, **kwargs):
        super().__init__()
        self._setup(config, kwargs)
        self._load_mark()

    def _load_mark(self):
        # host-side, one-off: identical PIL steps to tools/paste_watermark.py:15-30
        from PIL import Image
        from torchvision.transforms import functional as TF
        dims = (self.config.size,) * 2
        mark = TF.resize(Image.open(self.config.watermark).convert("RGBA"), dims)
        sheet = Image.new("RGBA", dims, "white")
        sheet.paste(mark, (0, 0), mask=mark)
        fg = TF.to_tensor(sheet.convert("RGB"))
        if self.config.opaque:
            bg = torch.zeros(1, *dims)
        else:
            clear = Image.new("RGBA", dims, (0,) * 4)
            clear.paste(mark, (0, 0), mask=mark)
            bg = (TF.to_tensor(clear)[3:] == 0).float()
        self._set_patch(fg, bg)


class RandomNoisePatch(_PatchTrigger):
    def __init__(self, config, **kwargs):
        super().__init__()
        self._setup(config, kwargs)
        self.reset()

    def reset(self):
        size = self.config.size
        fg = torch.rand(3, size, size)              # global CPU RNG, as the reference
        self._set_patch(fg, torch.zeros(1, size, size))


class RandomBitMask(nn.Module):
    def __init__(self, config, **kwargs):
        super().__init__()
        self.n = config.n_bit
        self.c = config.constant
        self.z_dim = config.z_dim
        self.reset()

    def forward(self, z):
        with torch.no_grad():
            return ops.bitmask_scatter(_stage(z, self._mask.device), self._mask, self.c)

    def reset(self):
        mask = torch.randperm(self.z_dim)[:self.n].unsqueeze(0)
        if hasattr(self, "_mask"):
            mask = mask.to(self._mask.device)
        self.register_buffer("_mask", mask)

    @property
    def mask(self):
        return self._mask

    @mask.setter
    def mask(self, mask):
        self._mask = mask


class TransformDist(nn.Module):
    def __init__(self, config, **kwargs):
        super().__init__()
        self.register_buffer("_anchor", torch.zeros(()), persistent=False)   # carries the module's device

    def forward(self, z):
        with torch.no_grad():
            return ops.transform_dist(_stage(z, self._anchor.device))

    def reset(self):
        pass


class TransformVar(nn.Module):
    def __init__(self, config, **kwargs):
        super().__init__()
        self.register_buffer("w", torch.ones(1, 128))
        self.register_buffer("a", torch.ones(1, 128))
        self.reset()

    def forward(self, z):
        with torch.no_grad():
            return ops.transform_var(_stage(z, self.w.device), self.a, self.w)

    def reset(self):
        dev = self.w.device
        self.w = torch.exp(torch.randn(1, 128).abs()).to(dev)
        self.a = (torch.rand(1, 128) < 0.25).float().to(dev)
