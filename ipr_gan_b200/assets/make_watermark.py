"""Synthetic RGBA watermark recipe (the reference's ./data/watermarks/*.png are not in its repo).

64x64 RGBA: a coloured ring + diagonal stripes on a transparent ground, with a partially
transparent rim so that both the opaque (bg = 0) and the alpha-masked (bg = alpha == 0)
variants of PasteWatermark are exercised.  Deterministic; run to regenerate watermark_a.png.
"""
import os

import numpy as np
from PIL import Image


def make(size=64):
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float64)
    c = (size - 1) / 2.0
    rad = np.hypot(yy - c, xx - c)
    img = np.zeros((size, size, 4), dtype=np.uint8)
    ring = (rad > size * 0.22) & (rad < size * 0.42)
    stripes = ((xx + yy).astype(np.int64) // 6) % 2 == 0
    img[..., 0] = np.where(ring, 220, np.where(stripes, 30, 90))
    img[..., 1] = np.where(ring, 40, np.where(stripes, 160, 200))
    img[..., 2] = np.where(ring & stripes, 60, np.where(stripes, 210, 20))
    alpha = np.where(rad < size * 0.47, 255, 0)
    rim = (rad >= size * 0.42) & (rad < size * 0.47)
    alpha = np.where(rim, 128, alpha)
    img[..., 3] = alpha
    return Image.fromarray(img, "RGBA")


if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "watermark_a.png")
    make().save(out)
    print(out)
