"""oracle/shims/pdqhash -- TEST INFRASTRUCTURE ONLY.

Stand-in for the third-party ``pdqhash==0.2.2`` Cython package (absent from this image; pinned
at /root/reference/requirements.txt:5) so that /root/reference/tools/phash_pvalue.py imports and
runs on CPU.  ``compute(rgb_uint8_hwc) -> (bits[256], quality)`` is served by the C restatement
in ``oracle/pdq_oracle.c`` (see its header: PARITY UNPINNED).  ``quality`` is unused by the
reference (tools/phash_pvalue.py:13-14) and returned as 100.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE = os.path.normpath(os.path.join(_HERE, "..", ".."))
_SO = os.path.join(_ORACLE, "_build", "libpdq_oracle.so")


def _load():
    src = os.path.join(_ORACLE, "pdq_oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _ORACLE, "-s"])
    lib = ctypes.CDLL(_SO)
    lib.pdq_oracle_hash_rgb.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.c_void_p]
    lib.pdq_oracle_hash_rgb.restype = None
    lib.pdq_oracle_hash_batch.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_void_p]
    lib.pdq_oracle_hash_batch.restype = None
    lib.pdq_oracle_dct_matrix.argtypes = [ctypes.c_void_p]
    lib.pdq_oracle_dct_matrix.restype = None
    return lib


_lib = _load()


def compute(image):
    img = np.ascontiguousarray(image, dtype=np.uint8)
    assert img.ndim == 3 and img.shape[2] == 3, "expected an HxWx3 uint8 RGB array"
    bits = np.zeros(256, dtype=np.uint8)
    _lib.pdq_oracle_hash_rgb(img.ctypes.data, img.shape[0], img.shape[1], bits.ctypes.data, None)
    return bits, 100


def compute_with_coeffs(image):
    img = np.ascontiguousarray(image, dtype=np.uint8)
    bits = np.zeros(256, dtype=np.uint8)
    coeffs = np.zeros(256, dtype=np.float32)
    _lib.pdq_oracle_hash_rgb(img.ctypes.data, img.shape[0], img.shape[1], bits.ctypes.data,
                             coeffs.ctypes.data)
    return bits, coeffs


def compute_batch(images):
    """images: (B, H, W, 3) uint8 -> (B, 256) uint8 bits."""
    imgs = np.ascontiguousarray(images, dtype=np.uint8)
    bits = np.zeros((imgs.shape[0], 256), dtype=np.uint8)
    _lib.pdq_oracle_hash_batch(imgs.ctypes.data, imgs.shape[0], imgs.shape[1], imgs.shape[2],
                               bits.ctypes.data)
    return bits


def dct_matrix():
    d = np.zeros((16, 64), dtype=np.float32)
    _lib.pdq_oracle_dct_matrix(d.ctypes.data)
    return d
