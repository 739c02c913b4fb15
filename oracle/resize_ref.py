"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the reference's input transform:

    transforms.Compose([Resize(224, interpolation=BICUBIC), ToTensor(), Normalize(mean, std)])      ref qat_trainer.py:210-216

applied to the 32x32x3 uint8 CIFAR-10 images `datasets.CIFAR10` hands out as PIL images (ref :218-219).  The arithmetic lives
in third-party code absent from /root/reference: Pillow (unpinned, pulled in by torchvision; installed 12.2.0) and torchvision
(unpinned, ref requirements.txt; installed 0.26.0).  Restated algorithms:

  * Pillow src/libImaging/Resample.c: `precompute_coeffs` (double: centre, support = 2.0 * max(scale, 1), window bounds by
    C-truncation, bicubic a = -0.5, weights normalised by their sum), `normalize_coeffs_8bpc` (fixed point, PRECISION_BITS = 22,
    round half away from zero by +-0.5 and truncation), `ImagingResampleHorizontal_8bpc` then `ImagingResampleVertical_8bpc`
    (int32 accumulate from 1 << 21, arithmetic shift, clip to 0..255) -- two passes with a uint8 intermediate image;
  * torchvision ToTensor: uint8 -> float32, true division by 255;  Normalize: (x - mean[c]) / std[c] in float32.

Pinned bit-for-bit against the LIVE Pillow + torchvision pipeline (tests/test_oracle.py, golden fixture tests/golden/resize.npz
written by oracle/gen_golden.py).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def _bicubic(x: float, a: float = -0.5) -> float:
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int, support: float = 2.0):
    """-> (bounds int32 [out, 2] = (first input index, tap count), coefficients int32 [out, ksize] in 22-bit fixed point)."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = support * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = int(center - support + 0.5)            # C (int) cast: truncation toward zero
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_axis(img: np.ndarray, bounds: np.ndarray, kk: np.ndarray, axis: int) -> np.ndarray:
    """One 8-bit pass along `axis` of an [H, W, C] uint8 image."""
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((bounds.shape[0],) + src.shape[1:], np.uint8)
    for xx in range(bounds.shape[0]):
        xmin, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        ss = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(n):
            ss += src[xmin + x] * int(kk[xx, x])
        out[xx] = np.clip(ss >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bicubic_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """PIL Image.resize((out_w, out_h), BICUBIC) of an [H, W, C] uint8 image: horizontal pass first, then vertical."""
    bh, kh = precompute_coeffs(img.shape[1], out_w)
    bv, kv = precompute_coeffs(img.shape[0], out_h)
    tmp = _resample_axis(img, bh, kh, axis=1)
    return _resample_axis(tmp, bv, kv, axis=0)


def transform(img_u8: np.ndarray, size: int = 224, mean=IMAGENET_MEAN, std=IMAGENET_STD) -> np.ndarray:
    """The whole reference transform for one [H, W, 3] uint8 image -> float32 [3, size', size''] (smaller edge -> `size`)."""
    h, w = img_u8.shape[:2]
    if w <= h:
        ow, oh = size, int(size * h / w)
    else:
        oh, ow = size, int(size * w / h)
    r = resize_bicubic_u8(img_u8, oh, ow)
    t = r.transpose(2, 0, 1).astype(np.float32) / np.float32(255)                     # ToTensor
    m = np.asarray(mean, np.float32)[:, None, None]
    s = np.asarray(std, np.float32)[:, None, None]
    return np.ascontiguousarray(((t - m) / s).astype(np.float32))                     # Normalize
