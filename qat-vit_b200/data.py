"""GPU input transform: drop-in for the reference's per-image CPU transform (ref/src/training/qat_trainer.py:210-216)

    transform = transforms.Compose([transforms.Resize(224, interpolation=BICUBIC), transforms.ToTensor(),
                                    transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])

applied to a whole BATCH of raw uint8 HWC images (what ``datasets.CIFAR10(...).data`` holds: [N, 32, 32, 3]) that is already
on the GPU: 0.8 MB of uint8 cross PCIe per 256-image batch instead of 154 MB of fp32, and the 4-worker CPU DataLoader -- the
bottleneck at > 4 000 img/s -- is out of the loop.  The kernel (csrc/resize.cu, qv_resize_normalize_u8)
reproduces Pillow's 8-bit two-pass bicubic resample and torchvision's ToTensor / Normalize bit for bit.

This file holds the size-only part of Pillow's algorithm: the fixed-point tap tables (src/libImaging/Resample.c,
precompute_coeffs + normalize_coeffs_8bpc), computed once per input geometry in double precision.
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, Sequence, Tuple

import torch

from . import _lib
from ._lib import check

_PRECISION_BITS = 32 - 8 - 2      # Pillow: 8 bits of pixel, 2 bits of head-room for the negative bicubic lobes


def _bicubic(x: float, a: float = -0.5) -> float:
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def bicubic_taps(in_size: int, out_size: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(bounds int32 [out, 2] = (first input index, tap count), taps int32 [out, ksize]) of one resample pass."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = torch.zeros(out_size, 2, dtype=torch.int32)
    taps = torch.zeros(out_size, ksize, dtype=torch.int32)
    inv = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)            # int(): truncation toward zero, like the C cast
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * inv) for x in range(xmax)]
        total = 0.0
        for v in w:
            total += v
        if total != 0.0:
            w = [v / total for v in w]
        for x, v in enumerate(w):
            taps[xx, x] = int(-0.5 + v * (1 << _PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << _PRECISION_BITS))
        bounds[xx, 0], bounds[xx, 1] = xmin, xmax
    return bounds, taps


class GpuImageTransform:
    """``GpuImageTransform(224)(batch_u8)``: uint8 CUDA [B, H, W, C] -> float32 CUDA [B, C, H', W'] (smaller edge -> size, like
    transforms.Resize(int)), normalised with (mean, std)."""

    def __init__(self, size: int = 224, mean: Sequence[float] = (0.485, 0.456, 0.406), std: Sequence[float] = (0.229, 0.224, 0.225)):
        self.size = int(size)
        self.mean, self.std = tuple(float(m) for m in mean), tuple(float(s) for s in std)
        self._tables: Dict[tuple, tuple] = {}

    def output_hw(self, h: int, w: int) -> Tuple[int, int]:
        """torchvision.transforms.functional.resize with an int size: the smaller edge becomes `size`."""
        if w <= h:
            return int(self.size * h / w), self.size
        return self.size, int(self.size * w / h)

    def _get_tables(self, h: int, w: int, dev) -> tuple:
        key = (h, w, str(dev))
        if key not in self._tables:
            oh, ow = self.output_hw(h, w)
            bh, kh = bicubic_taps(w, ow)
            bv, kv = bicubic_taps(h, oh)
            ks = max(kh.shape[1], kv.shape[1])
            pad = lambda t: torch.nn.functional.pad(t, (0, ks - t.shape[1]))  # noqa: E731
            self._tables[key] = (oh, ow, ks, bh.to(dev), pad(kh).contiguous().to(dev), bv.to(dev), pad(kv).contiguous().to(dev),
                                 torch.tensor(self.mean, dtype=torch.float32, device=dev),
                                 torch.tensor(self.std, dtype=torch.float32, device=dev))
        return self._tables[key]

    def __call__(self, batch_u8: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        if not batch_u8.is_cuda or batch_u8.dtype != torch.uint8 or batch_u8.dim() != 4 or not batch_u8.is_contiguous():
            raise RuntimeError("qatvit_b200: GpuImageTransform takes a contiguous uint8 CUDA tensor [B, H, W, C] (no CPU fallback)")
        B, H, W, C = batch_u8.shape
        if C != len(self.mean):
            raise RuntimeError(f"qatvit_b200: {C} channels but {len(self.mean)} mean / std entries")
        oh, ow, ks, bh, kh, bv, kv, mean, std = self._get_tables(H, W, batch_u8.device)
        if out is None:
            out = torch.empty(B, C, oh, ow, dtype=torch.float32, device=batch_u8.device)
        elif tuple(out.shape) != (B, C, oh, ow) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != batch_u8.device:
            raise RuntimeError("qatvit_b200: out must be a contiguous float32 [B, C, H', W'] tensor on the input's device")
        P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
        check(_lib.lib().qv_resize_normalize_u8(P(batch_u8), B, H, W, C, oh, ow, P(bh), P(kh), P(bv), P(kv), ks, P(mean), P(std),
                                                P(out), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
              "resize_normalize_u8")
        return out
