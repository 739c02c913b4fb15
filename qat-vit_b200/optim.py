"""Fused gradient clipping + AdamW over flat arenas (ref/src/training/qat_trainer.py:271-278,360-361).

``FusedClipAdamW(params, grad_arena, ...)`` re-homes every parameter's storage into ONE flat fp32 arena (the Parameters
themselves -- identity, names, state_dict keys -- are unchanged: ``p.data`` becomes a view of the arena), keeps the two
moment buffers as flat arenas, and performs ``clip_grad_norm_(params, max_norm)`` + ``AdamW.step()`` as two kernel launches
(qv_clip_adamw) with torch.optim.AdamW's arithmetic.  ``grad_arena`` is the engine's flat gradient buffer (parameter order).
"""
from __future__ import annotations

import ctypes
from typing import Iterable, Tuple

import torch

from . import _lib
from ._lib import check


class FusedClipAdamW:
    def __init__(self, params: Iterable[torch.nn.Parameter], grad_arena: torch.Tensor, lr: float, betas: Tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 1e-2, max_norm: float = 1.0):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        if grad_arena.numel() != total or grad_arena.dtype != torch.float32 or not grad_arena.is_cuda:
            raise ValueError("grad_arena must be the flat CUDA fp32 gradient buffer of exactly these parameters (no CPU fallback)")
        dev = grad_arena.device
        self.grad_arena = grad_arena
        self.flat = torch.empty(total, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("parameters must be fp32 on the gradient arena's device")
            n = p.numel()
            self.flat[off:off + n].copy_(p.detach().reshape(-1))
            p.data = self.flat[off:off + n].view_as(p)          # same Parameter object, storage now inside the arena
            off += n
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        self.partials = torch.empty(sms * 4, dtype=torch.float32, device=dev)
        self.total_norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.lr, self.betas, self.eps, self.weight_decay, self.max_norm = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay), float(max_norm)
        self.steps = 0

    def step(self, grad_scale: float = 1.0, write_back_grad: bool = False) -> torch.Tensor:
        """clip + AdamW; grad_scale folds the 1/world of a summed (all-reduced) gradient.  Returns the total-norm tensor."""
        self.steps += 1
        P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
        check(_lib.lib().qv_clip_adamw(P(self.flat), P(self.grad_arena), P(self.exp_avg), P(self.exp_avg_sq), self.flat.numel(),
                                       P(self.partials), self.partials.numel(), float(grad_scale), self.max_norm, self.lr,
                                       self.betas[0], self.betas[1], self.eps, self.weight_decay, self.steps, P(self.total_norm),
                                       int(bool(write_back_grad)), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
              "clip_adamw")
        return self.total_norm

    def zero_grad(self, set_to_none: bool = True) -> None:
        """No-op: the engine overwrites every gradient each step (kept for API symmetry with torch optimizers)."""
