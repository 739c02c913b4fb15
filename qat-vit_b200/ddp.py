"""Data-parallel glue for the fused step: ONE exchange per iteration.

The reference wraps the prepared student in ``DistributedDataParallel`` (ref/src/training/qat_trainer.py:310-313), which
(a) averages gradients with bucketed NCCL all-reduces during backward and (b) broadcasts every buffer -- i.e. the
observer state -- from rank 0 at each forward entry (torch/nn/parallel/distributed.py:1590-1591).

Here the engine writes all gradients into one flat fp32 arena; ``GradSync`` all-reduces it (SUM; the 1/world is folded
into the clip pass) and piggy-backs rank 0's activation-observer running min/max in the tail of the same buffer:
rank 0 contributes its values, every other rank contributes zeros, so after the SUM every rank holds rank 0's state
exactly -- the DDP buffer-broadcast semantics without a second collective.  (scale / zero_point are a deterministic
function of min/max and are recomputed on the next forward; weight observers see identical weights on every rank.)
Works with any torch.distributed backend (nccl on the B200 box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, n_grad: int, n_observers: int, device, dtype=torch.float32, bucket_bytes: int = 32 << 20):
        """One flat buffer [n_grad gradients | 2*n_observers running min/max].  Build the engine on ``grad_arena`` and
        then ``bind_observers`` the (min_val, max_val) buffers of the activation fake-quant modules."""
        self.n_grad = n_grad
        self.observers: List[Tuple[torch.Tensor, torch.Tensor]] = []
        self.n_tail = 2 * n_observers
        self.flat = torch.zeros(n_grad + self.n_tail, dtype=dtype, device=device)
        self.bucket_elems = max(1, bucket_bytes // 4)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.rehomed = False

    def bind_observers(self, observers: Sequence[Tuple[torch.Tensor, torch.Tensor]], rehome: Optional[bool] = None) -> None:
        """observers: the (min_val, max_val) buffers of the activation fake-quant modules, in a fixed order.
        rehome (default: on for CUDA): move the storage of every buffer INTO the tail of the flat buffer (the module keeps the
        same tensor objects -- names, state_dict keys and values unchanged; ``t.data`` becomes a view of the tail), so that
        packing rank 0's state and adopting it after the all-reduce cost no launch at all: the kernels that update the
        running ranges write straight into the exchange buffer.  Do not deep-copy / ``.to()`` the model afterwards."""
        observers = list(observers)
        if 2 * len(observers) != self.n_tail:
            raise ValueError(f"expected {self.n_tail // 2} observers, got {len(observers)}")
        self.observers = observers
        if rehome is None:
            rehome = self.flat.is_cuda
        self.rehomed = bool(rehome) and len(observers) > 0
        if self.rehomed:
            t = self.tail
            with torch.no_grad():
                for i, (mn, mx) in enumerate(observers):
                    for j, buf in enumerate((mn, mx)):
                        if buf.numel() != 1 or buf.dtype != t.dtype:
                            raise ValueError("activation observers are per-tensor fp32 scalars")
                        slot = t[2 * i + j:2 * i + j + 1]
                        slot.copy_(buf.reshape(1))
                        buf.data = slot.view(buf.shape)

    @property
    def grad_arena(self) -> torch.Tensor:
        return self.flat[:self.n_grad]

    @property
    def tail(self) -> torch.Tensor:
        return self.flat[self.n_grad:]

    def pack_observers(self) -> None:
        """Rank 0 contributes its running min / max, every other rank zeros: after the SUM everyone holds rank 0's state."""
        if self.n_tail == 0:
            return
        if len(self.observers) * 2 != self.n_tail:
            raise RuntimeError("GradSync.bind_observers() has not been called")
        if self.rehomed:
            if self.rank != 0:          # the local ranges are dead after the forward: clear them in place (one memset)
                if self.flat.is_cuda:
                    from . import ops
                    ops.zero_(self.tail)
                else:
                    self.tail.zero_()
            return
        if self.rank == 0:
            vals = [t.reshape(1) for pair in self.observers for t in pair]
            torch.cat(vals, out=self.tail)
        else:
            self.tail.zero_()

    def unpack_observers(self) -> None:
        if self.n_tail == 0 or self.world == 1 or self.rehomed:
            return
        t = self.tail
        for i, (mn, mx) in enumerate(self.observers):
            mn.copy_(t[2 * i].reshape(mn.shape))
            mx.copy_(t[2 * i + 1].reshape(mx.shape))

    def buckets(self) -> List[Tuple[int, int]]:
        """[start, end) element ranges, last bucket first (gradients are produced head -> patch-embed)."""
        total = self.flat.numel()
        out, end = [], total
        while end > 0:
            start = max(0, end - self.bucket_elems)
            out.append((start, end))
            end = start
        return out

    def all_reduce(self, async_op: bool = False):
        """SUM over ranks of gradients + observer tail.  Returns the list of work handles when async."""
        if self.world == 1:
            return []
        self.pack_observers()
        works = []
        for (s, e) in self.buckets():
            w = dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.SUM, async_op=async_op)
            if async_op:
                works.append(w)
        if not async_op:
            self.unpack_observers()
        return works

    def finish(self, works) -> None:
        for w in works:
            w.wait()
        self.unpack_observers()

    # ---- overlapped mode: the engine reports, layer by layer, how much of the arena is final -------------------------
    def begin_step(self, min_bucket_bytes: Optional[int] = None) -> None:
        """Call before backward.  Gradients are produced from the END of the arena (head, last block, ...) towards its start
        (embeddings), like DDP's reverse-order buckets (torch/nn/parallel/distributed.py:831-833): ``grads_final_from(lo)``
        all-reduces the newly completed suffix asynchronously so the transfer overlaps the remaining backward GEMMs."""
        if min_bucket_bytes is None:
            import os
            min_bucket_bytes = int(float(os.environ.get("QV_DDP_BUCKET_MB", "25")) * (1 << 20))
        self._hi = self.flat.numel()
        self._works = []
        self._min_bucket = max(1, min_bucket_bytes // 4)
        if self.world > 1:
            self.pack_observers()          # rank 0's running min/max are final after the forward; they ride in the first bucket

    def grads_final_from(self, lo: int) -> None:
        """Every gradient with arena offset >= lo is final.  lo == 0 flushes what is left."""
        if self.world == 1:
            return
        if lo > 0 and self._hi - lo < self._min_bucket:
            return
        if self._hi > lo:
            self._works.append(dist.all_reduce(self.flat[lo:self._hi], op=dist.ReduceOp.SUM, async_op=True))
            self._hi = lo

    def end_step(self) -> None:
        """Wait for the outstanding all-reduces (stream-ordered for NCCL) and adopt rank 0's observer state."""
        if self.world == 1:
            return
        self.grads_final_from(0)
        self.finish(self._works)
        self._works = []
