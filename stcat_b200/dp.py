"""Data-parallel plumbing of the hot path: one clip per GPU, one process per GPU (the reference's only regime,
datasets/build.py:151,157), gradients summed across ranks once per step.

The reference wraps the model in ``DistributedDataParallel(find_unused_parameters=True)`` (train_net.py:31-36):
bucketed NCCL all-reduces plus a per-iteration unused-parameter bitmap all-reduce, needed because
``ground_encoder.fusion.*`` and the six ``ca_qtime_proj.*`` never receive a gradient.  Here
every hot-path gradient is a view of ONE contiguous fp32 buffer: the wgrad kernels accumulate straight into it
(``ops.set_grad_fusion``), one memset clears it, ONE all-reduce per step exchanges it, and parameters the
forward never touches simply keep their zero slice (no unused-parameter search).
"""
from __future__ import annotations

import torch


class FlatGrads:
    """All gradients of ``model`` as views of one contiguous fp32 buffer."""

    def __init__(self, model):
        seen, params = set(), []
        for p in model.parameters():
            if id(p) not in seen:
                seen.add(id(p))
                params.append(p)
        n = sum(p.numel() for p in params)
        self.buf = torch.zeros(n, dtype=torch.float32, device=params[0].device)
        self.params = params
        o = 0
        for p in params:
            p.grad = self.buf[o:o + p.numel()].view_as(p)
            o += p.numel()
        self.numel = n

    def zero(self):
        self.buf.zero_()

    def all_reduce(self, average: bool = False):
        """Sum (or mean) of the flat buffer over the default process group; no-op without one."""
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        dist.all_reduce(self.buf)
        if average:
            self.buf.div_(dist.get_world_size())
