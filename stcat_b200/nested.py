"""``NestedTensor``: the boundary type of ``STCATNet.forward`` / ``build_encoder(cfg).forward``.

Mirrors the reference's ``utils.misc.NestedTensor`` (utils/misc.py:41-97): frames of all videos of
the batch concatenated along dim 0, a padding mask per frame and the per-video frame counts.  The
modules of this package only call ``decompose()``, so the reference's own class works as well.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch


class NestedTensor:
    def __init__(self, tensors: torch.Tensor, mask: Optional[torch.Tensor], durations: Sequence[int]):
        self.tensors = tensors
        self.mask = mask
        self.durations = list(durations)

    def to(self, *args, **kwargs) -> "NestedTensor":
        m = self.mask.to(*args, **kwargs) if self.mask is not None else None
        return type(self)(self.tensors.to(*args, **kwargs), m, self.durations)

    def decompose(self):
        return self.tensors, self.mask, self.durations

    def subsample(self, stride: int, start_idx: int = 0) -> "NestedTensor":
        """every ``stride``-th frame of every video (engine/evaluate.py:97-104 even/odd passes)."""
        ts = [v[start_idx::stride] for v in torch.split(self.tensors, self.durations, dim=0)]
        ms = [m[start_idx::stride] for m in torch.split(self.mask, self.durations, dim=0)]
        return NestedTensor(torch.cat(ts, 0), torch.cat(ms, 0), [x.shape[0] for x in ts])

    @classmethod
    def from_tensor_list(cls, clips: List[torch.Tensor]) -> "NestedTensor":
        """clips: list of [T_i, C, H_i, W_i]; zero-pads to the largest H, W; mask True = padding."""
        assert clips[0].ndim == 4
        c = max(x.shape[1] for x in clips)
        h = max(x.shape[2] for x in clips)
        w = max(x.shape[3] for x in clips)
        durations = [x.shape[0] for x in clips]
        out = clips[0].new_zeros((sum(durations), c, h, w))
        mask = torch.ones((sum(durations), h, w), dtype=torch.bool, device=clips[0].device)
        s = 0
        for x in clips:
            out[s:s + x.shape[0], :x.shape[1], :x.shape[2], :x.shape[3]].copy_(x)
            mask[s:s + x.shape[0], :x.shape[2], :x.shape[3]] = False
            s += x.shape[0]
        return cls(out, mask, durations)

    def __repr__(self):
        return f"NestedTensor(tensors={tuple(self.tensors.shape)}, durations={self.durations})"
