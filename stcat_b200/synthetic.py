"""Deterministic synthetic weights / inputs for the hot path.

There is no network for checkpoints or datasets, so benchmarks, golden fixtures and parity tests all
use the same recipe: every parameter is filled from a generator seeded by ``crc32(name) ^ seed`` so
that the reference model (in ``oracle/make_golden.py``), the oracle and the CUDA modules get
*identical* weights from the key name alone -- no 158 MB state_dict has to be committed.

Distributions follow the reference's initialisation in spirit (xavier-uniform for every dim>1
parameter, modal_encoder.py:35-38 / query_decoder.py:78-81) but biases and LayerNorm affines are made
non-trivial on purpose so that parity tests exercise them.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, List, Sequence

import torch


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def fill_param(name: str, shape: Sequence[int], seed: int = 0) -> torch.Tensor:
    """Value of parameter ``name`` (fp32, CPU)."""
    shape = tuple(shape)
    g = _gen(name, seed)
    u = torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1
    if len(shape) > 1:
        fan_out, fan_in = shape[0], shape[1]
        rf = 1
        for s in shape[2:]:
            rf *= s
        a = math.sqrt(6.0 / ((fan_in + fan_out) * rf))
        return u * a
    if name.endswith("weight") and ("norm" in name):
        return 1.0 + 0.1 * u
    return 0.05 * u


def fill_state_dict(sd: Dict[str, torch.Tensor], seed: int = 0, skip_suffixes=(".te",)) -> Dict[str, torch.Tensor]:
    """New dict with every floating-point entry of ``sd`` replaced by the recipe (buffers such as the
    sine tables ``*.te`` are kept)."""
    out = {}
    for k, v in sd.items():
        if any(k.endswith(s) for s in skip_suffixes) or not torch.is_floating_point(v):
            out[k] = v.detach().clone()
        else:
            out[k] = fill_param(k, v.shape, seed).to(v.dtype)
    return out


def make_inputs(durations: Sequence[int], H: int, W: int, L: int, d: int = 256, seed: int = 0,
                ragged: bool = False) -> dict:
    """Inputs at the hot-path seam (what ``input_proj`` / the text encoder hand to ``ground_encoder``).

    vis_features ~ N(0,1) [n,d,H,W]; vis_pos = the real image sine embedding of ``vis_mask``;
    text_memory ~ N(0,1) [L,b,d]; masks all-False unless ``ragged`` (then each video after the first
    loses trailing feature-map columns and every video has a different number of padded text tokens).
    """
    from . import posenc

    durations = list(durations)
    b, n = len(durations), sum(durations)
    g = _gen(f"inputs/{durations}/{H}x{W}/{L}", seed)
    vis = torch.randn(n, d, H, W, generator=g)
    txt = torch.randn(L, b, d, generator=g)
    vis_mask = torch.zeros(n, H, W, dtype=torch.bool)
    text_mask = torch.zeros(b, L, dtype=torch.bool)
    if ragged:
        s = 0
        for i, dur in enumerate(durations):
            if i > 0:
                cut = max(1, W - (i % 3) - 1)
                vis_mask[s:s + dur, :, cut:] = True
            pad = (2 * i + 1) % max(1, L - 1)
            if pad:
                text_mask[i, L - pad:] = True
            s += dur
        vis = vis * (~vis_mask)[:, None].float()
    vis_pos = posenc.image_sine_pos(vis_mask, d // 2)
    return {
        "vis_features": vis,
        "vis_mask": vis_mask,
        "vis_pos": vis_pos,
        "text_mask": text_mask,
        "text_memory": txt,
        "durations": durations,
    }


def make_targets(durations: Sequence[int], seed: int = 0) -> dict:
    """Targets for the fwd+bwd step: actioness = 1 on [dur/4, 3dur/4) of every video;
    boxes cxcywh with centre in [0.2,0.8], size in [0.1,0.4] (valid for the GIoU asserts)."""
    durations = list(durations)
    b, t = len(durations), max(durations)
    g = _gen(f"targets/{durations}", seed)
    act = torch.zeros(b, t)
    boxes: List[torch.Tensor] = []
    for i, dur in enumerate(durations):
        s, e = dur // 4, max(dur // 4 + 1, (3 * dur) // 4)
        act[i, s:e] = 1
        k = e - s
        c = 0.2 + 0.6 * torch.rand(k, 2, generator=g)
        wh = 0.1 + 0.3 * torch.rand(k, 2, generator=g)
        boxes.append(torch.cat([c, wh], 1))
    return {"actioness": act, "boxes": torch.cat(boxes, 0)}
