"""Parameter holders with the reference's ``state_dict`` key names.

The reference builds its hot path from ``torch.nn`` modules; what the rest of the reference depends on
is only the *names and shapes* of their parameters (``engine/optimizer.py:26-33`` LR groups,
``utils/checkpoint.py:122-201`` checkpoint remap / strict load).  These holders carry
fp32 master parameters under exactly those names and have no ``forward``: the arithmetic is done by
the sm_100a kernels behind ``stcat_b200.ops``.
"""
from __future__ import annotations

import math

import torch
from torch import nn


class LinearP(nn.Module):
    """weight [out, in], bias [out]  (nn.Linear's names / default init)."""

    def __init__(self, n_in: int, n_out: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(n_out, n_in))
        self.bias = nn.Parameter(torch.empty(n_out))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        bound = 1.0 / math.sqrt(n_in)
        nn.init.uniform_(self.bias, -bound, bound)

    def extra_repr(self):
        return f"in={self.weight.shape[1]}, out={self.weight.shape[0]}"


class NormP(nn.Module):
    """LayerNorm affine: weight [d] = 1, bias [d] = 0; eps 1e-5."""

    def __init__(self, d: int, eps: float = 1e-5):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))
        self.bias = nn.Parameter(torch.zeros(d))
        self.eps = eps


class MHAP(nn.Module):
    """nn.MultiheadAttention's parameters: packed in_proj_weight [3d, d], in_proj_bias [3d] (zeros),
    out_proj.{weight [d, d], bias [d] (zeros)}."""

    def __init__(self, d: int, nhead: int):
        super().__init__()
        self.embed_dim, self.num_heads = d, nhead
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d, d))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d))
        self.out_proj = LinearP(d, d)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.zeros_(self.out_proj.bias)


class OutProjOnly(nn.Module):
    """The reference's custom MultiheadAttention (attention.py:86-113) has no in-projections; its only
    parameters are out_proj.{weight [vdim, vdim], bias (zeros)}."""

    def __init__(self, vdim: int):
        super().__init__()
        self.out_proj = LinearP(vdim, vdim)
        nn.init.zeros_(self.out_proj.bias)


class TokenP(nn.Module):
    """nn.Embedding(1, d)'s parameter: weight [1, d] ~ N(0, 1)."""

    def __init__(self, d: int):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(1, d))


class MLPP(nn.Module):
    """Parameters of the reference MLP (net_utils.py:7-26): layers.{i}.{weight,bias}."""

    def __init__(self, n_in: int, hidden: int, n_out: int, num_layers: int, dropout: float = 0.0):
        super().__init__()
        dims = [n_in] + [hidden] * (num_layers - 1) + [n_out]
        self.num_layers = num_layers
        self.dropout_p = dropout
        self.layers = nn.ModuleList(LinearP(dims[i], dims[i + 1]) for i in range(num_layers))


class SineTable(nn.Module):
    """SeqEmbeddingSine (position_encoding.py:21-37): registered buffer ``te`` [max_len, 1, d]."""

    def __init__(self, max_len: int, d: int):
        super().__init__()
        from .posenc import seq_sine_table

        self.register_buffer("te", seq_sine_table(max_len, d))

    def rows(self, ln: int) -> torch.Tensor:
        return self.te[:ln, 0, :]


class LearnedTable(nn.Module):
    """SeqEmbeddingLearned (position_encoding.py:7-18): ``embed.weight`` [num_pos, d] ~ N(0, 1)."""

    def __init__(self, num_pos: int, d: int):
        super().__init__()
        self.embed = TokenP(d)
        self.embed.weight = nn.Parameter(torch.randn(num_pos, d))

    def rows(self, ln: int) -> torch.Tensor:
        return self.embed.weight[:ln]


def xavier_reset(module: nn.Module):
    """The reference's ``_reset_parameters`` (modal_encoder.py:35-38, query_decoder.py:78-81)."""
    for p in module.parameters():
        if p.dim() > 1:
            nn.init.xavier_uniform_(p)
