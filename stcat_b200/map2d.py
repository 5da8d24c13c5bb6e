"""The 2-D temporal proposal head of the reference, ``models/map2d_head.py`` (orphaned there: nothing in ``STCATNet`` calls
it, but it is part of the named path -- SURVEY.md 8a-13): ``Gen2DMap`` (clip features -> [d, N, N] proposal map),
``TempPredictionHead`` with its two interaction variants (``TEMP_HEAD`` = 'attn': row / column attention over the map,
the reference default; 'conv': a stack of masked-weight k x k convolutions) and the 1 x 1 predictor.

Same constructor argument (cfg), same ``forward(x [layers, b, T, d]) -> scores [layers, b, N, N]`` contract (train mode:
logits; eval mode: ``sigmoid * mask2d``) and the same ``state_dict`` names / shapes as the reference module.  The reference
reads ``cfg.MODEL.TEMPFORMER``, a node ``config/defaults.py`` never defines (it only runs with
``cfg.MODEL.TEMPFORMER = cfg.MODEL.STCAT``); this module reads ``cfg.MODEL.TEMPFORMER`` when present and
``cfg.MODEL.STCAT`` otherwise.

The arithmetic goes through the C ABI: the pooling cascade is ``stcat_map2d_pool``; the attention variant is packed
in-projection GEMMs + ``stcat_attention_*`` (sequence = a map row or column, batch = maps x N) + LayerNorm / FFN blocks; the
convolutions are GEMMs over an im2col of the map (``F.unfold`` is layout glue, the contraction [pixels, d k^2] x
[d k^2, d] runs in ``stcat_linear_*``).  Two quirks of the reference are reproduced on purpose (the oracle pins them
against reference fixtures): the attention passes the *valid-cell* mask as ``key_padding_mask`` (True = ignore in torch),
and indexes it [batch, key], i.e. transposed with respect to the map.
"""
from __future__ import annotations

from typing import List

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .params import LinearP, MHAP, NormP


def map2d_masks(N: int, pooling_counts):
    """Gen2DMap.__init__ (map2d_head.py:11-37): valid-cell mask, the (i, j) lists of every super-diagonal and the pooling
    schedule as (kernel, stride) pairs."""
    mask2d = torch.zeros(N, N, dtype=torch.bool)
    mask2d[range(N), range(N)] = True
    stride, offset = 1, 0
    maskij = []
    for c in pooling_counts:
        for _ in range(c):
            offset += stride
            i, j = list(range(0, N - offset, stride)), list(range(offset, N, stride))
            mask2d[i, j] = True
            maskij.append((i, j))
        stride *= 2
    poolers = [(2, 1)] * pooling_counts[0]
    for c in pooling_counts[1:]:
        poolers += [(3, 2)] + [(2, 1)] * (c - 1)
    return mask2d, maskij, poolers


def _pool_cascade_torch(x, N, maskij, poolers):
    """The reference's cascade with torch ops (map2d_head.py:53-61); only used to differentiate the map w.r.t. the clip
    features (the forward kernel has no backward of its own -- the head is not trained anywhere in the reference)."""
    B, d, _ = x.shape
    m = x.new_zeros(B, d, N, N)
    m[:, :, range(N), range(N)] = x
    for (kk, ss), (i, j) in zip(poolers, maskij):
        x = F.max_pool1d(x, kk, ss)
        m[:, :, i, j] = x
    return m


class _Map2dPoolFn(torch.autograd.Function):
    """[B, N, d] clip features -> [B, d, N, N] map: cell (i, j) of the valid set holds max over frames i..j."""

    @staticmethod
    def forward(ctx, x, valid_u8, N, maskij, poolers):
        xd = x.detach().float().contiguous()
        B, _, d = xd.shape
        out = torch.empty(B, d, N, N, dtype=torch.float32, device=x.device)
        ops.get_backend().map2d_pool(xd, valid_u8, out)
        ctx.save_for_backward(xd)
        ctx.meta = (N, maskij, poolers)
        return out

    @staticmethod
    def backward(ctx, g):
        (xd,) = ctx.saved_tensors
        N, maskij, poolers = ctx.meta
        with torch.enable_grad():
            xr = xd.detach().requires_grad_(True)
            m = _pool_cascade_torch(xr.permute(0, 2, 1), N, maskij, poolers)
            (gx,) = torch.autograd.grad(m, xr, g)
        return gx, None, None, None, None


class Gen2DMap(nn.Module):
    """map2d_head.py:9-62."""

    def __init__(self, cfg_node):
        super().__init__()
        self.map_size = int(cfg_node.MAX_MAP_SIZE)
        mask2d, self.maskij, self.poolers = map2d_masks(self.map_size, list(cfg_node.POOLING_COUNTS))
        self.register_buffer("mask2d", mask2d, persistent=False)
        self.register_buffer("valid_u8", mask2d.to(torch.uint8).contiguous(), persistent=False)

    def forward(self, x):
        """x [B, T, d] -> map2d [B, d, N, N]"""
        N = self.map_size
        if x.shape[1] != N:  # [B, d, T] -> N frames (:48-51); T == N makes both poolings the identity
            xt = x.permute(0, 2, 1)
            if xt.shape[-1] > N:
                xt = F.adaptive_avg_pool1d(xt, N)
            x = F.adaptive_max_pool1d(xt, N).permute(0, 2, 1)
        return _Map2dPoolFn.apply(x, self.valid_u8, N, self.maskij, self.poolers)


class _AttnLayerP(nn.Module):
    """Parameters of map2d_head.TransformerEncoderLayer (:151-171)."""

    def __init__(self, d, nhead, ffn):
        super().__init__()
        self.self_attn_row = MHAP(d, nhead)
        self.self_attn_col = MHAP(d, nhead)
        self.linear1 = LinearP(d, ffn)
        self.linear2 = LinearP(ffn, d)
        self.norm1, self.norm2 = NormP(d), NormP(d)


class _AttnEncoder(nn.Module):
    def __init__(self, d, nhead, ffn, num_layers):
        super().__init__()
        self.layers = nn.ModuleList(_AttnLayerP(d, nhead, ffn) for _ in range(num_layers))


class _ConvP(nn.Module):
    """nn.Conv2d's parameter names: weight [out, in, k, k], bias [out]."""

    def __init__(self, d_in, d_out, k):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(d_out, d_in, k, k))
        self.bias = nn.Parameter(torch.empty(d_out))
        nn.init.xavier_uniform_(self.weight)
        bound = 1.0 / (d_in * k * k) ** 0.5
        nn.init.uniform_(self.bias, -bound, bound)


class _ConvEncoder(nn.Module):
    def __init__(self, d, k, num_layers):
        super().__init__()
        self.convs = nn.ModuleList(_ConvP(d, d, k) for _ in range(num_layers))


def _mask2weight(mask2d, k, padding):
    """map2d_head.py:221-226: 1 / (number of valid cells under the kernel window), 0 where there is none."""
    w = torch.conv2d(mask2d[None, None].float(), torch.ones(1, 1, k, k), padding=padding)[0, 0]
    w[w > 0] = 1 / w[w > 0]
    return w


class TempPredictionHead(nn.Module):
    """map2d_head.py:65-127."""

    def __init__(self, cfg):
        super().__init__()
        node = cfg.MODEL.TEMPFORMER if hasattr(cfg.MODEL, "TEMPFORMER") else cfg.MODEL.STCAT
        d, self.nhead = int(node.HIDDEN), int(node.HEADS)
        if d != 256 or self.nhead != 8:
            raise NotImplementedError("the sm_100a kernels are built for HIDDEN=256, HEADS=8 (head dim 32)")
        self.d = d
        self.temp_head = node.TEMP_HEAD
        self.dropout_p = float(node.DROPOUT)
        self.map_maker = Gen2DMap(node)
        N = self.map_maker.map_size
        if self.temp_head == "attn":
            self.encoder = _AttnEncoder(d, self.nhead, int(node.FFN_DIM), int(node.TEMP_PRED_LAYERS))
            m = self.map_maker.mask2d
            # key_padding_mask exactly as the reference hands it over (:181-193): row attention -> mask2d[batch, key], column
            # attention -> mask2d^T[batch, key]; nonzero = ignored
            self.register_buffer("row_key_mask", m.to(torch.uint8).contiguous(), persistent=False)
            self.register_buffer("col_key_mask", m.t().to(torch.uint8).contiguous(), persistent=False)
        else:
            k, nconv = int(node.KERNAL_SIZE), int(node.CONV_LAYERS)
            self.kernel, self.first_padding = k, (k - 1) * nconv // 2
            self.encoder = _ConvEncoder(d, k, nconv)
            ws: List[torch.Tensor] = [_mask2weight(self.map_maker.mask2d, k, self.first_padding)]  # :232-245
            for _ in range(nconv - 1):
                ws.append(_mask2weight(ws[-1] > 0, k, 0))
            for i, w in enumerate(ws):
                self.register_buffer(f"conv_weight_{i}", w.contiguous(), persistent=False)
        self.predictor = _ConvP(d, 1, 1)
        assert N == self.map_maker.mask2d.shape[0]

    # -- interaction variants ------------------------------------------------------------------------------------------
    def _mha(self, a: MHAP, x, x_op, B, L, key_mask, p):
        """nn.MultiheadAttention with q = k = v = x (batch-major rows [B*L, d]): one packed in-projection GEMM, the
        attention core on column slices of its output, out-projection."""
        d = self.d
        qkv = ops.linear(x, a.in_proj_weight, a.in_proj_bias, out_bf16=True, x_op=x_op)  # [B*L, 3d]
        o, _ = ops.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], B, self.nhead, L, L, float(d // self.nhead) ** -0.5,
                             key_mask=key_mask, drop_p=p)
        return ops.linear(o, a.out_proj.weight, a.out_proj.bias)

    def _attn(self, maps):
        """maps [n, d, N, N] -> [n, N, N, d] after the row / column attention layers (:173-205)."""
        n, d, N, _ = maps.shape
        p = self.dropout_p if self.training else 0.0
        src = maps.permute(0, 2, 3, 1).reshape(n * N * N, d).contiguous()  # rows (map, i, j)
        row_mask = self.row_key_mask.repeat(n, 1)  # [n*N, N]: batch = (map, j), keys = i
        col_mask = self.col_key_mask.repeat(n, 1)  # batch = (map, i), keys = j
        tr = lambda t: t.view(n, N, N, d).transpose(1, 2).reshape(n * N * N, d)  # (map, i, j) <-> (map, j, i)
        for layer in self.encoder.layers:
            xr = tr(src)                                             # sequences run along i, one per (map, j)
            a = self._mha(layer.self_attn_row, xr, None, n * N, N, row_mask, p)   # rows (map, j, i)
            # the reference's permute(1, 0, 2) re-reads that output with the two axes swapped in role (:188): the column
            # attention's sequences run along j, one per (map, i)
            xc = tr(a)                                               # rows (map, i, j)
            a = self._mha(layer.self_attn_col, xc, None, n * N, N, col_mask, p)   # rows (map, i, j)
            src, src_op = ops.layer_norm(ops.dropout(a, p), src, layer.norm1.weight, layer.norm1.bias, layer.norm1.eps, want_op=True)
            src, _ = ops.ffn_block(src, src_op, layer.linear1.weight, layer.linear1.bias, layer.linear2.weight,
                                   layer.linear2.bias, layer.norm2.weight, layer.norm2.bias, layer.norm2.eps, drop_p=p)
        return src.view(n, N, N, d)

    def _conv(self, maps):
        """maps [n, d, N, N] -> [n, N, N, d]: conv -> ReLU -> * 1/valid-count weight, per layer (:247-250)."""
        n, d, _, _ = maps.shape
        k = self.kernel
        x = maps
        for i, conv in enumerate(self.encoder.convs):
            pad = self.first_padding if i == 0 else 0
            Ho, Wo = x.shape[2] + 2 * pad - k + 1, x.shape[3] + 2 * pad - k + 1
            outs = []
            for m in x.split(1, 0):  # one map at a time: the im2col of a 128 x 128 map with k = 9 is ~1 GB in bf16
                cols = F.unfold(ops.to_operand(m), k, padding=pad)  # [1, d*k*k, Ho*Wo], channel order (c, kh, kw) = the weight's
                cols = cols[0].t().contiguous()  # [Ho*Wo, d*k*k]
                outs.append(ops.linear(cols, conv.weight.view(d, -1), conv.bias, relu=True))
            y = torch.stack(outs).view(n, Ho, Wo, d) * getattr(self, f"conv_weight_{i}")[None, :, :, None]
            x = y.permute(0, 3, 1, 2)
        return x.permute(0, 2, 3, 1)

    def forward(self, x):
        """x [layers, b, T, d] -> scores [layers, b, N, N]"""
        nl, b, t, d = x.shape
        maps = self.map_maker(x.reshape(-1, t, d))
        N = self.map_maker.map_size
        y = self._attn(maps) if self.temp_head == "attn" else self._conv(maps)  # [n, N, N, d]
        s = ops.linear(y.reshape(-1, d), self.predictor.weight.view(1, d), self.predictor.bias).view(nl, b, N, N)
        if self.training:
            return s
        return torch.sigmoid(s) * self.map_maker.mask2d
