"""B200 query decoder: drop-in for the reference ``build_decoder(cfg)``.

Interface mirrored (reference models/grounding_model/query_decoder.py):
  ``QueryDecoder(cfg).forward(memory_cache, vis_pos=None, text_cls=None)
      -> ([hs [nl,b,t,d], reference [nl,b,t,4]], (time_hs [nl,b,t,d], weights [nl,b,t,t]))`` (:83-147),
  ``self.decoder.bbox_embed`` assigned from outside (pipeline.py:50), and the state_dict keys of
  SURVEY.md 8b.

Layout: queries are kept batch-major ``[b*t, d]`` (row = video * t + frame slot) and the encoder
memory frame-major ``[n*M, d]`` (row = frame * M + token), so the time-aligned cross attention
(one query per frame against the M = HW+L tokens of that frame, query_decoder.py:350-429, 615-651) is
a batch of n single-query attentions over contiguous key blocks; the per-head concat
[content(32) ; positional(32)] of the box decoder is never materialised (two-part score in the
attention kernel).
"""
from __future__ import annotations

import math
from typing import Optional

import os

import torch
from torch import nn

from . import ops
from .encoder import batch_indices, fused_glue, _check_cfg
from .params import LinearP, MHAP, MLPP, NormP, OutProjOnly, SineTable, LearnedTable, xavier_reset

_const_cache = {}
_MULTI_STREAM = os.environ.get("STCAT_SINGLE_STREAM", "0") == "0"
_stream_cache = {}


def set_multi_stream(on: bool):
    """Run the decoder's three independent branches on side streams (default on)."""
    global _MULTI_STREAM
    _MULTI_STREAM = bool(on)


_MEMSIDE_SMS = int(os.environ.get("STCAT_MEMSIDE_SMS", "112"))
_FUSED_HEAD = os.environ.get("STCAT_FUSED_HEAD", "1") != "0"  # ops.box_head / mul_operand in the anchor-update chain (A/B switch)


def _side_streams(device):
    key = str(device)
    st = _stream_cache.get(key)
    if st is None:
        # (M) memory-side GEMMs: default priority; (A), (B) the latency-bound query chains: high priority, so that their small
        # kernels are scheduled ahead of pending CTAs of the big GEMMs / leaf-stream weight gradients (STCAT_CHAIN_PRIO=0: off)
        prio = -1 if os.environ.get("STCAT_CHAIN_PRIO", "1") != "0" else 0
        st = (torch.cuda.Stream(device), torch.cuda.Stream(device, priority=prio), torch.cuda.Stream(device, priority=prio))
        _stream_cache[key] = st
    return st


def _anchor_freq(device) -> torch.Tensor:
    key = ("anchor_freq", str(device))
    t = _const_cache.get(key)
    if t is None:
        k = torch.arange(128, dtype=torch.float32, device=device)
        t = 10000 ** (2 * torch.div(k, 2, rounding_mode="floor") / 128)
        _const_cache[key] = t
    return t


def anchor_sine_embed(anchor: torch.Tensor) -> torch.Tensor:
    """[..., 4] (cx, cy, w, h) -> [..., 512] ordered (y, x, w, h); 128 dims per coordinate, sin on even /
    cos on odd dims of 2*pi*c / 10000^(2*floor(k/2)/128)  (net_utils.py:29-56)."""
    p = (anchor * (2 * math.pi))[..., None] / _anchor_freq(anchor.device)  # [..., 4, 128]
    e = torch.stack((p[..., 0::2].sin(), p[..., 1::2].cos()), dim=-1).flatten(-2)  # [..., 4, 128]
    return torch.cat((e[..., 1, :], e[..., 0, :], e[..., 2, :], e[..., 3, :]), dim=-1)


def inverse_sigmoid(x: torch.Tensor, eps: float = 1e-3) -> torch.Tensor:
    """clamped logit (net_utils.py:59-63)."""
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


class _AnchorSineFn(torch.autograd.Function):
    """gen_sineembed_for_position (net_utils.py:29-56) as one kernel; also writes the bf16 GEMM-operand copy."""

    @staticmethod
    def forward(ctx, anchor):
        be = ops.get_backend()
        a = anchor.detach().float().contiguous().view(-1, 4)
        n = a.shape[0]
        out = torch.empty(n, 512, dtype=torch.float32, device=a.device)
        out_op = torch.empty(n, 512, dtype=torch.bfloat16, device=a.device) if ops.get_precision() == "bf16" else None
        be.anchor_sine_fwd(a, out, out_op)
        ctx.save_for_backward(a)
        ctx.shape = anchor.shape
        ctx.set_materialize_grads(False)
        lead = anchor.shape[:-1]
        if out_op is not None:
            out_op = out_op.view(*lead, 512)
            ctx.mark_non_differentiable(out_op)
        return out.view(*lead, 512), out_op

    @staticmethod
    def backward(ctx, g, _unused):
        if g is None:
            return None
        (a,) = ctx.saved_tensors
        da = torch.empty_like(a)
        ops.get_backend().anchor_sine_bwd(a, g.contiguous().view(-1, 512).float(), da)
        return da.view(ctx.shape)


class _BoxRefineFn(torch.autograd.Function):
    """sigmoid(delta + inverse_sigmoid(anchor)) (query_decoder.py:212-217; pipeline.py:88-95) as one kernel each way."""

    @staticmethod
    def forward(ctx, delta, anchor):
        be = ops.get_backend()
        d = delta.detach().float().contiguous()
        a = anchor.detach().float().expand_as(d).contiguous()
        out = torch.empty_like(d)
        be.box_refine_fwd(d, a, out)
        ctx.save_for_backward(out, a)
        return out

    @staticmethod
    def backward(ctx, g):
        out, a = ctx.saved_tensors
        g = g.contiguous().float()
        dd = torch.empty_like(out)
        da = torch.empty_like(out) if ctx.needs_input_grad[1] else None
        ops.get_backend().box_refine_bwd(out, a, g, dd, da)
        return dd, da


def anchor_sine_embed_op(anchor: torch.Tensor):
    """(sine embedding [..., 512], its GEMM-operand copy or None).  CUDA: one kernel; CPU tensors (the tests of the host-side
    composition) take the torch restatement below."""
    if anchor.is_cuda:
        return _AnchorSineFn.apply(anchor)
    return anchor_sine_embed(anchor), None


def box_refine(delta: torch.Tensor, anchor: torch.Tensor) -> torch.Tensor:
    if delta.is_cuda and delta.shape == anchor.shape:
        return _BoxRefineFn.apply(delta, anchor)
    return torch.sigmoid(delta + inverse_sigmoid(anchor))


def run_mlp(m: MLPP, x: torch.Tensor, training: bool = False, x_op=None) -> torch.Tensor:
    """Linear-ReLU stack (net_utils.py:20-26).  Hidden activations feed exactly one GEMM, so without dropout they are
    written in the GEMM-operand dtype by the producing epilogue (no cast kernels); ``x_op`` is the caller's operand
    copy of ``x``."""
    p = getattr(m, "dropout_p", None)
    if p is None:  # the reference's own MLP module (net_utils.py:7-19) assigned from outside
        dm = getattr(m, "dropout", 0)
        p = dm.p if isinstance(dm, nn.Dropout) else float(dm or 0)
    for i, layer in enumerate(m.layers):
        hidden = i < m.num_layers - 1
        x = ops.linear(x, layer.weight, layer.bias, relu=hidden, out_bf16=hidden and not (training and p),
                       x_op=x_op if i == 0 else None)
        if training and p:
            # the reference applies dropout after EVERY layer of temp_embed / action_embed, the output
            # layer included (net_utils.py:23-25, p = 0.3 hard-wired at pipeline.py:42-47).  These are
            # [nl*b*t, <=256] tensors; masks from the same counter-based stream as every other site (stcat_dropout).
            x = ops.dropout(x, p)
    return x


def _lin(p: LinearP, x, **kw):
    return ops.linear(x, p.weight, p.bias, **kw)


def _mha_proj(a: MHAP, which: int, x, **kw):
    """x W_part^T + b_part for part 0/1/2 = q/k/v of the packed nn.MultiheadAttention in-projection."""
    d = a.embed_dim
    return ops.linear(x, a.in_proj_weight, a.in_proj_bias, rows=(which * d, (which + 1) * d), **kw)


class _Ctx:
    """Per-forward shared tensors of the decoder (memory-side operands are built once and reused by all
    12 layers)."""

    def __init__(self, idx, mem, mem_pos, key_mask, n_mem_tokens, operands=None, stream=None):
        self.idx = idx
        self.b, self.t, self.n = idx["b"], idx["t"], idx["n"]
        self.M = n_mem_tokens
        self.key_mask = key_mask
        self.query_mask = idx["query_mask"]
        if operands is not None:  # written by ops.mem_operands straight from the encoder stream ``stream`` = (X, POS) [n, 1 + M, d]
            self.mem_op, self.pos_op, self.mempos_op = operands
            return
        self.mem_op = ops.to_operand(mem)  # [n*M, d]
        self.pos_op = ops.to_operand(mem_pos)
        self.mempos_op = ops.add(mem, mem_pos, as_operand=True)

    def frames(self, x):
        """padded [b*t, c] -> frames [n, c]"""
        return x if self.idx["identity"] else x.index_select(0, self.idx["dec_scatter"])

    def padded(self, x):
        """frames [n, c] -> zero-padded [b*t, c]"""
        if self.idx["identity"]:
            return x
        return torch.cat([x, x.new_zeros(1, x.shape[1])], 0).index_select(0, self.idx["dec_gather"])


class TransformerDecoderLayer(nn.Module):
    """Box-decoder layer parameters + forward (query_decoder.py:250-438).  ``from_scratch`` selects the cross attention:
    True (both shipped experiment files): the reference's own attention without in-projections over the per-head
    [content ; position] concat, run as a two-part score; False (the MDETR initialisation branch, :287-288): an
    ``nn.MultiheadAttention`` named ``cross_attn_image`` over q = content + sine + time, k = content + position."""

    def __init__(self, d: int, nhead: int, ffn: int, first: bool, from_scratch: bool = True):
        super().__init__()
        for nm in ("sa_qcontent_proj", "sa_qpos_proj", "sa_qtime_proj", "sa_kcontent_proj", "sa_kpos_proj",
                   "sa_ktime_proj", "sa_v_proj"):
            setattr(self, nm, LinearP(d, d))
        self.self_attn = MHAP(d, nhead)
        self.ca_qcontent_proj = LinearP(d, d)
        self.ca_qpos_proj = LinearP(d, d) if first else None  # layers >= 1: None (query_decoder.py:166-167)
        for nm in ("ca_kcontent_proj", "ca_kpos_proj", "ca_qtime_proj", "ca_v_proj", "ca_qpos_sine_proj"):
            setattr(self, nm, LinearP(d, d))
        self.from_scratch = bool(from_scratch)
        self.cross_attn = OutProjOnly(d) if from_scratch else None
        self.cross_attn_image = None if from_scratch else MHAP(d, nhead)
        self.linear1 = LinearP(d, ffn)
        self.linear2 = LinearP(ffn, d)
        self.norm1, self.norm3, self.norm4 = NormP(d), NormP(d), NormP(d)
        self.nhead = nhead
        self.d = d
        self.dropout_p = 0.0  # set by QueryDecoder (MODEL.STCAT.DROPOUT); active in train mode only

    def memory_side(self, c: _Ctx, is_first: bool):
        """Key / value projections of the encoder memory for this layer (query_decoder.py:355-366).  They do not
        depend on the queries, so the decoder computes them for all layers up front on a side stream."""
        if not self.from_scratch:
            # k = ca_kcontent(memory) + ca_kpos(pos), with ca_kpos(pos) a second time in the first layer (:365 and :384), then
            # the in-projections of cross_attn_image (torch functional.py:5866-5873) on k and on v = ca_v(memory)
            kpos = (c.pos_op, self.ca_kpos_proj.weight, self.ca_kpos_proj.bias)
            k = ops.linear_sum([(c.mem_op, self.ca_kcontent_proj.weight, self.ca_kcontent_proj.bias), kpos] +
                               ([kpos] if is_first else []), out_bf16=True)
            v = _lin(self.ca_v_proj, c.mem_op, out_bf16=True)
            return (_mha_proj(self.cross_attn_image, 1, k, out_bf16=True), None,
                    _mha_proj(self.cross_attn_image, 2, v, out_bf16=True))
        kp = _lin(self.ca_kpos_proj, c.pos_op, out_bf16=True)  # [n*M, d]
        vv = _lin(self.ca_v_proj, c.mem_op, out_bf16=True)
        if is_first:
            kc = ops.linear_sum([(c.mem_op, self.ca_kcontent_proj.weight, self.ca_kcontent_proj.bias),
                                 (c.pos_op, self.ca_kpos_proj.weight, self.ca_kpos_proj.bias)], out_bf16=True)
        else:
            kc = _lin(self.ca_kcontent_proj, c.mem_op, out_bf16=True)
        return kc, kp, vv

    def run(self, c: _Ctx, tgt, tgt_op, query_pos, query_time, time_op, query_sine, sine_op, is_first: bool, mem_kv, pos_op=None):
        """``*_op`` are the (non-differentiable) GEMM-operand copies of the tensors before them: every activation is
        cast once, by its producer where possible, not once per consuming Linear.  Intermediates that feed exactly
        one GEMM / attention (q, k, v, qc, qs) are written in the operand dtype by the producing epilogue."""
        d, H = self.d, self.nhead
        p = self.dropout_p if self.training else 0.0  # query_decoder.py:269,286,293-303: attention, dropout1/3/4, FFN inner
        pos_op = pos_op if pos_op is not None else ops.operand_copy(query_pos)
        # ---- temporal self attention over the t queries of each video (:329-345) ----
        L = lambda p: (p.weight, p.bias)
        q, k, v = ops.linear_group(
            [(tgt, tgt_op), (query_time, time_op), (query_pos, pos_op)],
            [{"terms": [(0, *L(self.sa_qcontent_proj)), (1, *L(self.sa_qtime_proj)), (2, *L(self.sa_qpos_proj))], "out_bf16": True},
             {"terms": [(0, *L(self.sa_kcontent_proj)), (1, *L(self.sa_ktime_proj)), (2, *L(self.sa_kpos_proj))], "out_bf16": True},
             {"terms": [(0, *L(self.sa_v_proj))], "out_bf16": True}])
        sa = self.self_attn
        Q, K, V = ops.linear_group(
            [(q, None), (k, None), (v, None)],
            [{"terms": [(i, sa.in_proj_weight, sa.in_proj_bias, (i * d, (i + 1) * d))], "out_bf16": True} for i in range(3)])
        o, _ = ops.attention(Q, K, V, c.b, H, c.t, c.t, float(d // H) ** -0.5, key_mask=c.query_mask, drop_p=p)
        tgt, tgt_op = ops.out_ln(o, tgt, sa.out_proj.weight, sa.out_proj.bias, self.norm1.weight, self.norm1.bias, self.norm1.eps, p)
        # ---- time-aligned cross attention: query of frame f sees only frame f's tokens (:350-429) ----
        kc, kp, vv = mem_kv() if callable(mem_kv) else mem_kv
        qc_terms = [(0, *L(self.ca_qcontent_proj))] + ([(1, *L(self.ca_qpos_proj))] if is_first else [])
        if self.from_scratch:
            qc, qs = ops.linear_group(
                [(tgt, tgt_op), (query_pos, pos_op), (query_sine, sine_op)],
                [{"terms": qc_terms, "out_bf16": True}, {"terms": [(2, *L(self.ca_qpos_sine_proj))], "out_bf16": True}])
            o, _ = ops.attention(c.frames(qc), kc, vv, c.n, H, 1, c.M, float(2 * d // H) ** -0.5, key_mask=c.key_mask,
                                 q2=c.frames(qs), k2=kp, drop_p=p)
            cross_out = self.cross_attn.out_proj
        else:
            # q = ca_qcontent(tgt) [+ ca_qpos(query_pos)] + ca_qpos_sine(sine) + ca_qtime(time) (:355-376: the per-head views
            # of :371-375 add element-wise), then nn.MultiheadAttention: in-projection, one query per frame, out-projection
            qa, qt = ops.linear_group(
                [(tgt, tgt_op), (query_pos, pos_op), (query_sine, sine_op), (query_time, time_op)],
                [{"terms": qc_terms + [(2, *L(self.ca_qpos_sine_proj))]}, {"terms": [(3, *L(self.ca_qtime_proj))]}])
            ca = self.cross_attn_image
            Q = _mha_proj(ca, 0, c.frames(ops.add(qa, qt)), out_bf16=True)
            o, _ = ops.attention(Q, kc, vv, c.n, H, 1, c.M, float(d // H) ** -0.5, key_mask=c.key_mask, drop_p=p)
            cross_out = ca.out_proj
        if c.idx["identity"]:  # frames == query slots: out-projection + dropout + residual + norm as one node
            tgt, tgt_op = ops.out_ln(o, tgt, cross_out.weight, cross_out.bias, self.norm3.weight, self.norm3.bias, self.norm3.eps, p)
        else:
            o = ops.dropout(_lin(cross_out, o), p)
            tgt, tgt_op = ops.layer_norm(c.padded(o), tgt, self.norm3.weight, self.norm3.bias, self.norm3.eps, want_op=True)
        # ---- FFN (:435-437) ----
        return ops.ffn_block(tgt, tgt_op, self.linear1.weight, self.linear1.bias, self.linear2.weight,
                             self.linear2.bias, self.norm4.weight, self.norm4.bias, self.norm4.eps, drop_p=p)


_GROUP_MEMSIDE = os.environ.get("STCAT_GROUP_MEMSIDE", "1") != "0"


def memory_side_pair(c: _Ctx, box: "TransformerDecoderLayer", tim: "TimeDecoderLayer", is_first: bool):
    """The memory-side key / value projections of one box-decoder layer and one time-decoder layer (query_decoder.py:355-366,
    633-639) as ONE grouped launch: five [n M, 256] x [256, 256] GEMMs of one tile wave each become 5 x 106 tiles of one
    persistent launch (tile pipelining across jobs); backward: one grouped data-gradient launch (three terms accumulate in
    TMEM for the memory operand), one grouped weight-gradient launch, one grouped column-sum launch instead of 14 launches.
    Returns ((kc, kp, vv), (K, V))."""
    if not (_GROUP_MEMSIDE and box.from_scratch):
        return box.memory_side(c, is_first), tim.memory_side(c)
    d = box.d
    L = lambda p: (p.weight, p.bias)
    ca = tim.cross_attn_image
    kc_terms = [(0, *L(box.ca_kcontent_proj))] + ([(1, *L(box.ca_kpos_proj))] if is_first else [])
    kc, kp, vv, K, V = ops.linear_group(
        [(c.mem_op, None), (c.pos_op, None), (c.mempos_op, None)],
        [{"terms": kc_terms, "out_bf16": True},
         {"terms": [(1, *L(box.ca_kpos_proj))], "out_bf16": True},
         {"terms": [(0, *L(box.ca_v_proj))], "out_bf16": True},
         {"terms": [(2, ca.in_proj_weight, ca.in_proj_bias, (d, 2 * d))], "out_bf16": True},
         {"terms": [(0, ca.in_proj_weight, ca.in_proj_bias, (2 * d, 3 * d))], "out_bf16": True}])
    return (kc, kp, vv), (K, V)


class TransformerDecoder(nn.Module):
    """Anchor-refining box decoder (query_decoder.py:150-247)."""

    def __init__(self, d: int, nhead: int, ffn: int, num_layers: int, query_dim: int, from_scratch: bool = True):
        super().__init__()
        self.layers = nn.ModuleList(TransformerDecoderLayer(d, nhead, ffn, first=i == 0, from_scratch=from_scratch)
                                    for i in range(num_layers))
        self.num_layers = num_layers
        self.norm = NormP(d)
        self.query_scale = MLPP(d, d, d, 2)
        self.ref_point_head = MLPP(query_dim // 2 * d, d, d, 2)
        self.bbox_embed = None  # assigned by the pipeline (pipeline.py:50)
        self.query_dim = query_dim
        self.d_model = d

    def run(self, c: _Ctx, tgt, anchor, query_time, mem_kv):
        d = self.d_model
        out, out_op = tgt, None
        inter, refs = [], [anchor]
        time_op = ops.operand_copy(query_time)
        # bf16 mode on the device: the last Linear of bbox_embed + anchor refinement + the next layer's sine embedding are one
        # launch (ops.box_head) and query_sine is written directly as a GEMM operand (ops.mul_operand)
        be_ = self.bbox_embed
        fuse_head = (_FUSED_HEAD and fused_glue() and be_ is not None and self.query_dim == 4 and anchor.shape[-1] == 4 and len(be_.layers) >= 2
                     and be_.layers[-1].weight.shape[0] == 4 and all(l_.bias is not None for l_ in be_.layers))
        nxt = None  # (sine, sine_op) of the refined anchor, from the previous layer's box head
        for li, layer in enumerate(self.layers):
            qpos_op = None
            if nxt is not None:
                sine, sine_op = nxt
            else:
                sine, sine_op = anchor_sine_embed_op(anchor[..., : self.query_dim])  # [b*t, 512] (+ operand copy)
            if sine_op is None:
                sine_op = ops.operand_copy(sine)
            rp, qsc = self.ref_point_head.layers, self.query_scale.layers
            if li == 0:
                query_pos = run_mlp(self.ref_point_head, sine, x_op=sine_op)
                qsine, qsine_op = sine[..., :d], (None if sine_op is None else sine_op[..., :d])
            else:
                # ref_point_head(sine) and query_scale(out) are independent two-layer MLPs: layer by layer in one launch
                h1, h2 = ops.linear_group([(sine, sine_op), (out, out_op)],
                                          [{"terms": [(0, rp[0].weight, rp[0].bias)], "relu": True, "out_bf16": True},
                                           {"terms": [(1, qsc[0].weight, qsc[0].bias)], "relu": True, "out_bf16": True}])
                query_pos, scale = ops.linear_group([(h1, None), (h2, None)],
                                                    [{"terms": [(0, rp[1].weight, rp[1].bias)]},
                                                     {"terms": [(1, qsc[1].weight, qsc[1].bias)]}])
                if fuse_head and not sine.requires_grad and sine.dim() == 2:
                    qsine, qsine_op, qpos_op = ops.mul_operand(sine, scale, query_pos)  # + the operand copy of query_pos
                else:
                    qsine, qsine_op = sine[..., :d] * scale, None
            out, out_op = layer.run(c, out, out_op, query_pos, query_time, time_op, qsine, qsine_op, li == 0, mem_kv[li],
                                    pos_op=qpos_op)
            nxt = None
            if self.bbox_embed is not None:
                if fuse_head and out.dim() == 2:
                    last = li == self.num_layers - 1
                    new_anchor, s_, so_ = ops.box_mlp_head(be_.layers, out, out_op, anchor, want_sine=not last)
                    nxt = None if last else (s_, so_)
                else:
                    new_anchor = box_refine(run_mlp(self.bbox_embed, out, x_op=out_op), anchor)
                if li != self.num_layers - 1:
                    refs.append(new_anchor)
                anchor = new_anchor.detach()
            inter.append(out)
        # the shared output norm of every layer's queries (query_decoder.py:222) as ONE launch over the stacked outputs, after
        # the loop: it is not an input of the next layer, so it does not belong on the layers' dependent chain
        hs = ops.layer_norm(torch.stack(inter).view(self.num_layers * c.b * c.t, d), None, self.norm.weight, self.norm.bias,
                            self.norm.eps).view(self.num_layers, c.b, c.t, d)
        if self.bbox_embed is not None:
            ref = torch.stack(refs).view(len(refs), c.b, c.t, -1)
        else:
            ref = anchor.view(1, c.b, c.t, -1)
        return [hs, ref]


class TimeDecoderLayer(nn.Module):
    """Temporal decoder layer (query_decoder.py:553-660)."""

    def __init__(self, d: int, nhead: int, ffn: int):
        super().__init__()
        self.self_attn = MHAP(d, nhead)
        self.cross_attn_image = MHAP(d, nhead)
        self.linear1 = LinearP(d, ffn)
        self.linear2 = LinearP(ffn, d)
        self.norm1, self.norm3, self.norm4 = NormP(d), NormP(d), NormP(d)
        self.nhead = nhead
        self.d = d
        self.dropout_p = 0.0  # set by QueryDecoder

    def memory_side(self, c: _Ctx):
        """nn.MultiheadAttention in-projection of key = memory + pos and value = memory (:633-639)."""
        K = _mha_proj(self.cross_attn_image, 1, c.mempos_op, out_bf16=True)
        V = _mha_proj(self.cross_attn_image, 2, c.mem_op, out_bf16=True)
        return K, V

    def run(self, c: _Ctx, tgt, tgt_op, query_pos, query_pos_frames, qpos_plus_time, mem_kv):
        d, H = self.d, self.nhead
        scale = float(d // H) ** -0.5
        p = self.dropout_p if self.training else 0.0  # query_decoder.py:565-580
        qk = tgt + qpos_plus_time
        qk_op = ops.operand_copy(qk)  # one cast for the q and k projections
        sa = self.self_attn
        Q, K, V = ops.linear_group(
            [(qk, qk_op), (tgt, tgt_op)],
            [{"terms": [(0 if i < 2 else 1, sa.in_proj_weight, sa.in_proj_bias, (i * d, (i + 1) * d))], "out_bf16": True}
             for i in range(3)])
        o, weights = ops.attention(Q, K, V, c.b, H, c.t, c.t, scale, key_mask=c.query_mask, need_pavg=True, drop_p=p)
        tgt, _ = ops.out_ln(o, tgt, sa.out_proj.weight, sa.out_proj.bias, self.norm1.weight, self.norm1.bias, self.norm1.eps, p)
        # cross attention, one query per frame (:615-651)
        Q = _mha_proj(self.cross_attn_image, 0, c.frames(tgt) + query_pos_frames, out_bf16=True)
        K, V = mem_kv() if callable(mem_kv) else mem_kv
        o, _ = ops.attention(Q, K, V, c.n, H, 1, c.M, scale, key_mask=c.key_mask, drop_p=p)
        ca_out = self.cross_attn_image.out_proj
        if c.idx["identity"]:
            tgt, tgt_op = ops.out_ln(o, tgt, ca_out.weight, ca_out.bias, self.norm3.weight, self.norm3.bias, self.norm3.eps, p)
        else:
            o = ops.dropout(_lin(ca_out, o), p)
            tgt, tgt_op = ops.layer_norm(c.padded(o), tgt, self.norm3.weight, self.norm3.bias, self.norm3.eps, want_op=True)
        tgt, tgt_op = ops.ffn_block(tgt, tgt_op, self.linear1.weight, self.linear1.bias, self.linear2.weight,
                                    self.linear2.bias, self.norm4.weight, self.norm4.bias, self.norm4.eps, drop_p=p)
        return tgt, tgt_op, weights


class TimeDecoder(nn.Module):
    def __init__(self, d: int, nhead: int, ffn: int, num_layers: int):
        super().__init__()
        self.layers = nn.ModuleList(TimeDecoderLayer(d, nhead, ffn) for _ in range(num_layers))
        self.num_layers = num_layers
        self.norm = NormP(d)
        self.d_model = d

    def run(self, c: _Ctx, tgt, query_pos, query_time, mem_kv):
        out, out_op = tgt, None
        inter, ws = [], []
        qpt = query_pos + query_time
        qpf = c.frames(query_pos)
        for li, layer in enumerate(self.layers):
            out, out_op, w = layer.run(c, out, out_op, query_pos, qpf, qpt, mem_kv[li])
            inter.append(out)
            ws.append(w)
        # one launch for the shared output norm of all layers (query_decoder.py:527), off the layers' dependent chain
        hs = ops.layer_norm(torch.stack(inter).view(self.num_layers * c.b * c.t, self.d_model), None, self.norm.weight,
                            self.norm.bias, self.norm.eps)
        return hs.view(self.num_layers, c.b, c.t, self.d_model), torch.stack(ws)


class TemplateGenerator(nn.Module):
    """FiLM-style anchors and temporal content query (query_decoder.py:441-475)."""

    def __init__(self, d: int, query_dim: int):
        super().__init__()
        self.content_proj = LinearP(d, d)
        self.gamma_proj = LinearP(d, d)
        self.beta_proj = LinearP(d, d)
        self.anchor_proj = LinearP(d, query_dim)

    def run(self, idx, frames_cls, videos_cls):
        content = _lin(self.content_proj, videos_cls)
        gamma = torch.tanh(_lin(self.gamma_proj, videos_cls))
        beta = torch.tanh(_lin(self.beta_proj, videos_cls))
        if idx["identity"]:
            mod = gamma * frames_cls + beta
            temp_query = content.expand(idx["n"], -1)
        else:
            f2v = idx["f2v"]
            mod = gamma.index_select(0, f2v) * frames_cls + beta.index_select(0, f2v)
            temp_query = content.index_select(0, f2v)
        return _lin(self.anchor_proj, mod), temp_query

    def run_fused(self, idx, frames_cls, videos_cls):
        """(sigmoid(anchor_proj(mod)) [n, query_dim], temp_query [n, d]) in two launches (ops.template); bf16 mode."""
        one = idx["identity"]
        L = lambda p: (p.weight, p.bias)
        return ops.template(videos_cls, frames_cls, *L(self.content_proj), *L(self.gamma_proj), *L(self.beta_proj),
                            *L(self.anchor_proj), None if one else idx["f2v"], None if one else idx["vid_start"])


class QueryDecoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        _check_cfg(cfg)
        S = cfg.MODEL.STCAT
        d = S.HIDDEN
        self.d_model = d
        self.query_pos_dim = S.QUERY_DIM
        self.nhead = S.HEADS
        self.video_max_len = cfg.INPUT.MAX_VIDEO_LEN
        self.return_weights = cfg.SOLVER.USE_ATTN
        self.dropout_p = float(S.DROPOUT)
        self.template_generator = TemplateGenerator(d, S.QUERY_DIM)
        self.decoder = TransformerDecoder(d, S.HEADS, S.FFN_DIM, S.DEC_LAYERS, S.QUERY_DIM, bool(S.FROM_SCRATCH))
        self.temp_decoder = TimeDecoder(d, S.HEADS, S.FFN_DIM, S.DEC_LAYERS)
        max_len = self.video_max_len + 1
        self.time_embed = LearnedTable(max_len, d) if S.USE_LEARN_TIME_EMBED else SineTable(max_len, d)
        for layer in (*self.decoder.layers, *self.temp_decoder.layers):
            layer.dropout_p = self.dropout_p
        xavier_reset(self)

    def forward(self, memory_cache: dict, vis_pos: Optional[torch.Tensor] = None, text_cls=None):
        d = self.d_model
        mem_sf = memory_cache["encoded_memory"]  # [M, n, d]
        memory_mask = memory_cache["mask"]  # [n, M] bool
        durations = list(memory_cache["durations"])
        H, W = memory_cache["fea_map_size"]
        n_vis = H * W
        M, n, _ = mem_sf.shape
        idx = batch_indices(durations, mem_sf.device)
        b, t = idx["b"], idx["t"]
        if t > self.video_max_len + 1:
            raise ValueError(f"clip of {t} frames exceeds INPUT.MAX_VIDEO_LEN={self.video_max_len}")
        key_mask = memory_mask.to(torch.uint8).contiguous()
        enc_stream = memory_cache.get("_stream") if fused_glue() else None
        if enc_stream is not None and (vis_pos is None or enc_stream[2] != vis_pos.data_ptr()
                                       or tuple(enc_stream[0].shape) != (n, M + 1, d)
                                       or self.template_generator.anchor_proj.weight.shape[0] > 8):
            enc_stream = None
        if enc_stream is not None:
            # this package's encoder handed over its frame-major stream: the three memory-side GEMM operands and the frame-CLS
            # rows in one launch, the template generator in two (csrc/assembly.cu); one / three launches in the backward pass
            operands = ops.mem_operands(enc_stream[0], enc_stream[1])
            memory_cache["_stream_used"] = True  # dp.GradSync.attach: the gradient boundary is the stream, not its views
            frames_cls = operands[3]
            operands = operands[:3]
            mem = mem_pos = None
            anchor_frames, temp_query = self.template_generator.run_fused(idx, frames_cls, memory_cache["videos_cls"])
            like = frames_cls
        else:
            operands = None
            mem = mem_sf.transpose(0, 1).reshape(n * M, d).float()  # frame-major copy (one pass over 14 MB)
            p_v = vis_pos.flatten(2).transpose(1, 2)  # [n, HW, d]
            mem_pos = torch.cat([p_v, p_v.new_zeros(n, M - n_vis, d)], 1).reshape(n * M, d).float()
            # templates (:97-120)
            pos_query, temp_query = self.template_generator.run(idx, memory_cache["frames_cls"], memory_cache["videos_cls"])
            anchor_frames = torch.sigmoid(pos_query)
            like = mem
        qt = self.time_embed.rows(t)
        query_time = (qt if b == 1 else qt.repeat(b, 1)).contiguous()
        tgt = like.new_zeros(b * t, d)
        nl = self.decoder.num_layers
        use_streams = like.is_cuda and _MULTI_STREAM
        if not use_streams:
            c = _Ctx(idx, mem, mem_pos, key_mask, M, operands, enc_stream)
            anchors = c.padded(anchor_frames)  # [b*t, 4]
            query_temporal = c.padded(temp_query)  # [b*t, d]
            pairs = [memory_side_pair(c, self.decoder.layers[i], self.temp_decoder.layers[i], i == 0) for i in range(nl)]
            box_kv = [p_[0] for p_ in pairs]
            time_kv = [p_[1] for p_ in pairs]
            outputs = self.decoder.run(c, tgt, anchors, query_time, box_kv)
            outputs_temp = self.temp_decoder.run(c, tgt.clone(), query_temporal, query_time, time_kv)
            return outputs, outputs_temp
        # Three concurrent branches (they only share read-only inputs): (M) the memory-side key/value projections of
        # all 12 layers -- big GEMMs that do not depend on the queries; (A) the box decoder's query chain; (B) the time
        # decoder's query chain.  A and B are chains of tiny [t, 256] ops (launch-latency bound), so running them side
        # by side with M fills the machine.  autograd replays each node's backward on its forward stream, so the
        # backward pass has the same concurrency.
        cur = torch.cuda.current_stream()
        sM, sA, sB = _side_streams(like.device)
        fork = cur.record_event()
        for st_ in (sM, sA, sB):
            st_.wait_event(fork)
        with torch.cuda.stream(sM):
            c = _Ctx(idx, mem, mem_pos, key_mask, M, operands, enc_stream)
            ctx_ready = sM.record_event()
            box_kv, time_kv, ev_box, ev_time = [], [], [], []
            # these persistent GEMMs (and their backward) share the machine with the two query chains: keep some SMs free for
            # the chains' small kernels (ops.sm_limit; STCAT_MEMSIDE_SMS overrides, 0 = no cap)
            with ops.sm_limit(_MEMSIDE_SMS):
                for i in range(nl):
                    bkv, tkv = memory_side_pair(c, self.decoder.layers[i], self.temp_decoder.layers[i], i == 0)
                    box_kv.append(bkv)
                    time_kv.append(tkv)
                    ev = sM.record_event()
                    ev_box.append(ev)
                    ev_time.append(ev)

        def waiter(stream, evs, vals):
            def mk(i):
                def get():
                    stream.wait_event(evs[i])
                    for tns in vals[i]:  # allocated on sM, consumed here (and in backward) on another stream
                        if tns is not None:  # FROM_SCRATCH False has no separate positional key part
                            tns.record_stream(stream)
                    return vals[i]
                return get
            return [mk(i) for i in range(len(vals))]

        for tns in (anchor_frames, temp_query, query_time, tgt):  # allocated on `cur`, consumed on the side streams
            tns.record_stream(sA)
            tns.record_stream(sB)
        for tns in ((mem, mem_pos) if operands is None else operands):
            tns.record_stream(sM)

        with torch.cuda.stream(sA):
            sA.wait_event(ctx_ready)
            anchors = c.padded(anchor_frames)
            outputs = self.decoder.run(c, tgt, anchors, query_time, waiter(sA, ev_box, box_kv))
        with torch.cuda.stream(sB):
            sB.wait_event(ctx_ready)
            query_temporal = c.padded(temp_query)
            outputs_temp = self.temp_decoder.run(c, tgt.clone(), query_temporal, query_time, waiter(sB, ev_time, time_kv))
        for st_ in (sM, sA, sB):
            cur.wait_stream(st_)
        for tns in (*outputs, *outputs_temp):
            tns.record_stream(cur)
        return outputs, outputs_temp


def build_decoder(cfg) -> QueryDecoder:
    """Mirror of models/grounding_model/__init__.py:8-9."""
    return QueryDecoder(cfg)
