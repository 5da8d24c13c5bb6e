"""TEST INFRASTRUCTURE ONLY -- CPU restatement (the "oracle") of the STCAT hot path.

This file is the checker the CUDA path is compared with.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import
it; nothing under ``stcat_b200/`` does (the product path has no CPU fallback).

Parity status: the reference ships no tests or golden vectors for this path (SURVEY.md 4, 8c:
"parity unpinned by the reference").  This restatement is therefore pinned against **outputs of the
reference itself run in the build container**: ``oracle/make_golden.py`` imports the unmodified
reference from /root/reference, runs it on seeded inputs/weights and commits inputs-by-recipe and
outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file against those
fixtures (and against the live reference when /root/reference exists).

It is a *restatement*, not a copy: a purely functional implementation over a flat ``{name: tensor}``
parameter dict that uses the reference's ``state_dict`` key names.  Every function cites the
reference lines whose arithmetic it follows.  All arithmetic is floating point; ``Prec`` selects the
dtype (fp32 like the reference, or fp64) and optionally emulates bf16 rounding of matmul operands
(used to gate the bf16 tensor-core kernels).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]


class _RoundBF16(torch.autograd.Function):
    """round-to-nearest-even to bf16, kept in the working dtype; the gradient passes through unrounded (the oracle's
    gradients are those of the working dtype -- fp64 in the gradient gates -- on bf16-rounded operands)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


@dataclass
class Prec:
    """dtype: working dtype.  round_operands="bf16": every matmul operand is rounded to bf16 first.
    kernel_stores (with round_operands="bf16"): additionally round, to bf16, every intermediate that the B200 path keeps in
    bf16 between kernels (projection outputs that feed only a GEMM / attention, the FFN hidden activation, the attention
    output) and the probabilities where the attention kernels round them -- the rounding points are listed function by
    function below and cite the kernel / host line that makes them ("kernel-matched oracle")."""
    dtype: torch.dtype = torch.float32
    round_operands: Optional[str] = None
    kernel_stores: bool = False

    def r(self, x: Tensor) -> Tensor:
        if self.round_operands == "bf16":
            return _RoundBF16.apply(x)
        return x

    def s(self, x: Tensor) -> Tensor:
        """a value the CUDA path stores as bf16 (identity unless kernel_stores)"""
        if self.kernel_stores and self.round_operands == "bf16":
            return _RoundBF16.apply(x)
        return x

    @property
    def km(self) -> bool:
        return bool(self.kernel_stores and self.round_operands == "bf16")


FP32 = Prec()
BF16_KERNEL = Prec(round_operands="bf16", kernel_stores=True)


# ----------------------------------------------------------------------------------------------
# small building blocks
# ----------------------------------------------------------------------------------------------
def linear(x: Tensor, w: Tensor, b: Optional[Tensor], prec: Prec = FP32, store: bool = False) -> Tensor:
    """y = x W^T + b  (torch.nn.Linear; used everywhere on the path).  ``store``: the CUDA path writes this output in bf16
    from the GEMM epilogue (``out_bf16=True`` at the call site in stcat_b200/{ops,decoder}.py)."""
    y = prec.r(x) @ prec.r(w).t()
    y = y if b is None else y + b
    return prec.s(y) if store else y


def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """nn.LayerNorm(256), eps 1e-5 (modal_encoder.py:218-219, query_decoder.py:296-299,573-576)."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def mlp(P: Params, prefix: str, x: Tensor, num_layers: int, prec: Prec = FP32) -> Tensor:
    """Linear-ReLU stack, no activation after the last layer (net_utils.py:7-26; dropout omitted: eval)."""
    for i in range(num_layers):
        x = linear(x, P[f"{prefix}.layers.{i}.weight"], P[f"{prefix}.layers.{i}.bias"], prec)
        if i < num_layers - 1:
            x = prec.s(torch.relu(x))  # hidden activations feed one GEMM: bf16 from the epilogue (decoder.py run_mlp)
    return x


def inverse_sigmoid(x: Tensor, eps: float = 1e-3) -> Tensor:
    """logit with clamps (net_utils.py:59-63)."""
    x = x.clamp(0, 1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def seq_sine_table(max_len: int, d_model: int, dtype=torch.float32) -> Tensor:
    """SeqEmbeddingSine buffer ``te`` [max_len,1,d] (position_encoding.py:21-33).

    The reference builds the table in fp32 with torch.exp / sin / cos; we do the same in fp32 and
    cast, so the table is bit-identical to the reference's buffer.
    """
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    te = torch.zeros(max_len, 1, d_model)
    te[:, 0, 0::2] = torch.sin(position * div_term)
    te[:, 0, 1::2] = torch.cos(position * div_term)
    return te.to(dtype)


def anchor_sine_embed(anchor: Tensor) -> Tensor:
    """gen_sineembed_for_position for 4-d anchors (net_utils.py:29-56).

    anchor [..., 4] = (cx, cy, w, h) in (0,1).  Output [..., 512] ordered (y, x, w, h), 128 dims each,
    dims interleaved sin (even k) / cos (odd k) with frequency 10000^(2*floor(k/2)/128), scale 2*pi.
    """
    scale = 2 * math.pi
    k = torch.arange(128, dtype=torch.float32)
    dim_t = (10000 ** (2 * torch.div(k, 2, rounding_mode="floor") / 128)).to(anchor.dtype)

    def emb(c):
        p = (c * scale)[..., None] / dim_t
        return torch.stack((p[..., 0::2].sin(), p[..., 1::2].cos()), dim=-1).flatten(-2)

    return torch.cat((emb(anchor[..., 1]), emb(anchor[..., 0]), emb(anchor[..., 2]), emb(anchor[..., 3])), dim=-1)


def image_sine_pos(mask: Tensor, num_pos_feats: int = 128, temperature: float = 10000.0,
                   dtype=torch.float32) -> Tensor:
    """PositionEmbeddingSine(128, normalize=True) (vision_model/position_encoding.py:70-94).

    Upstream of the hot path (it arrives as ``vis_pos``); restated here only so that synthetic inputs
    have the real positional structure.  mask [n,H,W] bool (True = padded) -> [n, 256, H, W].
    """
    not_mask = ~mask
    y_embed = not_mask.cumsum(1, dtype=torch.float32)
    x_embed = not_mask.cumsum(2, dtype=torch.float32)
    eps, scale = 1e-6, 2 * math.pi
    y_embed = y_embed / (y_embed[:, -1:, :] + eps) * scale
    x_embed = x_embed / (x_embed[:, :, -1:] + eps) * scale
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / num_pos_feats)
    pos_x = x_embed[:, :, :, None] / dim_t
    pos_y = y_embed[:, :, :, None] / dim_t
    pos_x = torch.stack((pos_x[..., 0::2].sin(), pos_x[..., 1::2].cos()), dim=4).flatten(3)
    pos_y = torch.stack((pos_y[..., 0::2].sin(), pos_y[..., 1::2].cos()), dim=4).flatten(3)
    return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2).to(dtype)


# ----------------------------------------------------------------------------------------------
# attention
# ----------------------------------------------------------------------------------------------
def kernel_attention_variant(N: int, nhead: int, Lq: int, Lk: int, two_part: bool, need_weights: bool) -> str:
    """Which bf16 attention kernel the C ABI runs for a shape (the dispatch order of csrc/attention_simt.cu:
    attention_fwd_impl), reduced to what matters for rounding:
      "tc"   tcgen05 kernel (csrc/attention_tc.cu: Lq == Lk in [64, 512], one score part, no weights output; sequences of
             <= 128 tokens only when N*nhead >= 16): P = bf16(exp(s - rowmax)) un-normalised, O = (P V) / rowsum(fp32 exp)
      "mma"  mma.sync kernel (csrc/attention_small_mma.cu: Lq, Lk <= 128, one part): P = bf16(softmax) then P V
      "f32p" single-query / shared-memory / generic kernels: probabilities stay fp32
    In every variant q is NOT pre-scaled (the kernels scale the fp32 scores) and the output is stored as bf16."""
    if Lq == 1 and not need_weights and Lk <= 4096:
        return "f32p"
    if not two_part and not need_weights and Lq == Lk and 64 <= Lq <= 512 and (Lq > 128 or N * nhead >= 16):
        return "tc"
    if not two_part and 2 <= Lq <= 128 and 1 <= Lk <= 128:
        return "mma"
    return "f32p"


TC_KEY_TILE = 512  # keys per softmax tile of the tcgen05 kernel: the whole row (csrc/attention_tc.cu AT_KMAX), one row max


def attention_core(q: Tensor, k: Tensor, v: Tensor, nhead: int, key_padding_mask: Optional[Tensor],
                   scaling: float, prec: Prec = FP32, q2: Optional[Tensor] = None, k2: Optional[Tensor] = None,
                   need_weights: bool = True) -> Tuple[Tensor, Tensor]:
    """softmax((q*scaling) k^T + mask) v per head.

    q [Lq,N,Eq], k [Lk,N,Eq], v [Lk,N,Ev] (seq-first); key_padding_mask [N,Lk] bool, True -> -inf.
    Returns (o [Lq,N,Ev], P [N,h,Lq,Lk]).  Follows torch F.multi_head_attention_forward
    (nn/functional.py:6630-6665 in torch 2.11: q scaled first, bmm, masked -inf, softmax, bmm) and the
    reference's custom variant attention.py:283-383 (explicit max-subtraction :379-380, which does
    not change the value of the softmax).

    q2 / k2 (kernel-matched mode only): a second score part, s = q.k + q2.k2 per head -- the reference's per-head
    concat [content ; position] (query_decoder.py:371-384) as the CUDA path evaluates it (csrc/attention_sq.cu).
    """
    Lq, N, Eq = q.shape
    Lk = k.shape[0]
    Ev = v.shape[2]
    dq, dv = Eq // nhead, Ev // nhead
    heads = lambda t, L, dd: t.reshape(L, N, nhead, dd).permute(1, 2, 0, 3)
    if not prec.km:
        assert q2 is None
        qh = heads(prec.r(q * scaling), Lq, dq)  # [N,h,Lq,dq]
        kh = heads(prec.r(k), Lk, dq)
        vh = heads(prec.r(v), Lk, dv)
        s = qh @ kh.transpose(-1, -2)  # [N,h,Lq,Lk]
        if key_padding_mask is not None:
            s = s.masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
        p = torch.softmax(s - s.max(dim=-1, keepdim=True)[0], dim=-1)
        o = prec.r(p) @ vh  # [N,h,Lq,dv]
        o = o.permute(2, 0, 1, 3).reshape(Lq, N, Ev)
        return o, p
    # ---- kernel-matched: bf16 q, k, v as stored by the projection epilogues; fp32 scores scaled after the product ----
    variant = kernel_attention_variant(N, nhead, Lq, Lk, q2 is not None, need_weights)
    s = heads(prec.r(q), Lq, dq) @ heads(prec.r(k), Lk, dq).transpose(-1, -2)
    if q2 is not None:
        s = s + heads(prec.r(q2), Lq, dq) @ heads(prec.r(k2), Lk, dq).transpose(-1, -2)
    s = s * scaling
    if key_padding_mask is not None:
        s = s.masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
    vh = heads(prec.r(v), Lk, dv)
    p = torch.softmax(s - s.max(dim=-1, keepdim=True)[0], dim=-1)
    if variant == "tc":
        # attention_tc.cu: per 256-key tile, e = exp(s - m_run) with m_run the running row max up to and including the
        # tile, rounded to bf16 for the P V product; earlier partial outputs are rescaled by exp(m_old - m_new) in fp32;
        # the normaliser is the fp32 sum of the unrounded e
        o = None
        m_run = None
        den = None
        for k0 in range(0, Lk, TC_KEY_TILE):
            st = s[..., k0:k0 + TC_KEY_TILE]
            m_t = st.max(dim=-1, keepdim=True)[0]
            m_new = m_t if m_run is None else torch.maximum(m_run, m_t)
            m_safe = torch.where(torch.isinf(m_new), torch.zeros_like(m_new), m_new)
            e = torch.exp(st - m_safe)
            part = prec.r(e) @ vh[:, :, k0:k0 + TC_KEY_TILE]
            if o is None:
                o, den = part, e.sum(-1, keepdim=True)
            else:
                alpha = torch.exp(torch.where(torch.isinf(m_run), torch.zeros_like(m_run), m_run) - m_safe)
                alpha = torch.where(torch.isinf(m_run), torch.zeros_like(alpha), alpha)
                o, den = o * alpha + part, den * alpha + e.sum(-1, keepdim=True)
            m_run = m_new
        o = o / den.clamp_min(1e-300)
    elif variant == "mma":
        o = prec.r(p) @ vh
    else:
        o = p @ vh
    o = prec.s(o.permute(2, 0, 1, 3).reshape(Lq, N, Ev))
    return o, p


def torch_mha(P: Params, prefix: str, query: Tensor, key: Tensor, value: Tensor, nhead: int,
              key_padding_mask: Optional[Tensor], prec: Prec = FP32, need_weights: bool = True) -> Tuple[Tensor, Tensor]:
    """torch.nn.MultiheadAttention forward with packed in_proj (used at modal_encoder.py:212,236;
    query_decoder.py:269,341,565-566,604,633).  Returns (out [Lq,N,E], head-averaged P [N,Lq,Lk]).
    ``need_weights``: whether the CUDA path asks the kernel for the head-averaged probabilities (it does only where the
    reference uses them, query_decoder.py:604); selects the kernel in kernel-matched mode, the value is returned either way."""
    E = query.shape[-1]
    W, b = P[f"{prefix}.in_proj_weight"], P[f"{prefix}.in_proj_bias"]
    q = linear(query, W[:E], b[:E], prec, store=True)  # q, k, v feed only the attention kernel: bf16 from the epilogue
    k = linear(key, W[E:2 * E], b[E:2 * E], prec, store=True)
    v = linear(value, W[2 * E:], b[2 * E:], prec, store=True)
    o, p = attention_core(q, k, v, nhead, key_padding_mask, float(E // nhead) ** -0.5, prec, need_weights=need_weights)
    out = linear(o, P[f"{prefix}.out_proj.weight"], P[f"{prefix}.out_proj.bias"], prec)
    return out, p.mean(dim=1)


def custom_mha(P: Params, prefix: str, query: Tensor, key: Tensor, value: Tensor, nhead: int,
               key_padding_mask: Optional[Tensor], prec: Prec = FP32, q2: Optional[Tensor] = None,
               k2: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """Reference attention.MultiheadAttention: no in-projection, embed 2d (dk=64/head), vdim d (dv=32),
    scaling = dk^-0.5, out_proj Linear(d,d) (attention.py:86-113,275-393).  With q2 / k2 (kernel-matched mode) the
    per-head concat is given as its two 32-wide parts and ``query`` / ``key`` are the content parts."""
    Eq = query.shape[-1] * (2 if q2 is not None else 1)
    o, p = attention_core(query, key, value, nhead, key_padding_mask, float(Eq // nhead) ** -0.5, prec, q2=q2, k2=k2,
                          need_weights=False)
    out = linear(o, P[f"{prefix}.out_proj.weight"], P[f"{prefix}.out_proj.bias"], prec)
    return out, p.mean(dim=1)


# ----------------------------------------------------------------------------------------------
# encoder  (modal_encoder.py)
# ----------------------------------------------------------------------------------------------
def encoder_layer(P: Params, prefix: str, src: Tensor, mask: Optional[Tensor], pos: Tensor, nhead: int,
                  prec: Prec = FP32) -> Tensor:
    """Post-norm TransformerEncoderLayer.forward (modal_encoder.py:228-242), dropout = identity."""
    qk = src + pos
    a, _ = torch_mha(P, f"{prefix}.self_attn", qk, qk, src, nhead, mask, prec, need_weights=False)
    src = layer_norm(src + a, P[f"{prefix}.norm1.weight"], P[f"{prefix}.norm1.bias"])
    h = prec.s(torch.relu(linear(src, P[f"{prefix}.linear1.weight"], P[f"{prefix}.linear1.bias"], prec)))  # ops.FFNBlockFn: h bf16
    y = linear(h, P[f"{prefix}.linear2.weight"], P[f"{prefix}.linear2.bias"], prec)
    return layer_norm(src + y, P[f"{prefix}.norm2.weight"], P[f"{prefix}.norm2.bias"])


def encoder_forward(P: Params, cfg, vis_features: Tensor, vis_mask: Tensor, durations: Sequence[int],
                    vis_pos: Tensor, text_mask: Tensor, text_memory: Tensor, prec: Prec = FP32,
                    prefix: str = "ground_encoder") -> dict:
    """CrossModalEncoder.forward (modal_encoder.py:40-101) + SpatialTemporalEncoder.forward (:130-204).

    vis_features [n,d,H,W], vis_mask [n,H,W] bool, vis_pos [n,d,H,W], text_mask [b,L] bool (True = pad),
    text_memory [L,b,d].  Returns the ``memory_cache`` dict of :92-99.
    """
    S = cfg.MODEL.STCAT
    d, nhead, nlayers = S.HIDDEN, S.HEADS, S.ENC_LAYERS
    durations = list(durations)
    b, n, t = len(durations), sum(durations), max(durations)
    dt = prec.dtype
    vis_features, vis_pos, text_memory = vis_features.to(dt), vis_pos.to(dt), text_memory.to(dt)
    assert vis_pos.shape[0] == n
    _, _, H, W = vis_features.shape
    vis_mask = vis_mask.clone()
    vis_mask[:, 0, 0] = False  # :46
    x_v = vis_features.flatten(2).permute(2, 0, 1)  # [HW,n,d]  :52
    pos_v = vis_pos.flatten(2).permute(2, 0, 1)
    m_v = vis_mask.flatten(1)  # [n,HW]
    frame_to_video = torch.repeat_interleave(torch.arange(b), torch.tensor(durations))
    m_t = text_mask[frame_to_video]  # [n,L]   :62-68
    x_t = text_memory[:, frame_to_video]  # [L,n,d] :71-77
    x = torch.cat([x_v, x_t], 0)  # :80
    mask = torch.cat([m_v, m_t], 1)  # :81
    pos = torch.cat([pos_v, torch.zeros_like(x_t)], 0)  # :82  text tokens get pos = 0

    ep = f"{prefix}.encoder"
    # SpatialTemporalEncoder.forward :145-159
    frame_cls = P[f"{ep}.frame_cls.weight"].to(dt)  # [1,d]
    x = torch.cat([frame_cls[None].expand(1, n, d), x], 0)  # [S,n,d]
    pos = torch.cat([P[f"{ep}.local_pos_embed.weight"].to(dt)[None].expand(1, n, d), pos], 0)
    kp_mask = torch.cat([torch.zeros(n, 1, dtype=torch.bool), mask], 1)  # [n,S]
    video_src = P[f"{ep}.video_cls.weight"].to(dt).expand(b, d).clone()  # [b,d]
    temp_pos = P[f"{ep}.time_embed.te"].to(dt)[: t + 1].expand(t + 1, b, d)
    temp_mask = torch.ones(b, t + 1, dtype=torch.bool)
    temp_mask[:, 0] = False
    for i, dur in enumerate(durations):
        temp_mask[i, 1:1 + dur] = False
    starts = [0]
    for dur in durations:
        starts.append(starts[-1] + dur)

    for li in range(nlayers):
        x = encoder_layer(P, f"{ep}.spatial_layers.{li}", x, kp_mask, pos, nhead, prec)  # :163-168
        y = torch.zeros(t + 1, b, d, dtype=dt)  # :170-177 (seq-first directly)
        for i, dur in enumerate(durations):
            y[0, i] = video_src[i]
            y[1:1 + dur, i] = x[0, starts[i]:starts[i + 1]]
        y = encoder_layer(P, f"{ep}.temporal_layers.{li}", y, temp_mask, temp_pos, nhead, prec)  # :180-185
        video_src = y[0].clone()  # :191
        cls_new = torch.cat([y[1:1 + dur, i] for i, dur in enumerate(durations)], 0)  # :192-194
        x = torch.cat([cls_new[None], x[1:]], 0)  # :195  (in-place row replacement, restated functionally)

    return {
        "encoded_memory": x[1:],  # [HW+L,n,d]
        "mask": mask,  # [n,HW+L]
        "frames_cls": x[0],  # [n,d]
        "videos_cls": video_src,  # [b,d]
        "durations": durations,
        "fea_map_size": (H, W),
    }


# ----------------------------------------------------------------------------------------------
# decoder  (query_decoder.py)
# ----------------------------------------------------------------------------------------------
def template_generator(P: Params, prefix: str, frames_cls: Tensor, videos_cls: Tensor,
                       durations: Sequence[int], prec: Prec = FP32) -> Tuple[Tensor, Tensor]:
    """TemplateGenerator.forward (query_decoder.py:451-475).  Returns (pos_query [n,4] pre-sigmoid,
    temp_query [n,d])."""
    b = len(durations)
    f2v = torch.repeat_interleave(torch.arange(b), torch.tensor(list(durations)))
    content = linear(videos_cls, P[f"{prefix}.content_proj.weight"], P[f"{prefix}.content_proj.bias"], prec)
    gamma = torch.tanh(linear(videos_cls, P[f"{prefix}.gamma_proj.weight"], P[f"{prefix}.gamma_proj.bias"], prec))
    beta = torch.tanh(linear(videos_cls, P[f"{prefix}.beta_proj.weight"], P[f"{prefix}.beta_proj.bias"], prec))
    pos_query = linear(gamma[f2v] * frames_cls + beta[f2v], P[f"{prefix}.anchor_proj.weight"],
                       P[f"{prefix}.anchor_proj.bias"], prec)
    return pos_query, content[f2v]


def _pad_per_video(x: Tensor, durations: Sequence[int], t: int) -> Tensor:
    """[n, c] -> [t, b, c], zero padded (query_decoder.py:108-119)."""
    out = x.new_zeros(t, len(durations), x.shape[-1])
    s = 0
    for i, dur in enumerate(durations):
        out[:dur, i] = x[s:s + dur]
        s += dur
    return out


def _frames_from_padded(x: Tensor, durations: Sequence[int]) -> Tensor:
    """[t, b, c] -> [1, n, c] (query_decoder.py:386-398, 618-629)."""
    return torch.cat([x[:dur, i] for i, dur in enumerate(durations)], 0)[None]


def _padded_from_frames(x: Tensor, durations: Sequence[int], t: int) -> Tensor:
    """[1, n, c] -> [t, b, c] zero padded (query_decoder.py:419-429, 641-651)."""
    return _pad_per_video(x[0], durations, t)


def box_decoder_layer(P: Params, prefix: str, tgt, memory, query_mask, memory_mask, pos, query_pos,
                      query_time, query_sine, durations, is_first: bool, nhead: int, prec: Prec = FP32,
                      from_scratch: bool = True):
    """TransformerDecoderLayer.forward (query_decoder.py:310-438): the FROM_SCRATCH=True branch (custom attention over the
    per-head [content ; position] concat) and the FROM_SCRATCH=False branch (MDETR-style nn.MultiheadAttention,
    :372-376, 381-384, 409-416)."""
    L = lambda name, x: linear(x, P[f"{prefix}.{name}.weight"], P[f"{prefix}.{name}.bias"], prec)
    t, b, c = tgt.shape
    # self attention over the t queries of each video :329-345.  Kernel-matched mode: each sum of Linears is one
    # multi-term GEMM whose epilogue writes bf16 (decoder.py TransformerDecoderLayer.run: linear_group, out_bf16)
    q = prec.s(L("sa_qcontent_proj", tgt) + L("sa_qtime_proj", query_time) + L("sa_qpos_proj", query_pos))
    k = prec.s(L("sa_kcontent_proj", tgt) + L("sa_ktime_proj", query_time) + L("sa_kpos_proj", query_pos))
    v = prec.s(L("sa_v_proj", tgt))
    a, weights = torch_mha(P, f"{prefix}.self_attn", q, k, v, nhead, query_mask, prec, need_weights=False)
    tgt = layer_norm(tgt + a, P[f"{prefix}.norm1.weight"], P[f"{prefix}.norm1.bias"])
    # time-aligned cross attention :350-429
    n_tok, n, f = memory.shape
    qc = L("ca_qcontent_proj", tgt)
    kc = L("ca_kcontent_proj", memory)
    vv = prec.s(L("ca_v_proj", memory))
    kp = L("ca_kpos_proj", pos)
    if is_first:
        qc = qc + L("ca_qpos_proj", query_pos)
        kc = kc + kp
    dh = c // nhead
    qs = L("ca_qpos_sine_proj", query_sine)
    if from_scratch and prec.km:
        # the CUDA path never builds the concat: two-part score over the bf16 projections (decoder.py memory_side / run)
        o, _ = custom_mha(P, f"{prefix}.cross_attn", _frames_from_padded(prec.s(qc), durations), prec.s(kc), vv, nhead,
                          memory_mask, prec, q2=_frames_from_padded(prec.s(qs), durations), k2=prec.s(kp))
    elif from_scratch:
        q2 = torch.cat([qc.view(t, b, nhead, dh), qs.view(t, b, nhead, dh)], 3).reshape(t, b, 2 * c)
        k2 = torch.cat([kc.view(n_tok, n, nhead, dh), kp.view(n_tok, n, nhead, dh)], 3).reshape(n_tok, n, 2 * c)
        q_cross = _frames_from_padded(q2, durations)  # [1,n,2c]
        o, _ = custom_mha(P, f"{prefix}.cross_attn", q_cross, k2, vv, nhead, memory_mask, prec)
    else:
        # :375-376 q = (q + sine) + ca_qtime_proj(time); :384 k = k + k_pos (a second time in the first layer)
        q1 = (qc + qs) + L("ca_qtime_proj", query_time)
        k1 = prec.s(kc + kp)
        q_cross = _frames_from_padded(q1, durations)  # [1,n,c]
        o, _ = torch_mha(P, f"{prefix}.cross_attn_image", q_cross, k1, vv, nhead, memory_mask, prec, need_weights=False)
    o = _padded_from_frames(o, durations, t)
    tgt = layer_norm(tgt + o, P[f"{prefix}.norm3.weight"], P[f"{prefix}.norm3.bias"])
    # FFN :435-437
    y = L("linear2", prec.s(torch.relu(L("linear1", tgt))))
    tgt = layer_norm(tgt + y, P[f"{prefix}.norm4.weight"], P[f"{prefix}.norm4.bias"])
    return tgt, weights


def box_decoder(P: Params, prefix: str, bbox_prefix: str, tgt, memory, query_mask, memory_mask, pos, anchor,
                query_time, durations, nlayers: int, nhead: int, prec: Prec = FP32, from_scratch: bool = True):
    """TransformerDecoder.forward with bbox_embed set (query_decoder.py:169-247).
    Returns (hs [nl,b,t,d], refs [nl,b,t,4])."""
    d = tgt.shape[-1]
    out = tgt
    inter, refs = [], [anchor]
    for li in range(nlayers):
        sine = anchor_sine_embed(anchor)  # [t,b,512] :191
        query_pos = mlp(P, f"{prefix}.ref_point_head", sine, 2, prec)  # :192
        scale = 1 if li == 0 else mlp(P, f"{prefix}.query_scale", out, 2, prec)  # :195-198
        qsine = sine[..., :d] * scale  # :201
        out, _ = box_decoder_layer(P, f"{prefix}.layers.{li}", out, memory, query_mask, memory_mask, pos,
                                   query_pos, query_time, qsine, durations, li == 0, nhead, prec, from_scratch)
        new_anchor = torch.sigmoid(mlp(P, bbox_prefix, out, 3, prec) + inverse_sigmoid(anchor))  # :213-215
        if li != nlayers - 1:
            refs.append(new_anchor)
        anchor = new_anchor.detach()  # :219
        inter.append(layer_norm(out, P[f"{prefix}.norm.weight"], P[f"{prefix}.norm.bias"]))  # :222
    return torch.stack(inter).transpose(1, 2), torch.stack(refs).transpose(1, 2)


def time_decoder_layer(P: Params, prefix: str, tgt, memory, query_mask, memory_mask, pos, query_pos,
                       query_time_pos, durations, nhead: int, prec: Prec = FP32):
    """TimeDecoderLayer.forward (query_decoder.py:587-660)."""
    t, b, c = tgt.shape
    qk = tgt + (query_pos + query_time_pos)  # :597-599 (q = k = tgt + query_pos + time; the positional sum is formed once)
    a, weights = torch_mha(P, f"{prefix}.self_attn", qk, qk, tgt, nhead, query_mask, prec)
    tgt = layer_norm(tgt + a, P[f"{prefix}.norm1.weight"], P[f"{prefix}.norm1.bias"])
    q_cross = _frames_from_padded(tgt, durations) + _frames_from_padded(query_pos, durations)
    o, _ = torch_mha(P, f"{prefix}.cross_attn_image", q_cross, memory + pos, memory, nhead, memory_mask, prec,
                     need_weights=False)
    o = _padded_from_frames(o, durations, t)
    tgt = layer_norm(tgt + o, P[f"{prefix}.norm3.weight"], P[f"{prefix}.norm3.bias"])
    L = lambda name, x: linear(x, P[f"{prefix}.{name}.weight"], P[f"{prefix}.{name}.bias"], prec)
    y = L("linear2", prec.s(torch.relu(L("linear1", tgt))))
    tgt = layer_norm(tgt + y, P[f"{prefix}.norm4.weight"], P[f"{prefix}.norm4.bias"])
    return tgt, weights


def time_decoder(P: Params, prefix: str, tgt, memory, query_mask, memory_mask, pos, query_pos,
                 query_time_pos, durations, nlayers: int, nhead: int, prec: Prec = FP32):
    """TimeDecoder.forward, return_intermediate & return_weights (query_decoder.py:494-550)."""
    out = tgt
    inter, ws = [], []
    for li in range(nlayers):
        out, w = time_decoder_layer(P, f"{prefix}.layers.{li}", out, memory, query_mask, memory_mask, pos,
                                    query_pos, query_time_pos, durations, nhead, prec)
        inter.append(layer_norm(out, P[f"{prefix}.norm.weight"], P[f"{prefix}.norm.bias"]))
        ws.append(w)
    return torch.stack(inter).transpose(1, 2), torch.stack(ws)


def decoder_forward(P: Params, cfg, memory_cache: dict, vis_pos: Tensor, prec: Prec = FP32,
                    prefix: str = "ground_decoder", bbox_prefix: str = "bbox_embed"):
    """QueryDecoder.forward (query_decoder.py:83-147).  ``text_cls`` is accepted by the reference and
    never used (:451-475), so it is not an argument here."""
    S = cfg.MODEL.STCAT
    d, nhead, nlayers = S.HIDDEN, S.HEADS, S.DEC_LAYERS
    dt = prec.dtype
    memory = memory_cache["encoded_memory"]
    memory_mask = memory_cache["mask"]
    durations = list(memory_cache["durations"])
    H, W = memory_cache["fea_map_size"]
    n_vis = H * W
    b, t = len(durations), max(durations)
    pos_query, temp_query = template_generator(P, f"{prefix}.template_generator", memory_cache["frames_cls"],
                                               memory_cache["videos_cls"], durations, prec)
    anchors = _pad_per_video(torch.sigmoid(pos_query), durations, t)  # [t,b,4]
    query_temporal = _pad_per_video(temp_query, durations, t)  # [t,b,d]
    query_mask = torch.ones(b, t, dtype=torch.bool)
    query_mask[:, 0] = False
    for i, dur in enumerate(durations):
        query_mask[i, :dur] = False
    query_time = P[f"{prefix}.time_embed.te"].to(dt)[:t].expand(t, b, d)
    mem_pos = vis_pos.to(dt).flatten(2).permute(2, 0, 1)
    mem_pos = torch.cat([mem_pos, torch.zeros_like(memory[n_vis:])], 0)
    tgt = torch.zeros(t, b, d, dtype=dt)
    outputs = box_decoder(P, f"{prefix}.decoder", bbox_prefix, tgt, memory, query_mask, memory_mask, mem_pos,
                          anchors, query_time, durations, nlayers, nhead, prec, bool(S.FROM_SCRATCH))
    outputs_temp = time_decoder(P, f"{prefix}.temp_decoder", tgt.clone(), memory, query_mask, memory_mask,
                                mem_pos, query_temporal, query_time, durations, nlayers, nhead, prec)
    return outputs, outputs_temp


# ----------------------------------------------------------------------------------------------
# heads, post-process  (pipeline.py, post_processor.py)
# ----------------------------------------------------------------------------------------------
def heads_forward(P: Params, cfg, outputs, outputs_temp, prec: Prec = FP32) -> dict:
    """The prediction tail of STCATNet.forward (pipeline.py:82-121), eval mode (dropout identity)."""
    hs, reference = outputs
    time_hs, weights = outputs_temp
    out = {}
    if cfg.SOLVER.USE_ATTN:
        out["weights"] = weights[-1]
    coord = torch.sigmoid(mlp(P, "bbox_embed", hs, 3, prec) + inverse_sigmoid(reference)).flatten(1, 2)
    out["pred_boxes"] = coord[-1]
    sted = mlp(P, "temp_embed", time_hs, 2, prec)
    out["pred_sted"] = sted[-1]
    if cfg.MODEL.STCAT.USE_ACTION:
        act = mlp(P, "action_embed", time_hs, 2, prec)
        out["pred_actioness"] = act[-1]
    if cfg.SOLVER.USE_AUX_LOSS:
        out["aux_outputs"] = []
        for i in range(hs.shape[0] - 1):
            a = {"pred_sted": sted[i], "pred_boxes": coord[i]}
            if cfg.SOLVER.USE_ATTN:
                a["weights"] = weights[i]
            if cfg.MODEL.STCAT.USE_ACTION:
                a["pred_actioness"] = act[i]
            out["aux_outputs"].append(a)
    return out


def hot_path_forward(P: Params, cfg, vis_features, vis_mask, durations, vis_pos, text_mask, text_memory,
                     prec: Prec = FP32) -> dict:
    """encoder -> decoder -> heads: everything between ``input_proj``/text encoder and the loss
    (pipeline.py:72-121)."""
    cache = encoder_forward(P, cfg, vis_features, vis_mask, durations, vis_pos, text_mask, text_memory, prec)
    outputs, outputs_temp = decoder_forward(P, cfg, cache, vis_pos, prec)
    out = heads_forward(P, cfg, outputs, outputs_temp, prec)
    out["_memory_cache"] = cache
    out["_hs"], out["_reference"] = outputs
    out["_time_hs"], out["_weights_all"] = outputs_temp
    return out


def post_process(pred_sted: Tensor, pred_boxes: Tensor, target_sizes: Tensor, frames_id, durations):
    """PostProcess.forward (post_processor.py:17-55).

    score[i,j] = logsoftmax_t(sted[:,0])[i] + logsoftmax_t(sted[:,1])[j] - 1e32 * [j<=i or i>=dur or j>=dur]
    (the reference's ``.tril(0)`` masks the diagonal too, :37); flat argmax -> (start, end+1) frame ids.
    Returns (boxes_xyxy_scaled [n,4], steds list, score_map [b,t,t]).
    """
    cx, cy, w, h = pred_boxes.unbind(-1)
    boxes = torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1)
    img_h, img_w = target_sizes.unbind(1)
    boxes = (boxes * torch.stack([img_w, img_h, img_w, img_h], 1)).clamp(min=0)
    b, t, _ = pred_sted.shape
    neg = -1e32
    ii = torch.arange(t)[:, None]
    jj = torch.arange(t)[None, :]
    maps = []
    for i_b in range(b):
        dur = durations[i_b]
        m = torch.zeros(t, t, dtype=pred_sted.dtype)
        m[(jj <= ii) | (ii >= dur) | (jj >= dur)] = neg
        maps.append(m)
    score = torch.stack(maps) + F.log_softmax(pred_sted[:, :, 0], dim=1)[:, :, None] \
        + F.log_softmax(pred_sted[:, :, 1], dim=1)[:, None, :]
    steds = []
    for i_b in range(b):
        idx = int(score[i_b].flatten().argmax())
        s, e = idx // t, idx % t
        steds.append([frames_id[i_b][s], frames_id[i_b][e] + 1])
    return boxes, steds, score


# ----------------------------------------------------------------------------------------------
# loss (criterion.py) -- downstream of the hot path; restated so fwd+bwd has the reference's entry
# gradient.  Used by bench.py's CPU baseline and by gradient-parity tests.
# ----------------------------------------------------------------------------------------------
def _box_cxcywh_to_xyxy(x):
    cx, cy, w, h = x.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1)


def _giou_diag(b1, b2):
    """diag of generalized_box_iou (utils/box_utils.py:94-115) computed pairwise-aligned."""
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    lt = torch.max(b1[:, :2], b2[:, :2])
    rb = torch.min(b1[:, 2:], b2[:, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[:, 0] * wh[:, 1]
    union = a1 + a2 - inter
    iou = inter / union
    lt2 = torch.min(b1[:, :2], b2[:, :2])
    rb2 = torch.max(b1[:, 2:], b2[:, 2:])
    wh2 = (rb2 - lt2).clamp(min=0)
    area = wh2[:, 0] * wh2[:, 1]
    return iou - (area - union) / area


def stg_loss(cfg, out: dict, target_boxes: Tensor, actioness: Tensor, durations: Sequence[int],
             weight_dict: Optional[dict] = None) -> Tuple[Tensor, dict]:
    """VideoSTGLoss.forward (criterion.py:151-207) for b videos, world size 1.

    target_boxes: [num_gt_frames_total, 4] cxcywh for the frames with actioness==1 (video-major order);
    actioness: [b, t] {0,1}.  Returns (weighted total, dict of the individual losses)."""
    S = cfg.SOLVER
    b, t = actioness.shape
    dev = out["pred_boxes"].device
    bounds, sl = [], []
    for i in range(b):
        idx = torch.nonzero(actioness[i].cpu()).flatten().tolist()
        bounds.append((idx[0], idx[-1]))
        sl.extend(range(i * t + idx[0], i * t + idx[-1] + 1))
    sl = torch.tensor(sl, dtype=torch.long, device=dev)
    num_boxes = max(float(target_boxes.shape[0]), 1.0)
    time_mask = torch.zeros(b, t, dtype=torch.bool, device=dev)
    for i, dur in enumerate(durations):
        time_mask[i, :dur] = True
    positive = torch.zeros(b, t, dtype=torch.bool, device=dev)
    for i, (s, e) in enumerate(bounds):
        positive[i, s:e + 1] = True
    eps = 1e-6
    ar = torch.arange(t, device=dev)[None, :]
    tstart = torch.tensor([x[0] for x in bounds], device=dev)[:, None]
    tend = torch.tensor([x[1] for x in bounds], device=dev)[:, None]

    def one(o):
        L = {}
        pb = o["pred_boxes"][sl]
        L["loss_bbox"] = (pb - target_boxes).abs().sum() / num_boxes  # criterion.py:26-39
        L["loss_giou"] = (1 - _giou_diag(_box_cxcywh_to_xyxy(pb), _box_cxcywh_to_xyxy(target_boxes))).sum() / num_boxes
        sted = o["pred_sted"].masked_fill(~time_mask[:, :, None], -1e32)  # :64-109
        tot = 0
        for ch, tt in ((0, tstart), (1, tend)):
            distrib = F.normalize((-((ar - tt) ** 2) / (2 * S.SIGMA ** 2)).exp() + eps, p=1, dim=1)
            prob = sted[:, :, ch].softmax(1)
            tot = tot + prob * ((prob + eps) / distrib).log() * time_mask
        L["loss_sted"] = tot.mean()
        if S.USE_ATTN:  # :111-130
            w = o["weights"]
            pm = positive | (~time_mask)
            la = -(1 - w + eps).log()
            la = la.masked_fill(pm[:, :, None], 0)
            nb_neg = (~pm).sum(1) + eps
            L["loss_guided_attn"] = (la.sum(2) / nb_neg[:, None]).sum(1).mean()
        if cfg.MODEL.STCAT.USE_ACTION:  # :46-62
            pa = o["pred_actioness"].squeeze(-1)
            wgt = torch.full(pa.shape, float(S.EOS_COEF), device=dev, dtype=pa.dtype)
            for i, (s, e) in enumerate(bounds):
                wgt[i, s:e + 1] = 1
            la = F.binary_cross_entropy_with_logits(pa, actioness.to(pa.dtype), weight=wgt, reduction="none")
            L["loss_actioness"] = (la * time_mask).mean()
        return L

    losses = one(out)
    for i, aux in enumerate(out.get("aux_outputs", [])):
        losses.update({f"{k}_{i}": v for k, v in one(aux).items()})
    if weight_dict is None:
        weight_dict = loss_weight_dict(cfg)
    total = sum(losses[k] * weight_dict[k] for k in losses if k in weight_dict)
    return total, losses


def loss_weight_dict(cfg) -> dict:
    """build_model's weight_dict (models/__init__.py:12-29)."""
    S = cfg.SOLVER
    wd = {"loss_bbox": S.BBOX_COEF, "loss_giou": S.GIOU_COEF, "loss_sted": S.TEMP_COEF}
    if cfg.MODEL.STCAT.USE_ACTION:
        wd["loss_actioness"] = S.ACTIONESS_COEF
    if S.USE_ATTN:
        wd["loss_guided_attn"] = S.ATTN_COEF
    if S.USE_AUX_LOSS:
        base = dict(wd)
        for i in range(cfg.MODEL.STCAT.DEC_LAYERS - 1):
            wd.update({f"{k}_{i}": v for k, v in base.items()})
    return wd


# ----------------------------------------------------------------------------------------------
# map2d_head.py (orphaned in the reference, named by north_star): the 2-D proposal map + conv head
# ----------------------------------------------------------------------------------------------
def map2d_masks(N: int, pooling_counts: Sequence[int]):
    """Gen2DMap.__init__ (map2d_head.py:11-37): valid-cell mask and the (i, j) index lists of every
    super-diagonal, plus the pooling schedule as (kernel, stride) pairs."""
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


def gen_2d_map(x: Tensor, N: int, pooling_counts: Sequence[int]) -> Tensor:
    """Gen2DMap.forward (map2d_head.py:39-62).  x [B,T,d] -> map2d [B,d,N,N]."""
    mask2d, maskij, poolers = map2d_masks(N, pooling_counts)
    x = x.permute(0, 2, 1)
    if x.shape[-1] > N:
        x = F.adaptive_avg_pool1d(x, N)
    x = F.adaptive_max_pool1d(x, N)
    B, d, _ = x.shape
    m = x.new_zeros(B, d, N, N)
    m[:, :, range(N), range(N)] = x
    for (kk, ss), (i, j) in zip(poolers, maskij):
        x = F.max_pool1d(x, kk, ss)
        m[:, :, i, j] = x
    return m


def map2d_conv_weights(mask2d: Tensor, k: int, num_layers: int) -> List[Tensor]:
    """mask2weight cascade of TempConvInteraction.__init__ (map2d_head.py:221-245)."""
    kernel = torch.ones(1, 1, k, k)
    first_padding = (k - 1) * num_layers // 2

    def m2w(m, padding):
        w = torch.conv2d(m[None, None].float(), kernel, padding=padding)[0, 0]
        w[w > 0] = 1 / w[w > 0]
        return w

    ws = [m2w(mask2d, first_padding)]
    for _ in range(num_layers - 1):
        ws.append(m2w(ws[-1] > 0, 0))
    return ws


def map2d_conv_head(P: Params, prefix: str, x: Tensor, cfg_stcat, training: bool = False,
                    prec: Prec = FP32) -> Tensor:
    """TempPredictionHead.forward with TEMP_HEAD='conv' (map2d_head.py:105-127, 228-250).
    x [layers,b,T,d] -> scores [layers,b,N,N]."""
    N, counts = cfg_stcat.MAX_MAP_SIZE, cfg_stcat.POOLING_COUNTS
    k, nconv = cfg_stcat.KERNAL_SIZE, cfg_stcat.CONV_LAYERS
    nl, b, t, d = x.shape
    mask2d, _, _ = map2d_masks(N, counts)
    m = gen_2d_map(x.reshape(-1, t, d), N, counts)
    ws = map2d_conv_weights(mask2d, k, nconv)
    first_padding = (k - 1) * nconv // 2
    for i in range(nconv):
        m = F.conv2d(prec.r(m), prec.r(P[f"{prefix}.encoder.convs.{i}.weight"]), P[f"{prefix}.encoder.convs.{i}.bias"],
                     padding=first_padding if i == 0 else 0).relu() * ws[i].to(m.dtype)
    s = F.conv2d(m, P[f"{prefix}.predictor.weight"], P[f"{prefix}.predictor.bias"]).squeeze(1)
    s = s.view(nl, b, N, N)
    return s if training else torch.sigmoid(s) * mask2d


def map2d_attn_head(P: Params, prefix: str, x: Tensor, cfg_stcat, training: bool = False, prec: Prec = FP32) -> Tensor:
    """TempPredictionHead.forward with TEMP_HEAD='attn' (map2d_head.py:105-127, 151-205), dropout = identity.
    Per map [d,N,N] -> [N,N,d]; per layer: row attention (sequence = first axis, batch = second, key_padding_mask =
    mask2d AS IS: the reference passes the *valid* mask where torch expects True = ignore, and indexes it [batch, key],
    i.e. transposed w.r.t. the map -- both quirks are reproduced), column attention on the row output with the transposed
    mask, residual + norm1, FFN + norm2.  x [layers,b,T,d] -> scores [layers,b,N,N]."""
    N, counts = cfg_stcat.MAX_MAP_SIZE, cfg_stcat.POOLING_COUNTS
    nhead, nlayers = cfg_stcat.HEADS, cfg_stcat.TEMP_PRED_LAYERS
    nl, b, t, d = x.shape
    mask2d, _, _ = map2d_masks(N, counts)
    maps = gen_2d_map(x.reshape(-1, t, d), N, counts)  # [nl*b, d, N, N]
    outs = []
    for m in maps:
        src = m.permute(1, 2, 0)  # [N,N,d]
        for li in range(nlayers):
            pf = f"{prefix}.encoder.layers.{li}"
            a, _ = torch_mha(P, f"{pf}.self_attn_row", src, src, src, nhead, mask2d, prec)          # :181-185
            a = a.permute(1, 0, 2)                                                                      # :188
            a, _ = torch_mha(P, f"{pf}.self_attn_col", a, a, a, nhead, mask2d.permute(1, 0), prec)     # :189-193
            a = a.permute(1, 0, 2)                                                                      # :194
            src = layer_norm(src + a, P[f"{pf}.norm1.weight"], P[f"{pf}.norm1.bias"])
            h = torch.relu(linear(src, P[f"{pf}.linear1.weight"], P[f"{pf}.linear1.bias"], prec))
            y = linear(h, P[f"{pf}.linear2.weight"], P[f"{pf}.linear2.bias"], prec)
            src = layer_norm(src + y, P[f"{pf}.norm2.weight"], P[f"{pf}.norm2.bias"])
        outs.append(src.permute(2, 0, 1))
    m = torch.stack(outs)
    s = F.conv2d(m, P[f"{prefix}.predictor.weight"], P[f"{prefix}.predictor.bias"]).squeeze(1)
    s = s.view(nl, b, N, N)
    return s if training else torch.sigmoid(s) * mask2d


# ----------------------------------------------------------------------------------------------
# either side of the hot path
# ----------------------------------------------------------------------------------------------
def input_proj(P: Params, prefix: str, feats: Tensor) -> Tensor:
    """``nn.Conv2d(2048, 256, kernel_size=1)`` (pipeline.py:41,64): feats [n, C, H, W] -> [n, 256, H, W]."""
    w = P[f"{prefix}.weight"]
    y = torch.einsum("nchw,oc->nohw", feats, w.view(w.shape[0], -1))
    return y + P[f"{prefix}.bias"].view(1, -1, 1, 1)


def feature_resizer(P: Params, prefix: str, x: Tensor) -> Tensor:
    """FeatureResizer.forward in eval mode (language_model/bert.py:91-110): Linear -> LayerNorm(eps 1e-12)."""
    y = x @ P[f"{prefix}.fc.weight"].t() + P[f"{prefix}.fc.bias"]
    return layer_norm(y, P[f"{prefix}.layer_norm.weight"], P[f"{prefix}.layer_norm.bias"], eps=1e-12)


def linear_interp(bbox_dict: dict) -> dict:
    """engine/evaluate.py:20-38: fill every frame id between two predicted ones with the linear blend of their boxes."""
    frame_ids = sorted(bbox_dict)
    if len(frame_ids) < 2:
        return bbox_dict
    out = dict(bbox_dict)
    for left, right in zip(frame_ids[:-1], frame_ids[1:]):
        interval = right - left
        if interval > 1:
            bl, br = bbox_dict[left][0], bbox_dict[right][0]
            delta = [(br[c] - bl[c]) / interval for c in range(4)]
            for step in range(1, interval):
                out[left + step] = [[bl[c] + step * delta[c] for c in range(4)]]
    return {fid: out[fid] for fid in sorted(out)}


def merge_even_odd(pred1, sted1, pred2, sted2):
    """engine/evaluate.py:112-121: boxes of the even and the odd pass merged and interpolated, union of the two segments."""
    boxes = dict(pred1)
    boxes.update(pred2)
    return linear_interp(boxes), [min(sted1[0], sted2[0]), max(sted1[1], sted2[1])]
