"""B200 cross-modal encoder: drop-in for the reference ``build_encoder(cfg)``.

Interface mirrored (reference models/grounding_model/modal_encoder.py):
  ``CrossModalEncoder(cfg).forward(videos: NestedTensor, vis_pos, texts) -> memory_cache`` (:40-101),
  attribute ``d_model``, and the state_dict keys of SURVEY.md 8b (``encoder.spatial_layers.N.*``,
  ``encoder.temporal_layers.N.*``, ``encoder.{time_embed.te, local_pos_embed, frame_cls, video_cls}``,
  ``fusion.*``).

Data layout (differs from the reference on purpose).  The reference keeps tokens sequence-first
``[S, n_frames, d]`` so that every frame's tokens are strided by ``n_frames*d``.  Here the residual
stream is one frame-major matrix ``X [n_frames * S, d]`` (row = frame * S + token): the per-frame
spatial attention then reads contiguous ``S x 32`` head slices, all GEMMs see a plain row-major
``[N_s, 256]`` activation, and the decoder's time-aligned cross attention (one query per frame) finds
the keys of its frame in one contiguous block.  The views handed back in ``memory_cache`` have the
reference's shapes (``encoded_memory [HW+L, n, d]`` etc.).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import os

import torch
from torch import nn

from . import ops
from .params import LinearP, MHAP, NormP, SineTable, LearnedTable, TokenP, xavier_reset


def _check_cfg(cfg):
    S = cfg.MODEL.STCAT
    if S.HIDDEN != 256 or S.HEADS != 8:
        raise NotImplementedError(
            f"stcat_b200 kernels are built for MODEL.STCAT.HIDDEN=256, HEADS=8 (head dim 32); got {S.HIDDEN}/{S.HEADS}")


_index_cache = {}
# fused layout glue (ops.token_assembly / mem_operands / template: csrc/assembly.cu) in bf16 mode; STCAT_FUSED_GLUE=0 keeps the
# torch.cat / slice composition (the only path in exact-fp32 mode)
_FUSED_GLUE = os.environ.get("STCAT_FUSED_GLUE", "1") != "0"
_EARLY_QKV = os.environ.get("STCAT_EARLY_QKV", "1") != "0"  # ops.qkv_early: next layer's in-projection under the temporal layer
_CLS_KERNELS = os.environ.get("STCAT_CLS_KERNELS", "1") != "0"  # ops.cls_gather / cls_scatter with fused operand copies (A/B switch)


def set_fused_glue(on: bool):
    global _FUSED_GLUE
    _FUSED_GLUE = bool(on)


def fused_glue() -> bool:
    return _FUSED_GLUE and ops.get_precision() == "bf16"


def batch_indices(durations: Sequence[int], device) -> dict:
    """Index / mask tensors that map the frame-concatenated layout (n = sum(durations) rows) to the
    padded per-video layout ([b, t] rows) and back.  Cached per (durations, device) so that the steady
    state does no host->device copies (the reference does 43 per forward)."""
    key = (tuple(durations), str(device))
    hit = _index_cache.get(key)
    if hit is not None:
        return hit
    durations = list(durations)
    b, n, t = len(durations), sum(durations), max(durations)
    f2v, f2i = [], []
    for j, dur in enumerate(durations):
        f2v += [j] * dur
        f2i += list(range(dur))
    f2v_t = torch.tensor(f2v, dtype=torch.long)
    f2i_t = torch.tensor(f2i, dtype=torch.long)
    # temporal sequence of the encoder: [b, t+1] rows; source rows are [video_src (b) ; frame cls (n) ; zero (1)]
    enc_gather = torch.full((b, t + 1), b + n, dtype=torch.long)
    enc_gather[:, 0] = torch.arange(b)
    enc_gather[f2v_t, 1 + f2i_t] = b + torch.arange(n)
    enc_scatter = f2v_t * (t + 1) + 1 + f2i_t  # frame f <- row of the [b*(t+1)] temporal output
    temp_mask = torch.ones(b, t + 1, dtype=torch.uint8)
    temp_mask[:, 0] = 0
    temp_mask[f2v_t, 1 + f2i_t] = 0
    # decoder queries: [b, t] rows; source rows are [frames (n) ; zero (1)]
    dec_gather = torch.full((b, t), n, dtype=torch.long)
    dec_gather[f2v_t, f2i_t] = torch.arange(n)
    dec_scatter = f2v_t * t + f2i_t
    query_mask = torch.ones(b, t, dtype=torch.uint8)
    query_mask[:, 0] = 0
    query_mask[f2v_t, f2i_t] = 0
    vid_start = torch.tensor([0] + [sum(durations[: j + 1]) for j in range(b)], dtype=torch.long)
    out = {
        "b": b, "n": n, "t": t, "identity": b == 1, "vid_start": vid_start.to(device),
        "f2v": f2v_t.to(device), "enc_gather": enc_gather.flatten().to(device), "enc_scatter": enc_scatter.to(device),
        "temp_mask": temp_mask.to(device), "dec_gather": dec_gather.flatten().to(device),
        "dec_scatter": dec_scatter.to(device), "query_mask": query_mask.to(device),
    }
    _index_cache[key] = out
    return out


class TransformerEncoderLayer(nn.Module):
    """Parameters of one post-norm encoder layer (modal_encoder.py:207-242)."""

    def __init__(self, d_model: int, nhead: int, dim_feedforward: int, dropout: float):
        super().__init__()
        self.self_attn = MHAP(d_model, nhead)
        self.linear1 = LinearP(d_model, dim_feedforward)
        self.linear2 = LinearP(dim_feedforward, d_model)
        self.norm1 = NormP(d_model)
        self.norm2 = NormP(d_model)
        self.nhead = nhead
        self.dropout_p = dropout

    def run(self, x, x_op, pos, key_mask, B: int, L: int, pos_cls=None, qk_op=None, pre=None):
        """x, pos: [B*L, d] batch-major rows.  Returns (y, y_op).  In train mode the four dropout sites of the reference
        layer (attention probabilities, dropout1, the FFN's inner dropout, dropout2; modal_encoder.py:212-241) are active."""
        a = self.self_attn
        p = self.dropout_p if self.training else 0.0
        x, x_op = ops.self_attn_block(x, x_op, pos, key_mask, a.in_proj_weight, a.in_proj_bias, a.out_proj.weight,
                                      a.out_proj.bias, self.norm1.weight, self.norm1.bias, B, L, self.nhead,
                                      self.norm1.eps, pos_cls=pos_cls, drop_p=p, qk_op=qk_op, pre=pre)
        return ops.ffn_block(x, x_op, self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias,
                             self.norm2.weight, self.norm2.bias, self.norm2.eps, drop_p=p)


class SpatialTemporalEncoder(nn.Module):
    """6 x [spatial layer per frame -> temporal layer over the frame CLS tokens] (modal_encoder.py:104-204)."""

    def __init__(self, cfg, num_layers: int):
        super().__init__()
        S = cfg.MODEL.STCAT
        d = S.HIDDEN
        mk = lambda: TransformerEncoderLayer(d, S.HEADS, S.FFN_DIM, S.DROPOUT)
        self.spatial_layers = nn.ModuleList(mk() for _ in range(num_layers))
        self.temporal_layers = nn.ModuleList(mk() for _ in range(num_layers))
        max_len = cfg.INPUT.MAX_VIDEO_LEN + 1
        self.time_embed = LearnedTable(max_len, d) if S.USE_LEARN_TIME_EMBED else SineTable(max_len, d)
        self.local_pos_embed = TokenP(d)
        self.frame_cls = TokenP(d)
        self.video_cls = TokenP(d)
        self.num_layers = num_layers
        self.d_model = d

    def run(self, X, POS, key_mask, n: int, S_len: int, durations, pos_cls=None, first_ops=None):
        """X, POS: [n*S, d] frame-major (row 0 of every frame = CLS slot).  Returns (X, video_src [b, d]).
        ``pos_cls``: see ops.self_attn_block (POS is then a constant whose CLS rows equal this parameter).
        ``first_ops``: (bf16(X + POS), bf16(X)) when the token assembly already wrote the first layer's GEMM operands."""
        d = self.d_model
        idx = batch_indices(durations, X.device)
        b, t = idx["b"], idx["t"]
        max_t = self.time_embed.rows(t + 1).shape[0]
        if t + 1 > max_t:
            raise ValueError(f"clip of {t} frames exceeds INPUT.MAX_VIDEO_LEN={max_t - 1}")
        video_src = self.video_cls.weight.expand(b, d)
        temp_pos = self.time_embed.rows(t + 1)  # [t+1, d]
        temp_pos = temp_pos if b == 1 else temp_pos.repeat(b, 1)
        temp_pos = temp_pos.contiguous()
        qk_op, X_op = first_ops if first_ops is not None else (None, None)
        pre = None
        # optional observer of every block's input (dp.GradSync hangs its bucketed all-reduce on their gradients);
        # nothing is stored here: holding these tensors would keep the step's autograd graph alive
        on_input = getattr(self, "layer_input_callback", None)
        for li, (sp, tp) in enumerate(zip(self.spatial_layers, self.temporal_layers)):
            if on_input is not None:
                on_input(li, X)
            X, X_op = sp.run(X, X_op, POS, key_mask, n, S_len, pos_cls=pos_cls, qk_op=qk_op if li == 0 else None, pre=pre)
            pre = None
            if idx["identity"]:
                # one un-padded video: the frame-CLS exchange with the temporal layer as two row-sized nodes
                # Y = [video token ; frame-CLS rows]; in bf16 mode the same launch writes the temporal layer's GEMM operands and
                # the scatter refreshes the CLS rows of the stream's operand copy (3 launches fewer per block on the chain)
                fuse = _CLS_KERNELS and fused_glue()
                if (fuse and _EARLY_QKV and X_op is not None and pos_cls is not None and li + 1 < self.num_layers
                        and POS.dtype == torch.float32):
                    # the next spatial layer's q/k/v projection of every row the temporal layer does not touch, on a side stream
                    nxt = self.spatial_layers[li + 1].self_attn
                    pre = ops.qkv_early(X, X_op, POS, nxt.in_proj_weight, nxt.in_proj_bias)
                X3, Y, tq_op, ty_op = ops.cls_gather(X.view(n, S_len, d), video_src, 0, pos=temp_pos if fuse else None)
                Y, _ = tp.run(Y, ty_op, temp_pos, idx["temp_mask"], b, t + 1, qk_op=tq_op)
                X3, video_src = ops.cls_scatter(X3, Y, 0, x_op=X_op if fuse else None, pre=pre, pos=POS)  # modal_encoder.py:191-195
                cls_new = None if (fuse and X_op is not None) else Y.detach()[1:]
            else:
                X3, cls = ops.take_rows(X.view(n, S_len, d), 0)  # frame-CLS rows (copy) + the stream itself
                Y = torch.cat([video_src, cls, cls.new_zeros(1, d)], 0).index_select(0, idx["enc_gather"])
                Y, _ = tp.run(Y, None, temp_pos, idx["temp_mask"], b, t + 1)
                video_src = Y.view(b, t + 1, d)[:, 0, :]
                cls_new = Y.index_select(0, idx["enc_scatter"])
                X3 = ops.put_rows(X3, cls_new, 0)  # modal_encoder.py:191-195
            X = X3.view(n * S_len, d)
            if X_op is not None and cls_new is not None:
                X_op.view(n, S_len, d)[:, 0, :] = cls_new.detach()
        return X, video_src


class CrossModalEncoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        _check_cfg(cfg)
        S = cfg.MODEL.STCAT
        self.d_model = S.HIDDEN
        self.dropout_p = float(S.DROPOUT)
        self.encoder = SpatialTemporalEncoder(cfg, S.ENC_LAYERS)
        self.fusion = LinearP(S.HIDDEN, S.HIDDEN)  # never used by the reference forward; checkpoint contract
        xavier_reset(self)

    def forward(self, videos=None, vis_pos: Optional[torch.Tensor] = None, texts: Optional[Tuple] = None) -> dict:
        vis_features, vis_mask, durations = videos.decompose()
        durations = list(durations)
        n, d, H, W = vis_features.shape
        if vis_pos.shape[0] != sum(durations) or n != sum(durations):
            raise AssertionError("vis_pos / videos frame count does not match sum(durations)")
        vis_mask[:, 0, 0] = False  # the reference mutates the caller's mask too (modal_encoder.py:46)
        text_mask, text_memory, _ = texts  # [b, L] bool (True = pad), [L, b, d]
        b = len(durations)
        assert text_memory.shape[1] == b and text_mask.shape[0] == b
        L = text_memory.shape[0]
        HW = H * W
        S_len = 1 + HW + L
        idx = batch_indices(durations, vis_features.device)
        enc = self.encoder
        pos_const = not (vis_pos.requires_grad and torch.is_grad_enabled())
        mask = None
        if fused_glue() and pos_const and L > 0 and all(t_.dtype == torch.float32 for t_ in (vis_features, vis_pos, text_memory)):
            # one launch: X, POS and the first layer's operand copies (ops.token_assembly); backward: one launch
            m_t = text_mask.expand(n, L) if idx["identity"] else text_mask.index_select(0, idx["f2v"])
            mask = torch.cat([vis_mask.flatten(1), m_t], 1)  # [n, HW+L] bool (returned)
            key_mask = torch.cat([mask.new_zeros(n, 1), mask], 1).to(torch.uint8).contiguous()  # [n, S]
            X3, POS3, qk_op, x_op = ops.token_assembly(vis_features, vis_pos, text_memory, enc.frame_cls.weight,
                                                       enc.local_pos_embed.weight.detach(),
                                                       None if idx["identity"] else idx["f2v"],
                                                       None if idx["identity"] else idx["vid_start"])
            X, video_src = enc.run(X3.view(n * S_len, d), POS3.view(n * S_len, d), key_mask, n, S_len, durations,
                                   pos_cls=enc.local_pos_embed.weight, first_ops=(qk_op.view(n * S_len, d), x_op.view(n * S_len, d)))
            X3 = X.view(n, S_len, d)
            return {
                "encoded_memory": X3[:, 1:, :].transpose(0, 1), "mask": mask, "frames_cls": X3[:, 0, :], "videos_cls": video_src,
                "durations": durations, "fea_map_size": (H, W),
                # private: the frame-major stream and its positional stream for this package's decoder (ops.mem_operands)
                "_stream": (X3, POS3, vis_pos.data_ptr()),
            }
        # ---- token assembly, frame-major: [cls ; HW visual tokens ; L text tokens] per frame ----
        x_v = vis_features.flatten(2).transpose(1, 2)  # [n, HW, d] view
        p_v = vis_pos.flatten(2).transpose(1, 2)
        x_t = text_memory.transpose(0, 1)  # [b, L, d]
        if idx["identity"]:
            x_t = x_t.expand(n, L, d)
            m_t = text_mask.expand(n, L)
        else:
            x_t = x_t.index_select(0, idx["f2v"])
            m_t = text_mask.index_select(0, idx["f2v"])
        X = torch.cat([enc.frame_cls.weight.expand(n, 1, d), x_v, x_t], 1).reshape(n * S_len, d)
        # the positional stream is constant except its CLS row (local_pos_embed, a parameter): when vis_pos needs no
        # gradient (it is the backbone's sine embedding) the big tensor is built detached and the parameter's gradient
        # is taken from the CLS rows directly (ops.self_attn_block, pos_cls)
        pos_cls = None if (vis_pos.requires_grad and torch.is_grad_enabled()) else enc.local_pos_embed.weight
        lpe = enc.local_pos_embed.weight if pos_cls is None else enc.local_pos_embed.weight.detach()
        POS = torch.cat([lpe.expand(n, 1, d), p_v, p_v.new_zeros(n, L, d)], 1).reshape(n * S_len, d)
        mask = torch.cat([vis_mask.flatten(1), m_t], 1)  # [n, HW+L] bool (returned)
        key_mask = torch.cat([mask.new_zeros(n, 1), mask], 1).to(torch.uint8).contiguous()  # [n, S]
        X = X.float()
        POS = POS.float()
        X, video_src = enc.run(X, POS, key_mask, n, S_len, durations, pos_cls=pos_cls)
        X3 = X.view(n, S_len, d)
        return {
            "encoded_memory": X3[:, 1:, :].transpose(0, 1),  # [HW+L, n, d] (view of the frame-major stream)
            "mask": mask,
            "frames_cls": X3[:, 0, :],  # [n, d]
            "videos_cls": video_src,  # [b, d]
            "durations": durations,
            "fea_map_size": (H, W),
        }


def build_encoder(cfg) -> CrossModalEncoder:
    """Mirror of models/grounding_model/__init__.py:5-6."""
    return CrossModalEncoder(cfg)
