"""The checkpoint (state_dict) contract of the hot path, as data.

Key names and shapes that ``engine/optimizer.py:26-33`` (LR groups by substring),
``utils/checkpoint.py:122-201`` (MDETR remap, strict load) and the published checkpoints depend on
(probe p5).  ``tests/test_boundary.py`` checks the modules of this package against this
table, and ``oracle/make_golden.py`` checks the table against the reference's own ``state_dict()``.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict


def _linear(out: OrderedDict, prefix: str, n_out: int, n_in: int):
    out[f"{prefix}.weight"] = (n_out, n_in)
    out[f"{prefix}.bias"] = (n_out,)


def _norm(out: OrderedDict, prefix: str, d: int):
    out[f"{prefix}.weight"] = (d,)
    out[f"{prefix}.bias"] = (d,)


def _mha(out: OrderedDict, prefix: str, d: int):
    out[f"{prefix}.in_proj_weight"] = (3 * d, d)
    out[f"{prefix}.in_proj_bias"] = (3 * d,)
    _linear(out, f"{prefix}.out_proj", d, d)


def _mlp(out: OrderedDict, prefix: str, n_in: int, hidden: int, n_out: int, layers: int):
    dims = [n_in] + [hidden] * (layers - 1) + [n_out]
    for i in range(layers):
        _linear(out, f"{prefix}.layers.{i}", dims[i + 1], dims[i])


def encoder_spec(cfg, prefix: str = "ground_encoder") -> "OrderedDict[str, Tuple[int, ...]]":
    S = cfg.MODEL.STCAT
    d, F, nl = S.HIDDEN, S.FFN_DIM, S.ENC_LAYERS
    out: OrderedDict = OrderedDict()
    for kind in ("spatial_layers", "temporal_layers"):
        for i in range(nl):
            p = f"{prefix}.encoder.{kind}.{i}"
            _mha(out, f"{p}.self_attn", d)
            _linear(out, f"{p}.linear1", F, d)
            _linear(out, f"{p}.linear2", d, F)
            _norm(out, f"{p}.norm1", d)
            _norm(out, f"{p}.norm2", d)
    out[f"{prefix}.encoder.time_embed.te"] = (cfg.INPUT.MAX_VIDEO_LEN + 1, 1, d)  # buffer
    out[f"{prefix}.encoder.local_pos_embed.weight"] = (1, d)
    out[f"{prefix}.encoder.frame_cls.weight"] = (1, d)
    out[f"{prefix}.encoder.video_cls.weight"] = (1, d)
    _linear(out, f"{prefix}.fusion", d, d)  # never used in forward; checkpoint contract
    return out


def decoder_spec(cfg, prefix: str = "ground_decoder", with_bbox_alias: bool = True):
    S = cfg.MODEL.STCAT
    d, F, nl, qd = S.HIDDEN, S.FFN_DIM, S.DEC_LAYERS, S.QUERY_DIM
    out: OrderedDict = OrderedDict()
    tg = f"{prefix}.template_generator"
    for nm in ("content_proj", "gamma_proj", "beta_proj"):
        _linear(out, f"{tg}.{nm}", d, d)
    _linear(out, f"{tg}.anchor_proj", qd, d)
    for i in range(nl):
        p = f"{prefix}.decoder.layers.{i}"
        for nm in ("sa_qcontent_proj", "sa_qpos_proj", "sa_qtime_proj", "sa_kcontent_proj", "sa_kpos_proj",
                   "sa_ktime_proj", "sa_v_proj"):
            _linear(out, f"{p}.{nm}", d, d)
        _mha(out, f"{p}.self_attn", d)
        _linear(out, f"{p}.ca_qcontent_proj", d, d)
        if i == 0:
            _linear(out, f"{p}.ca_qpos_proj", d, d)  # layers >= 1 set it to None (query_decoder.py:166-167)
        for nm in ("ca_kcontent_proj", "ca_kpos_proj", "ca_qtime_proj", "ca_v_proj", "ca_qpos_sine_proj"):
            _linear(out, f"{p}.{nm}", d, d)
        if S.FROM_SCRATCH:  # custom MHA without in-projections (query_decoder.py:285-286)
            _linear(out, f"{p}.cross_attn.out_proj", d, d)
        else:               # MDETR-initialised nn.MultiheadAttention (query_decoder.py:287-288)
            _mha(out, f"{p}.cross_attn_image", d)
        _linear(out, f"{p}.linear1", F, d)
        _linear(out, f"{p}.linear2", d, F)
        for nm in ("norm1", "norm3", "norm4"):
            _norm(out, f"{p}.{nm}", d)
    _norm(out, f"{prefix}.decoder.norm", d)
    _mlp(out, f"{prefix}.decoder.query_scale", d, d, d, 2)
    _mlp(out, f"{prefix}.decoder.ref_point_head", qd // 2 * d, d, d, 2)
    if with_bbox_alias:  # assigned from outside (pipeline.py:50); aliases the top-level bbox_embed
        _mlp(out, f"{prefix}.decoder.bbox_embed", d, d, 4, 3)
    for i in range(nl):
        p = f"{prefix}.temp_decoder.layers.{i}"
        _mha(out, f"{p}.self_attn", d)
        _mha(out, f"{p}.cross_attn_image", d)
        _linear(out, f"{p}.linear1", F, d)
        _linear(out, f"{p}.linear2", d, F)
        for nm in ("norm1", "norm3", "norm4"):
            _norm(out, f"{p}.{nm}", d)
    _norm(out, f"{prefix}.temp_decoder.norm", d)
    out[f"{prefix}.time_embed.te"] = (cfg.INPUT.MAX_VIDEO_LEN + 1, 1, d)  # buffer
    return out


def heads_spec(cfg):
    d = cfg.MODEL.STCAT.HIDDEN
    out: OrderedDict = OrderedDict()
    _mlp(out, "temp_embed", d, d, 2, 2)
    _mlp(out, "bbox_embed", d, d, 4, 3)
    if cfg.MODEL.STCAT.USE_ACTION:
        _mlp(out, "action_embed", d, d, 1, 2)
    return out


def hot_path_spec(cfg) -> "OrderedDict[str, Tuple[int, ...]]":
    out: OrderedDict = OrderedDict()
    out.update(encoder_spec(cfg))
    out.update(decoder_spec(cfg))
    out.update(heads_spec(cfg))
    return out


def synthetic_params(cfg, seed: int = 0) -> Dict[str, "torch.Tensor"]:
    """{name: fp32 CPU tensor} for the whole hot path from the deterministic recipe; buffers are built
    by their defining formula."""
    import torch  # noqa: F401
    from . import posenc
    from .synthetic import fill_param

    out = {}
    for k, shape in hot_path_spec(cfg).items():
        if k.endswith(".te"):
            out[k] = posenc.seq_sine_table(shape[0], shape[2])
        else:
            out[k] = fill_param(k, shape, seed)
    return out
