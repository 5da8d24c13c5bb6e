"""The hot-path slice of ``STCATNet`` (reference models/pipeline.py:12-121) as one module.

``STCATHotPath`` owns exactly the sub-modules of ``STCATNet`` that lie on the hot path, under the same
attribute names (``ground_encoder``, ``ground_decoder``, ``bbox_embed``, ``temp_embed``,
``action_embed``), so a reference checkpoint's keys for these sub-modules load into it unchanged.
Its ``forward`` starts at the seam where the reference hands over to ``ground_encoder``
(pipeline.py:72: ``input_proj``-ed visual features as a NestedTensor, the backbone's positional
embedding and the text encoder's output tuple) and returns the output dict of pipeline.py:82-121.

Inside the reference itself the drop-in is one level lower: ``build_encoder`` / ``build_decoder``
(see INTEGRATION.md); this class is what ``bench.py`` and the tests drive.
"""
from __future__ import annotations

from typing import Sequence

import torch
from torch import nn

from . import ops
from .decoder import box_refine, build_decoder, run_mlp
from .encoder import build_encoder
from .params import MLPP


class STCATHotPath(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg.clone() if hasattr(cfg, "clone") else cfg
        self.use_attn = cfg.SOLVER.USE_ATTN
        self.use_aux_loss = cfg.SOLVER.USE_AUX_LOSS
        self.use_actioness = cfg.MODEL.STCAT.USE_ACTION
        self.query_dim = cfg.MODEL.STCAT.QUERY_DIM
        d = cfg.MODEL.STCAT.HIDDEN
        self.ground_encoder = build_encoder(cfg)
        self.ground_decoder = build_decoder(cfg)
        self.temp_embed = MLPP(d, d, 2, 2, dropout=0.3)
        self.bbox_embed = MLPP(d, d, 4, 3)
        self.action_embed = MLPP(d, d, 1, 2, dropout=0.3) if self.use_actioness else None
        self.ground_decoder.decoder.bbox_embed = self.bbox_embed  # iterative anchor update (pipeline.py:50)

    def heads(self, outputs, outputs_temp) -> dict:
        """pipeline.py:82-121."""
        out = {}
        time_hs, weights = outputs_temp
        if self.use_attn:
            out["weights"] = weights[-1]
        hs, reference = outputs
        coord = box_refine(run_mlp(self.bbox_embed, hs, self.training), reference).flatten(1, 2)
        out["pred_boxes"] = coord[-1]
        ths_op = ops.operand_copy(time_hs)  # one operand cast for both temporal heads
        sted = run_mlp(self.temp_embed, time_hs, self.training, x_op=ths_op)
        out["pred_sted"] = sted[-1]
        act = None
        if self.use_actioness:
            act = run_mlp(self.action_embed, time_hs, self.training, x_op=ths_op)
            out["pred_actioness"] = act[-1]
        out["_coord_all"], out["_sted_all"], out["_act_all"] = coord, sted, act  # stacked over layers (loss.py)
        if self.use_aux_loss:
            aux = []
            for i in range(hs.shape[0] - 1):
                a = {"pred_sted": sted[i], "pred_boxes": coord[i]}
                if self.use_attn:
                    a["weights"] = weights[i]
                if self.use_actioness:
                    a["pred_actioness"] = act[i]
                aux.append(a)
            out["aux_outputs"] = aux
        return out

    def forward(self, videos, vis_pos: torch.Tensor, texts) -> dict:
        """videos: NestedTensor(vis_features [n,256,H,W], mask [n,H,W] bool, durations);
        vis_pos [n,256,H,W]; texts = (text_mask [b,L] bool, text_memory [L,b,256], tokenized-or-None)."""
        cache = self.ground_encoder(videos=videos, vis_pos=vis_pos, texts=texts)
        outputs, outputs_temp = self.ground_decoder(memory_cache=cache, vis_pos=vis_pos, text_cls=None)
        out = self.heads(outputs, outputs_temp)
        out["_memory_cache"] = cache
        out["_hs"], out["_reference"] = outputs
        out["_time_hs"], out["_weights_all"] = outputs_temp
        return out

    def load_flat_params(self, P: dict, strict: bool = True):
        """Load a flat {reference state_dict key: tensor} dict (e.g. ``param_spec.synthetic_params``)."""
        sd = self.state_dict()
        missing = [k for k in sd if k not in P]
        extra = [k for k in P if k not in sd]
        if strict and (missing or extra):
            raise KeyError(f"state_dict mismatch: missing {missing[:5]} unexpected {extra[:5]}")
        with torch.no_grad():
            for k, v in sd.items():
                if k in P:
                    v.copy_(P[k].to(v.device, v.dtype))
        ops.clear_weight_cache()
        return self


class STCATNet(STCATHotPath):
    """The whole of the reference's ``STCATNet`` (models/pipeline.py:12-121) at its OUTER seam: ``model(videos, texts)`` with
    ``videos`` a NestedTensor of frames [sum(T), 3, H, W] and ``texts`` a list of captions (or pre-tokenised tensors), the same
    output dict, the same ``state_dict`` names (``vis_encoder.0.body.*``, ``text_encoder.body.*``, ``text_encoder.resizer.*``,
    ``input_proj.*`` and the hot-path modules).  The backbone trunk and the language model are library code (torchvision /
    cuDNN convolutions with FrozenBatchNorm folded, Hugging Face RobertaModel), as in the reference; ``input_proj``, the
    positional encoding, the resizer and everything after them run on this package's C ABI.

    ``text_body`` / ``tokenizer``: pass a constructed ``RobertaModel`` and tokenizer to build offline (no checkpoint
    download); default loads ``cfg.MODEL.TEXT_MODEL.NAME`` like the reference."""

    def __init__(self, cfg, text_body=None, tokenizer=None):
        super().__init__(cfg)
        from .text import TextEncoder
        from .vision import InputProj, VisionEncoder

        self.vis_encoder = VisionEncoder(cfg)
        self.text_encoder = TextEncoder(cfg.MODEL.TEXT_MODEL.NAME, cfg.MODEL.STCAT.HIDDEN, bool(cfg.MODEL.TEXT_MODEL.FREEZE),
                                        body=text_body, tokenizer=tokenizer)
        self.input_proj = InputProj(self.vis_encoder.num_channels, cfg.MODEL.STCAT.HIDDEN)

    def forward(self, videos, texts, logger=None) -> dict:  # pipeline.py:52-121
        from .nested import NestedTensor

        vis_outputs, vis_pos = self.vis_encoder(videos)
        feats, vis_mask, durations = vis_outputs.decompose()
        feats = self.input_proj(feats)
        text_outputs, _text_cls = self.text_encoder(texts, feats.device)
        out = super().forward(NestedTensor(feats, vis_mask, durations), vis_pos, text_outputs)
        return out


class PostProcess(nn.Module):
    """Mirror of models/post_processor.py:17-55 with the T x T start/end scoring on the device."""

    @torch.no_grad()
    def forward(self, outputs: dict, target_sizes: torch.Tensor, frames_id: Sequence[Sequence[int]], durations=None):
        pred_boxes, pred_sted = outputs["pred_boxes"], outputs["pred_sted"]
        b, t, _ = pred_sted.shape
        if durations is None:
            durations = [len(f) for f in frames_id]
        cx, cy, w, h = pred_boxes.unbind(-1)
        boxes = torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1)
        img_h, img_w = target_sizes.unbind(1)
        boxes = (boxes * torch.stack([img_w, img_h, img_w, img_h], 1)).clamp(min=0)
        best, _ = ops.sted_score(pred_sted, durations)
        best = best.cpu().tolist()  # the one host sync of the eval path: frame ids are python ints
        steds = []
        for i_b, flat in enumerate(best):
            s, e = flat // t, flat % t
            steds.append([frames_id[i_b][s], frames_id[i_b][e] + 1])
        return boxes, steds
