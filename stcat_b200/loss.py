"""Device-side restatement of the loss that enters the hot path's backward (reference
models/criterion.py:26-207 ``VideoSTGLoss``), batched over the 6 decoder layers and free of host
synchronisation.

The loss is *downstream* of the hot path: O(T) scalars on small tensors.
It is expressed with torch tensor ops on the device (no custom kernels yet); what matters for the
hot path is that (a) it produces the reference's entry gradient for fwd+bwd, and (b) unlike the
reference (``.cpu()``, ``.item()``, ``torch.LongTensor(...).to(device)`` per step, criterion.py:163-178)
every target-dependent constant is built once per target set (``STGLossPlan``) so the step itself
never blocks the host.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch
import torch.nn.functional as F


def loss_weight_dict(cfg) -> Dict[str, float]:
    """``build_model``'s weight_dict (models/__init__.py:12-29)."""
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


def _xyxy(x):
    cx, cy, w, h = x.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1)


def _giou_aligned(b1, b2):
    """diag(generalized_box_iou(b1, b2)) (utils/box_utils.py:94-115), computed pair-aligned."""
    a1 = (b1[..., 2] - b1[..., 0]) * (b1[..., 3] - b1[..., 1])
    a2 = (b2[..., 2] - b2[..., 0]) * (b2[..., 3] - b2[..., 1])
    wh = (torch.min(b1[..., 2:], b2[..., 2:]) - torch.max(b1[..., :2], b2[..., :2])).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    union = a1 + a2 - inter
    wh2 = (torch.max(b1[..., 2:], b2[..., 2:]) - torch.min(b1[..., :2], b2[..., :2])).clamp(min=0)
    area = wh2[..., 0] * wh2[..., 1]
    return inter / union - (area - union) / area


class STGLossPlan:
    """Everything of ``VideoSTGLoss.forward`` that depends only on the targets, precomputed on the
    device: the gt frame slice, the Gaussian start/end distributions, masks and BCE weights."""

    def __init__(self, cfg, target_boxes: torch.Tensor, actioness: torch.Tensor, durations: Sequence[int], device,
                 world_size: int = 1, num_boxes_global: Optional[float] = None):
        S = cfg.SOLVER
        self.cfg = cfg
        self.use_attn = bool(S.USE_ATTN)
        self.use_action = bool(cfg.MODEL.STCAT.USE_ACTION)
        self.use_aux = bool(S.USE_AUX_LOSS)
        act = actioness.detach().cpu()
        b, t = act.shape
        bounds, sl = [], []
        for i in range(b):
            nz = torch.nonzero(act[i]).flatten().tolist()
            bounds.append((nz[0], nz[-1]))
            sl.extend(range(i * t + nz[0], i * t + nz[-1] + 1))
        # num_boxes averaged over ranks (criterion.py:172-178): passed in by the caller for DDP
        nb = float(target_boxes.shape[0]) if num_boxes_global is None else float(num_boxes_global) / world_size
        self.num_boxes = max(nb, 1.0)
        self.slice = torch.tensor(sl, dtype=torch.long).to(device)
        self.target_boxes = target_boxes.to(device=device, dtype=torch.float32)
        self.target_xyxy = _xyxy(self.target_boxes)
        time_mask = torch.zeros(b, t, dtype=torch.bool)
        positive = torch.zeros(b, t, dtype=torch.bool)
        weight = torch.full((b, t), float(S.EOS_COEF))
        for i, dur in enumerate(durations):
            time_mask[i, :dur] = True
        for i, (s, e) in enumerate(bounds):
            positive[i, s:e + 1] = True
            weight[i, s:e + 1] = 1
        ar = torch.arange(t)[None, :]
        eps = 1e-6
        ds = []
        for col in (0, 1):
            tt = torch.tensor([x[col] for x in bounds])[:, None]
            ds.append(F.normalize((-((ar - tt) ** 2) / (2 * S.SIGMA ** 2)).exp() + eps, p=1, dim=1))
        self.distrib = torch.stack(ds, -1).to(device)  # [b, t, 2]
        self.time_mask = time_mask.to(device)
        self.time_mask_u8 = time_mask.to(torch.uint8).to(device)
        self.time_mask_f = time_mask.float().to(device)
        pm = positive | (~time_mask)
        self.neg_f = (~pm).float().to(device)  # [b, t]
        self.nb_neg = ((~pm).sum(1).float() + eps).to(device)
        self.bce_weight = weight.to(device)
        self.actioness = act.float().to(device)
        self.weights = loss_weight_dict(cfg)
        self.b, self.t = b, t

    ORDER = ("loss_bbox", "loss_giou", "loss_sted", "loss_guided_attn", "loss_actioness")

    def _fused(self, out: dict):
        """CUDA path: one kernel for the values and the gradients of all layers (csrc/stg_loss.cu)."""
        nl = out["_hs"].shape[0]
        first = 0 if self.use_aux else nl - 1
        coord = out["_coord_all"][first:]
        sted = out["_sted_all"][first:]
        act = out["_act_all"][first:] if self.use_action else None
        attn = out["_weights_all"][first:] if self.use_attn else None
        coef = [self.weights["loss_bbox"], self.weights["loss_giou"], self.weights["loss_sted"],
                self.weights.get("loss_guided_attn", 0.0) if self.use_attn else 0.0,
                self.weights.get("loss_actioness", 0.0) if self.use_action else 0.0]
        total, losses = FusedSTGLossFn.apply(self, tuple(coef), coord, sted, act, attn)
        named = {}
        k = losses.shape[0]
        for ci, name in enumerate(self.ORDER):
            if (name == "loss_guided_attn" and not self.use_attn) or (name == "loss_actioness" and not self.use_action):
                continue
            named[name] = losses[k - 1, ci]
            if self.use_aux:
                for i in range(k - 1):
                    named[f"{name}_{i}"] = losses[i, ci]
        return total, named

    def __call__(self, out: dict):
        """out: the dict of STCATHotPath.forward.  Returns (total, {reference loss name: scalar tensor})."""
        if out["_coord_all"].is_cuda:
            return self._fused(out)
        return self.torch_restatement(out)

    def torch_restatement(self, out: dict):
        """The same loss with torch tensor ops: used on CPU by the tests of the host-side composition (no CUDA device in
        the build container) and as the on-device checker of the fused kernel (tests/test_gpu_hotpath.py)."""
        eps = 1e-6
        nl = out["_hs"].shape[0]
        coord = out["_coord_all"]  # [nl, b*t, 4]
        sted = out["_sted_all"]  # [nl, b, t, 2]
        layers = range(nl) if self.use_aux else [nl - 1]
        first = 0 if self.use_aux else nl - 1
        pb = coord[first:, self.slice]  # [k, K, 4]
        per = {}
        per["loss_bbox"] = (pb - self.target_boxes).abs().sum((1, 2)) / self.num_boxes
        per["loss_giou"] = (1 - _giou_aligned(_xyxy(pb), self.target_xyxy)).sum(1) / self.num_boxes
        s = sted[first:].masked_fill(~self.time_mask[None, :, :, None], -1e32)
        prob = s.softmax(2)
        kl = prob * ((prob + eps) / self.distrib).log() * self.time_mask_f[None, :, :, None]
        per["loss_sted"] = kl.sum(3).mean((1, 2))
        if self.use_attn:
            w = out["_weights_all"][first:]  # [k, b, t, t]
            la = -(1 - w + eps).log() * self.neg_f[None, :, :, None]
            per["loss_guided_attn"] = (la.sum(3) / self.nb_neg[None, :, None]).sum(2).mean(1)
        if self.use_action:
            pa = out["_act_all"][first:].squeeze(-1)  # [k, b, t]
            la = F.binary_cross_entropy_with_logits(pa, self.actioness.expand_as(pa), weight=self.bce_weight.expand_as(pa),
                                                    reduction="none")
            per["loss_actioness"] = (la * self.time_mask_f).mean((1, 2))
        total = 0
        named = {}
        for k, v in per.items():
            total = total + self.weights[k] * v.sum()
            named[k] = v[-1]
            if self.use_aux:
                for i in range(nl - 1):
                    named[f"{k}_{i}"] = v[i]
        return total, named


class FusedSTGLossFn(torch.autograd.Function):
    """total = sum_l sum_k coef_k loss_{l,k}: values and gradients from ONE kernel launch (stcat_stg_loss); backward just
    scales the stored gradients by the incoming scalar."""

    @staticmethod
    def forward(ctx, plan, coef, coord, sted, act, attn):
        from . import ops

        be = ops.get_backend()
        f32 = torch.float32
        c = coord.detach().to(f32).contiguous()
        s = sted.detach().to(f32).contiguous()
        a = None if act is None else act.detach().to(f32).contiguous()
        w = None if attn is None else attn.detach().to(f32).contiguous()
        k = c.shape[0]
        losses = torch.empty(k, 5, dtype=f32, device=c.device)
        dc, ds = torch.empty_like(c), torch.empty_like(s)
        da = None if a is None else torch.empty_like(a)
        dw = None if w is None else torch.empty_like(w)
        be.stg_loss(c, s, None if a is None else a.view(k, plan.b, plan.t), w, plan, coef, losses, dc, ds, da, dw)
        cvec = torch.tensor(coef, dtype=f32).to(c.device, non_blocking=True) if not hasattr(plan, "_coef_dev") else plan._coef_dev
        plan._coef_dev = cvec
        total = (losses * cvec).sum()
        ctx.save_for_backward(dc, ds, da, dw)
        ctx.mark_non_differentiable(losses)
        return total, losses

    @staticmethod
    def backward(ctx, g, _unused):
        dc, ds, da, dw = ctx.saved_tensors
        return (None, None, dc * g, ds * g, None if da is None else da * g, None if dw is None else dw * g)
