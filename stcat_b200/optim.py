"""Optimizer-side step of the training loop for the hot path's parameters, fused:

    torch.nn.utils.clip_grad_norm_(model.parameters(), MAX_GRAD_NORM)     train_net.py:139-140
    torch.optim.AdamW(param_list, lr, weight_decay).step()                engine/optimizer.py:36-46
    update_ema(model, model_ema, EMA_DECAY)                                engine/optimizer.py:5-22
    (+ the refresh of the bf16 GEMM-operand shadows of the weights)

The reference does these tensor by tensor from Python (three passes over every parameter plus ~4 kernels per tensor).  Here
parameters, gradients (``dp.FlatGrads``), both Adam moments, the EMA copy and the bf16 shadows live in flat buffers with
one shared layout, and a step is: one ``stcat_sumsq`` launch for the gradient norm (no host sync: the clip coefficient is
formed on the device) and one ``stcat_adamw_step`` launch per run of parameters that share (lr, weight_decay).

``FusedAdamW`` is a ``torch.optim.Optimizer``: ``param_groups`` (the reference's LR schedule writes ``group["lr"]``,
``engine/lr_scheduler.py`` / ``train_net.py:142``), ``state_dict`` / ``load_state_dict`` and ``zero_grad`` keep their
meaning; the moments are exposed per parameter as ``exp_avg`` / ``exp_avg_sq`` views like torch's.
"""
from __future__ import annotations

import bisect
from typing import Iterable, Optional

import torch

from . import ops
from .dp import FlatGrads


class FusedAdamW(torch.optim.Optimizer):
    """AdamW over the parameters of a ``FlatGrads`` layout.

    flat          : the gradient layout; every parameter of ``param_groups`` must belong to it
    param_groups  : torch-style list of dicts ({"params": [...], "lr": ..., "weight_decay": ...}) or a list of parameters
    max_grad_norm : > 0 clips by the global L2 norm of all gradients of ``flat`` plus those of ``extra_norm_params``
                    (parameters optimised elsewhere that the reference's single clip_grad_norm_ call also covers)
    ema_decay     : keeps ``ema = ema * decay + (1 - decay) * p`` in the same pass; ``attach_ema(model, model_ema)``
                    makes a deep-copied EMA model's parameters views of it
    frozen        : parameters that never receive a gradient (the reference leaves ``.grad`` None for them, so torch's AdamW
                    skips them: no weight decay, no moments); they are excluded from the fused ranges.  MANDATORY for the
                    parameters the forward never uses (``ground_encoder.fusion.*``; ``ca_qtime_proj.*`` and
                    ``cross_attn_image.*`` of the box decoder when FROM_SCRATCH is True): their flat gradient slice is
                    all zero, which -- unlike torch's ``grad is None`` -- would still apply weight decay.
                    ``unused_parameters(model, cfg)`` lists them.
    grad_scale    : the gradients are multiplied by this before the norm / update (1/world_size after a SUM all-reduce;
                    ``dp.GradSync`` averages by default, so 1)
    extra_norm_params : their gradients enter the global norm AND are scaled in place by the clip coefficient in ``step()``
                    (the reference's one ``clip_grad_norm_(model.parameters())`` call scales every gradient), so whatever
                    optimises them afterwards sees clipped gradients; ``clip_coef()`` exposes the coefficient of the last step.
    """

    def __init__(self, flat: FlatGrads, param_groups, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, max_grad_norm: float = 0.0, ema_decay: Optional[float] = None,
                 frozen: Iterable[torch.nn.Parameter] = (), extra_norm_params: Iterable[torch.nn.Parameter] = (),
                 grad_scale: float = 1.0):
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        super().__init__(param_groups, defaults)
        self.flat = flat
        self.max_grad_norm = float(max_grad_norm)
        self.grad_scale = float(grad_scale)
        self.ema_decay = ema_decay
        self.extra_norm_params = list(extra_norm_params)
        self._frozen = {id(p) for p in frozen}
        self._start = {id(p): st for p, st in zip(flat.params, flat.starts)}
        for g in self.param_groups:
            for p in g["params"]:
                if id(p) not in self._start:
                    raise ValueError("FusedAdamW: a parameter of param_groups is not part of the FlatGrads layout")
        dev = flat.buf.device
        n = flat.numel
        # fp32 masters: every parameter becomes a view of one buffer laid out like the gradients
        self.pbuf = torch.zeros(n, dtype=torch.float32, device=dev)
        for p, st in zip(flat.params, flat.starts):
            view = self.pbuf[st:st + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.ema = self.pbuf.clone() if ema_decay is not None else None
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.shadow = None
        self._steps = 0
        for g in self.param_groups:
            for p in g["params"]:
                st = self._start[id(p)]
                self.state[p] = {"step": torch.tensor(0.0), "exp_avg": self.m[st:st + p.numel()].view_as(p),
                                 "exp_avg_sq": self.v[st:st + p.numel()].view_as(p)}
        self._runs = None

    # -- layout ----------------------------------------------------------------------------------------------------
    def _padded(self, p):
        A = FlatGrads.ALIGN
        return (p.numel() + A - 1) // A * A

    def _build_runs(self):
        """per param group: maximal contiguous [lo, hi) ranges of its non-frozen parameters (alignment gaps between
        neighbours are zeros in every buffer and stay zeros under the update, so runs may span them)"""
        runs = []
        for gi, g in enumerate(self.param_groups):
            spans = sorted((self._start[id(p)], self._start[id(p)] + self._padded(p)) for p in g["params"] if id(p) not in self._frozen)
            merged = []
            for lo, hi in spans:
                if merged and merged[-1][1] == lo:
                    merged[-1][1] = hi
                else:
                    merged.append([lo, hi])
            runs.append([(lo, min(hi, self.flat.numel)) for lo, hi in merged])
        return runs

    # -- bf16 shadows ----------------------------------------------------------------------------------------------
    def enable_shadows(self):
        """Keep bf16 copies of all parameters, refreshed by the step kernel, and serve them to ``ops`` as the GEMM operands
        of the weights (bf16 mode): no cast kernels after an optimizer step."""
        self.shadow = self.pbuf.to(torch.bfloat16)
        base, end = self.pbuf.data_ptr(), self.pbuf.data_ptr() + 4 * self.pbuf.numel()
        starts = list(self.flat.starts)
        params = self.flat.params
        # The step kernel writes masters and shadows through raw pointers (no autograd version bump), so a parameter whose
        # ``_version`` moved since the snapshot was written by something else -- load_state_dict (checkpoint resume happens
        # AFTER the optimizer is built, train_net.py:60-75), load_flat_params, copying EMA weights in for evaluation --
        # and its shadow is re-cast before it is served.
        self._pver = [p._version for p in params]

        def provider(t):
            ptr = t.data_ptr()
            if t.dtype != torch.float32 or not (base <= ptr < end) or not t.is_contiguous():
                return None
            off = (ptr - base) // 4
            i = bisect.bisect_right(starts, off) - 1
            if t._version != self._pver[i]:  # views / detach() share the parameter's version counter
                n = params[i].numel()
                self.shadow[starts[i]:starts[i] + n].copy_(self.pbuf[starts[i]:starts[i] + n])
                self._pver[i] = t._version
            return self.shadow.as_strided(tuple(t.shape), tuple(t.stride()), off)

        ops.set_shadow_provider(provider)
        return self

    def refresh_shadows(self):
        """Re-cast every bf16 shadow from the fp32 masters (after any bulk write that is not ``step()``; the provider also
        detects such writes per parameter through the autograd version counter)."""
        if self.shadow is not None:
            self.shadow.copy_(self.pbuf)
            self._pver = [p._version for p in self.flat.params]
        else:
            ops.clear_weight_cache()
        return self

    def clip_coef(self) -> torch.Tensor:
        """Device scalar: the factor the last ``step()`` applied to every gradient (grad_scale x the global-norm clip)."""
        c = torch.full((1,), self.grad_scale, dtype=torch.float32, device=self.sumsq.device)
        if self.max_grad_norm > 0:  # sumsq is that of the already scaled gradients
            c = c * torch.clamp(self.max_grad_norm / (self.sumsq.sqrt() + 1e-6), max=1.0)
        return c

    # -- EMA -------------------------------------------------------------------------------------------------------
    def attach_ema(self, model, model_ema):
        """Make the parameters of ``model_ema`` (a deepcopy of ``model``, train_net.py:28) views of the EMA buffer.  Buffers of
        the EMA model (none on the hot path besides constant tables) are left alone."""
        assert self.ema is not None, "construct FusedAdamW with ema_decay"
        ema_params = dict(model_ema.named_parameters())
        for name, p in model.named_parameters():
            st = self._start.get(id(p))
            if st is None or name not in ema_params:
                continue
            q = ema_params[name]
            view = self.ema[st:st + p.numel()].view_as(p)
            view.copy_(q.data)
            q.data = view

    # -- step ------------------------------------------------------------------------------------------------------
    def zero_grad(self, set_to_none: bool = False):
        self.flat.zero()

    @torch.no_grad()
    def step(self, closure=None):
        assert closure is None
        be = ops.get_backend()
        ops.join_leaf_streams()  # weight gradients issued on leaf streams (ops.set_leaf_streams)
        if self._runs is None:
            self._runs = self._build_runs()
        self._steps += 1
        clip = self.max_grad_norm > 0
        if self.grad_scale != 1.0:
            self.flat.buf.mul_(self.grad_scale)
            for q in self.extra_norm_params:
                if q.grad is not None:
                    q.grad.mul_(self.grad_scale)
        if clip:
            self.sumsq.zero_()
            be.sumsq(self.flat.buf, self.sumsq)
            for q in self.extra_norm_params:
                if q.grad is not None:
                    self.sumsq.add_(q.grad.detach().float().pow(2).sum())
            if self.extra_norm_params:
                coef = torch.clamp(self.max_grad_norm / (self.sumsq.sqrt() + 1e-6), max=1.0)
                for q in self.extra_norm_params:
                    if q.grad is not None:
                        q.grad.mul_(coef.to(q.grad.dtype))
        for g, runs in zip(self.param_groups, self._runs):
            b1, b2 = g["betas"]
            for lo, hi in runs:
                be.adamw_step(self.pbuf[lo:hi], self.flat.buf[lo:hi], self.m[lo:hi], self.v[lo:hi],
                              None if self.ema is None else self.ema[lo:hi], None if self.shadow is None else self.shadow[lo:hi],
                              g["lr"], b1, b2, g["eps"], g["weight_decay"], self._steps, self.sumsq if clip else None,
                              self.max_grad_norm, 0.0 if self.ema_decay is None else self.ema_decay)
        for st in self.state.values():
            st["step"] += 1
        if self.shadow is None:
            # the masters moved under the version-keyed cast cache of ops (raw-pointer update, no version bump)
            ops.clear_weight_cache()
        return None

    # -- checkpoints -----------------------------------------------------------------------------------------------
    def load_state_dict(self, state_dict):
        """torch's loader replaces the state tensors; copy them back into the flat buffers and re-expose the views."""
        super().load_state_dict(state_dict)
        steps = 0
        for g in self.param_groups:
            for p in g["params"]:
                st = self.state.get(p)
                if not st:
                    continue
                o = self._start[id(p)]
                mv = self.m[o:o + p.numel()].view_as(p)
                vv = self.v[o:o + p.numel()].view_as(p)
                mv.copy_(st["exp_avg"])
                vv.copy_(st["exp_avg_sq"])
                st["exp_avg"], st["exp_avg_sq"] = mv, vv
                steps = max(steps, int(st["step"]))
        self._steps = steps
        self.refresh_shadows()


def unused_parameters(model, cfg=None):
    """Parameters of the hot path that its forward never touches (the reason the reference needs
    ``find_unused_parameters=True``): pass them as ``frozen`` so they get neither weight decay nor moments, like torch's
    AdamW skipping ``grad is None``."""
    from_scratch = True if cfg is None else bool(cfg.MODEL.STCAT.FROM_SCRATCH)
    out = []
    for name, p in model.named_parameters():
        if ".fusion." in name or name.startswith("fusion."):
            out.append(p)
        elif "decoder.layers." in name and "temp_decoder" not in name:
            if from_scratch and (".ca_qtime_proj." in name or ".cross_attn_image." in name):
                out.append(p)
            elif not from_scratch and ".cross_attn.out_proj." in name:
                out.append(p)
    return out


def make_optimizer(cfg, model, flat: FlatGrads, enable_shadows: bool = True) -> FusedAdamW:
    """The hot-path part of the reference's ``make_optimizer`` (engine/optimizer.py:25-58) + ``clip_grad_norm_``
    (train_net.py:139-140) + ``update_ema`` (:143) as one FusedAdamW: parameters under ``ground_decoder.temp_decoder`` at
    SOLVER.TEMP_LR, every other hot-path parameter at SOLVER.BASE_LR, SOLVER.WEIGHT_DECAY, SOLVER.MAX_GRAD_NORM,
    MODEL.EMA_DECAY.  (``vis_encoder`` / ``text_encoder`` parameters are outside the hot path; hand them to a torch
    optimizer and list them as ``extra_norm_params`` so the clip covers them.)"""
    temp = [p for n, p in model.named_parameters() if "ground_decoder.temp_decoder" in n and p.requires_grad]
    rest = [p for n, p in model.named_parameters() if "ground_decoder.temp_decoder" not in n and p.requires_grad]
    groups = [{"params": rest}, {"params": temp, "lr": float(cfg.SOLVER.TEMP_LR)}]
    opt = FusedAdamW(flat, groups, lr=float(cfg.SOLVER.BASE_LR), weight_decay=float(cfg.SOLVER.WEIGHT_DECAY),
                     max_grad_norm=float(cfg.SOLVER.MAX_GRAD_NORM),
                     ema_decay=float(cfg.MODEL.EMA_DECAY) if cfg.MODEL.EMA else None, frozen=unused_parameters(model, cfg))
    if enable_shadows and ops.get_precision() == "bf16":
        opt.enable_shadows()
    return opt
