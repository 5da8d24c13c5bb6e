"""Visual side of ``STCATNet`` in front of the hot path: reference models/vision_model/backbone.py:16-159,
position_encoding.py:70-94 and the ``input_proj`` 1x1 convolution of models/pipeline.py:41,62-65.

* ``InputProj``: the 2048 -> 256 1x1 convolution as ONE GEMM of this package's C ABI (tcgen05 in bf16 mode) over the
  channels-last feature map: rows = (frame, h, w), so the output [n * HW, 256] is already the token-major layout the
  encoder's assembly copies from (the reference materialises [n, 256, H, W] and transposes it per forward).  Same
  ``state_dict`` entries as ``nn.Conv2d(2048, 256, 1)`` (weight [256, 2048, 1, 1], bias [256]).
* ``PositionEmbeddingSine``: one kernel (``stcat_pos_sine``) from the padding mask, written channels-last; cached per mask
  shape when the mask has no padding (the benchmark's and most training clips' case: a constant table).
* ``FrozenBNResNet``: the torchvision ResNet trunk the reference uses (library code: cuDNN convolutions through torch, NOT a
  kernel of this package), with what the survey asks for around it: FrozenBatchNorm folded into the convolution weights
  (no separate scale/shift pass per conv), channels-last bf16 activations, the same ``state_dict`` names as the reference's
  ``vis_encoder.0.body.*`` (torchvision ResNet + FrozenBatchNorm2d buffers), the same ``requires_grad`` policy.
* ``VisionEncoder``: ``Joiner``-shaped: ``forward(NestedTensor frames) -> (NestedTensor features, pos)``.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .nested import NestedTensor


class InputProj(nn.Module):
    """pipeline.py:41 ``nn.Conv2d(vis_fea_dim, hidden_dim, kernel_size=1)`` as a GEMM over channels-last rows."""

    def __init__(self, in_channels: int = 2048, hidden: int = 256):
        super().__init__()
        self.in_channels, self.hidden = in_channels, hidden
        self.weight = nn.Parameter(torch.empty(hidden, in_channels, 1, 1))
        self.bias = nn.Parameter(torch.empty(hidden))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))  # nn.Conv2d.reset_parameters
        bound = 1 / math.sqrt(in_channels)
        nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, feats: torch.Tensor) -> torch.Tensor:
        """feats [n, C, H, W] (any memory format; channels-last is read in place) -> [n, hidden, H, W] fp32 as the permuted
        view of the token-major GEMM output (``.flatten(2).transpose(1, 2)`` of it is contiguous)."""
        n, C, H, W = feats.shape
        x2 = feats.permute(0, 2, 3, 1).reshape(n * H * W, C)  # view for channels-last input, one transpose copy otherwise
        y = ops.linear(x2, self.weight.view(self.hidden, C), self.bias)
        return y.view(n, H, W, self.hidden).permute(0, 3, 1, 2)


class PositionEmbeddingSine(nn.Module):
    """position_encoding.py:52-94 (num_pos_feats = hidden / 2, temperature 10000, normalize=True, scale 2 pi)."""

    def __init__(self, num_pos_feats: int = 128, temperature: float = 10000.0, scale: float = 2 * math.pi):
        super().__init__()
        self.num_pos_feats, self.temperature, self.scale = num_pos_feats, temperature, scale
        self._cache = {}

    @torch.no_grad()
    def forward(self, tensor_list: NestedTensor) -> torch.Tensor:
        mask = tensor_list.mask
        n, H, W = mask.shape
        key = None
        if not bool(mask.any()):  # no padding: the table depends on (H, W) only; one frame's worth, expanded
            key = (H, W, str(mask.device))
            hit = self._cache.get(key)
            if hit is not None:
                return hit.expand(n, -1, -1, -1)
            mask = mask[:1]
        out = torch.empty(mask.shape[0], H, W, 2 * self.num_pos_feats, dtype=torch.float32, device=mask.device)
        ops.get_backend().pos_sine(mask.to(torch.uint8).contiguous(), out, self.num_pos_feats, self.temperature, self.scale)
        pos = out.permute(0, 3, 1, 2)  # [n, 2F, H, W] view of the channels-last buffer
        if key is not None:
            self._cache[key] = pos
            return pos.expand(n, -1, -1, -1)
        return pos


class FrozenBatchNorm2d(nn.Module):
    """backbone.py:16-66: buffers only (weight, bias, running_mean, running_var); eps 1e-5 inside the rsqrt."""

    def __init__(self, n: int):
        super().__init__()
        self.register_buffer("weight", torch.ones(n))
        self.register_buffer("bias", torch.zeros(n))
        self.register_buffer("running_mean", torch.zeros(n))
        self.register_buffer("running_var", torch.ones(n))

    def _load_from_state_dict(self, state_dict, prefix, *args):
        state_dict.pop(prefix + "num_batches_tracked", None)
        super()._load_from_state_dict(state_dict, prefix, *args)

    def scale_shift(self):
        scale = self.weight * (self.running_var + 1e-5).rsqrt()
        return scale, self.bias - self.running_mean * scale

    def forward(self, x):
        s, b = self.scale_shift()
        return x * s.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)


def _conv_bn(x, conv: nn.Conv2d, bn: FrozenBatchNorm2d, relu: bool):
    """conv -> frozen BN (-> ReLU) with the BN folded into the convolution: w' = w * scale[:, None, None, None], b' = shift.
    Differentiable w.r.t. ``conv.weight`` (layers 2-4 train in the reference), one cuDNN call instead of conv + 2 passes."""
    s, b = bn.scale_shift()
    w = conv.weight * s.view(-1, 1, 1, 1)
    y = F.conv2d(x, w.to(x.dtype), b.to(x.dtype), conv.stride, conv.padding, conv.dilation, conv.groups)
    return F.relu(y, inplace=True) if relu else y


class FrozenBNResNet(nn.Module):
    """The ResNet trunk (conv1 ... layer4) of backbone.py:93-121 with folded FrozenBatchNorm, channels-last, optional bf16."""

    def __init__(self, name: str = "resnet101", train_backbone: bool = True, dilation: bool = False):
        super().__init__()
        import torchvision

        net = getattr(torchvision.models, name)(weights=None, replace_stride_with_dilation=[False, False, dilation],
                                                norm_layer=FrozenBatchNorm2d)
        # torchvision.models._utils.IntermediateLayerGetter keeps the children up to the returned layer: same keys
        self.body = nn.ModuleDict({k: m for k, m in net.named_children() if k not in ("avgpool", "fc")})
        for pname, p in self.body.named_parameters():
            if not train_backbone or ("layer2" not in pname and "layer3" not in pname and "layer4" not in pname):
                p.requires_grad_(False)
        self.num_channels = 512 if name in ("resnet18", "resnet34") else 2048
        self.compute_dtype: Optional[torch.dtype] = None  # torch.bfloat16 on the GPU (set by VisionEncoder)

    def _block(self, blk, x):
        idn = x
        if blk.downsample is not None:
            idn = _conv_bn(x, blk.downsample[0], blk.downsample[1], False)
        if hasattr(blk, "conv3"):  # Bottleneck
            y = _conv_bn(x, blk.conv1, blk.bn1, True)
            y = _conv_bn(y, blk.conv2, blk.bn2, True)
            y = _conv_bn(y, blk.conv3, blk.bn3, False)
        else:  # BasicBlock
            y = _conv_bn(x, blk.conv1, blk.bn1, True)
            y = _conv_bn(y, blk.conv2, blk.bn2, False)
        return F.relu(y + idn, inplace=True)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        b = self.body
        if self.compute_dtype is not None:
            x = x.to(self.compute_dtype)
        x = x.contiguous(memory_format=torch.channels_last)
        x = _conv_bn(x, b["conv1"], b["bn1"], True)
        x = b["maxpool"](x)
        for ln in ("layer1", "layer2", "layer3", "layer4"):
            for blk in b[ln]:
                x = self._block(blk, x)
        return x  # [n, C, H/32, W/32], channels-last


class _Backbone(nn.Module):
    """``BackboneBase`` (backbone.py:69-102): the trunk under ``.body`` plus the mask down-sampling."""

    def __init__(self, name, train_backbone, dilation):
        super().__init__()
        trunk = FrozenBNResNet(name, train_backbone, dilation)
        self.body = trunk.body
        self._trunk = [trunk]  # not a sub-module twice: the parameters live under ``body``
        self.num_channels = trunk.num_channels

    def forward(self, tensor_list: NestedTensor) -> NestedTensor:
        x = self._trunk[0](tensor_list.tensors)
        m = tensor_list.mask
        mask = F.interpolate(m[None].float(), size=x.shape[-2:]).to(torch.bool)[0]
        return NestedTensor(x, mask, tensor_list.durations)


class VisionEncoder(nn.Sequential):
    """``Joiner(backbone, position_embedding)`` (backbone.py:151-165; built by vision_model/__init__.py).  Children "0" and
    "1" like the reference, so a checkpoint's ``vis_encoder.0.body.*`` keys load unchanged."""

    def __init__(self, cfg=None, name: str = "resnet101", train_backbone: bool = True, dilation: bool = False, hidden: int = 256):
        if cfg is not None:
            name = cfg.MODEL.VISION_BACKBONE.NAME
            dilation = bool(cfg.MODEL.VISION_BACKBONE.DILATION)
            train_backbone = float(cfg.SOLVER.VIS_BACKBONE_LR) > 0  # vision_model/__init__.py:7
            hidden = cfg.MODEL.STCAT.HIDDEN
        super().__init__(_Backbone(name, train_backbone, dilation), PositionEmbeddingSine(hidden // 2))
        self.num_channels = self[0].num_channels

    def set_compute_dtype(self, dtype: Optional[torch.dtype]):
        self[0]._trunk[0].compute_dtype = dtype
        return self

    def forward(self, tensor_list: NestedTensor):
        out = self[0](tensor_list)
        return out, self[1](out)
