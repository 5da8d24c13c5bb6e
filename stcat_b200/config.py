"""Configuration surface for the STCAT hot path.

The reference drives everything from a yacs ``CfgNode`` (reference
``config/defaults.py:6-177``).  yacs is not installed in this image, and the
hot path only ever *reads* attributes (``cfg.MODEL.STCAT.HIDDEN`` ...), so a
small attribute-dict with the yacs methods the reference scripts call
(``clone / merge_from_file / merge_from_list / freeze / dump``) is enough.  Any
object with the same attributes (a real yacs node included) works with
``build_encoder`` / ``build_decoder``.

Only option *names and default values* are mirrored here (they are the
contract); unsupported values raise in the module constructors instead of
silently falling back.
"""
from __future__ import annotations

import ast
import copy
from typing import Any, Dict, Iterable


class CfgNode(dict):
    """Attribute-style nested dict with the subset of the yacs API the reference uses."""

    _FROZEN = "__frozen__"

    def __init__(self, init: Dict[str, Any] | None = None):
        super().__init__()
        object.__setattr__(self, CfgNode._FROZEN, False)
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    # attribute access -------------------------------------------------
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name, value):
        if object.__getattribute__(self, CfgNode._FROZEN):
            raise AttributeError(f"attempted to set {name} on a frozen CfgNode")
        self[name] = value

    # yacs API ---------------------------------------------------------
    def clone(self) -> "CfgNode":
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        out = CfgNode()
        for k, v in self.items():
            dict.__setitem__(out, k, copy.deepcopy(v, memo))
        object.__setattr__(out, CfgNode._FROZEN, object.__getattribute__(self, CfgNode._FROZEN))
        return out

    def freeze(self):
        object.__setattr__(self, CfgNode._FROZEN, True)
        for v in self.values():
            if isinstance(v, CfgNode):
                v.freeze()

    def defrost(self):
        object.__setattr__(self, CfgNode._FROZEN, False)
        for v in self.values():
            if isinstance(v, CfgNode):
                v.defrost()

    def is_frozen(self) -> bool:
        return object.__getattribute__(self, CfgNode._FROZEN)

    def _set_path(self, dotted: str, value):
        node = self
        parts = dotted.split(".")
        for p in parts[:-1]:
            if p not in node:
                raise KeyError(f"Non-existent config key: {dotted}")
            node = node[p]
        if parts[-1] not in node:
            raise KeyError(f"Non-existent config key: {dotted}")
        old = node[parts[-1]]
        if isinstance(value, str) and not isinstance(old, str):
            try:
                value = ast.literal_eval(value)
            except (ValueError, SyntaxError):
                pass
        if isinstance(old, tuple) and isinstance(value, list):
            value = tuple(value)
        if isinstance(old, list) and isinstance(value, tuple):
            value = list(value)
        if isinstance(old, float) and isinstance(value, int):
            value = float(value)
        dict.__setitem__(node, parts[-1], value)

    def merge_from_list(self, opts: Iterable[Any]):
        opts = list(opts)
        if len(opts) % 2:
            raise ValueError("override list must have an even length: KEY VALUE ...")
        for k, v in zip(opts[0::2], opts[1::2]):
            self._set_path(k, v)

    def merge_from_other(self, other: Dict[str, Any], _prefix=""):
        for k, v in other.items():
            if isinstance(v, dict):
                if k not in self:
                    raise KeyError(f"Non-existent config key: {_prefix}{k}")
                self[k].merge_from_other(v, _prefix + k + ".")
            else:
                self._set_path(k, v)

    def merge_from_file(self, path: str):
        import yaml

        with open(path, "r") as f:
            data = yaml.safe_load(f) or {}
        self.merge_from_other(data)

    def dump(self) -> str:
        import yaml

        def plain(n):
            if isinstance(n, CfgNode):
                return {k: plain(v) for k, v in n.items()}
            if isinstance(n, tuple):
                return list(n)
            return n

        return yaml.safe_dump(plain(self), default_flow_style=None)


def get_default_cfg() -> CfgNode:
    """Defaults of the options the hot path (and the synthetic train step around it) reads.

    Mirrors names/values of reference ``config/defaults.py``: INPUT (:20-41), MODEL.STCAT
    (:85-103), SOLVER loss coefficients / flags (:132-176).  The other sections of the
    reference config are for subsystems that are out of scope (datasets, backbone names,
    LR schedule) and are carried only where ``STCATNet``-level code reads them.
    """
    return CfgNode(
        {
            "FROM_SCRATCH": True,
            "INPUT": {
                "MAX_QUERY_LEN": 26,
                "MAX_VIDEO_LEN": 200,
                "TRAIN_SAMPLE_NUM": 64,
                "RESOLUTION": 224,
            },
            "MODEL": {
                "DEVICE": "cuda",
                "WEIGHT": "",
                "EMA": True,
                "EMA_DECAY": 0.9998,
                "QUERY_NUM": 1,
                "VISION_BACKBONE": {"NAME": "resnet101", "POS_ENC": "sine", "DILATION": False, "FREEZE": False},
                "TEXT_MODEL": {"NAME": "roberta-base", "FREEZE": False},
                "USE_LSTM": False,
                "STCAT": {
                    "HIDDEN": 256,
                    "QUERY_DIM": 4,
                    "ENC_LAYERS": 6,
                    "DEC_LAYERS": 6,
                    "FFN_DIM": 2048,
                    "DROPOUT": 0.1,
                    "HEADS": 8,
                    "USE_LEARN_TIME_EMBED": False,
                    "USE_ACTION": True,
                    "FROM_SCRATCH": True,
                    "TEMP_PRED_LAYERS": 6,
                    "CONV_LAYERS": 4,
                    "TEMP_HEAD": "attn",
                    "KERNAL_SIZE": 9,
                    "MAX_MAP_SIZE": 128,
                    "POOLING_COUNTS": [15, 8, 8, 8],
                },
            },
            "SOLVER": {
                "BATCH_SIZE": 1,
                "BASE_LR": 2e-5,
                "VIS_BACKBONE_LR": 1e-5,
                "TEXT_LR": 2e-5,
                "TEMP_LR": 1e-4,
                "WEIGHT_DECAY": 0.0001,
                "MAX_GRAD_NORM": 0.1,
                "BBOX_COEF": 5,
                "GIOU_COEF": 2,
                "TEMP_COEF": 2,
                "ATTN_COEF": 1,
                "ACTIONESS_COEF": 2,
                "USE_ATTN": True,
                "SIGMA": 2.0,
                "USE_AUX_LOSS": True,
                "EOS_COEF": 0.1,
            },
        }
    )


#: BASELINE.json configs -> overrides (experiments/*/e2e_STCAT_R101_*.yaml values that the hot path reads)
NAMED_CONFIGS = {
    "cfg1_T8_res224_L8": dict(T=8, res=224, L=8, overrides=["INPUT.MAX_VIDEO_LEN", 200]),
    "cfg2_hcstvg_T48_res416_L16": dict(T=48, res=416, L=16, overrides=["INPUT.MAX_VIDEO_LEN", 200]),
    "cfg3_vidstg_T64_res448_L16": dict(T=64, res=448, L=16, overrides=["INPUT.MAX_VIDEO_LEN", 300]),
    "cfg4_long_T128_res448_L16": dict(T=128, res=448, L=16, overrides=["INPUT.MAX_VIDEO_LEN", 300]),
}
