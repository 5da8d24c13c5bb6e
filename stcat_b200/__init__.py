"""stcat_b200: the spatio-temporal grounding hot path of jy0205/STCAT on B200 (sm_100a).

Public surface (mirrors reference models/grounding_model/__init__.py and models/pipeline.py):
    build_encoder(cfg), build_decoder(cfg)      -- drop-ins for the reference factories
    STCATHotPath(cfg)                           -- encoder + decoder + prediction heads
    STCATNet(cfg)                               -- the reference's outer seam: model(videos, texts) incl. backbone / text side
    PostProcess                                 -- T x T start/end scoring on device
    evaluate.double_pass                        -- the even/odd evaluation pass as one ragged batch + device interpolation
    optim.make_optimizer / FusedAdamW           -- clip + AdamW + EMA + bf16 weight refresh in one pass
    NestedTensor, get_default_cfg
The arithmetic lives in libstcat_sm100.so (stcat_b200/csrc, C ABI in include/stcat_b200.h); there is
no CPU / PyTorch fallback.
"""
from .config import CfgNode, get_default_cfg  # noqa: F401
from .nested import NestedTensor  # noqa: F401


def build_encoder(cfg):
    from .encoder import build_encoder as _b

    return _b(cfg)


def build_decoder(cfg):
    from .decoder import build_decoder as _b

    return _b(cfg)


def __getattr__(name):
    if name in ("STCATHotPath", "STCATNet", "PostProcess"):
        from . import pipeline

        return getattr(pipeline, name)
    raise AttributeError(name)
