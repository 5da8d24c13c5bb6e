"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference (jy0205/STCAT) from /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  Used by
``oracle/make_golden.py`` to generate the committed fixtures under ``tests/golden/`` and by the
``-m "not gpu"`` tests that pin ``oracle/stcat_oracle.py`` against the live reference when it is
present.  Nothing in ``stcat_b200/`` may import this module.

The reference imports four third-party packages that are absent from this image:
``yacs`` (config/defaults.py:1), ``pytorch_pretrained_bert`` (models/language_model/bert.py:8),
``torchtext`` (models/language_model/lstm.py) and ``ffmpeg`` (datasets/vidstg.py:13).  None of them
is touched by the hot path, so in-memory stubs are sufficient.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("STCAT_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "grounding_model", "modal_encoder.py"))


def _install_stubs():
    from stcat_b200.config import CfgNode

    def mod(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    if "yacs" not in sys.modules:
        mod("yacs")
        mod("yacs.config", CfgNode=CfgNode)
    if "pytorch_pretrained_bert" not in sys.modules:
        mod("pytorch_pretrained_bert")
        mod("pytorch_pretrained_bert.modeling", BertModel=object)
        mod("pytorch_pretrained_bert.tokenization", BertTokenizer=object)
    if "torchtext" not in sys.modules:
        mod("torchtext", vocab=mod("torchtext.vocab"))
    if "ffmpeg" not in sys.modules:
        mod("ffmpeg")


def import_reference():
    """Returns a namespace with the reference's hot-path entry points."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from config import cfg as ref_cfg  # reference config/defaults.py
    from models.grounding_model import build_encoder, build_decoder
    from models.net_utils import MLP, inverse_sigmoid, gen_sineembed_for_position
    from models.post_processor import PostProcess
    from models.criterion import VideoSTGLoss
    from utils.misc import NestedTensor
    from models.vision_model.position_encoding import PositionEmbeddingSine

    ns = types.SimpleNamespace(
        cfg=ref_cfg,
        build_encoder=build_encoder,
        build_decoder=build_decoder,
        MLP=MLP,
        inverse_sigmoid=inverse_sigmoid,
        gen_sineembed_for_position=gen_sineembed_for_position,
        PostProcess=PostProcess,
        VideoSTGLoss=VideoSTGLoss,
        NestedTensor=NestedTensor,
        PositionEmbeddingSine=PositionEmbeddingSine,
    )
    return ns


def import_reference_map2d():
    """map2d_head.py is orphaned in the reference: it reads cfg.MODEL.TEMPFORMER and
    hard-codes ``.to("cuda")`` (map2d_head.py:35).  We alias the cfg node and load the module source
    with that one device string patched in memory (the file on disk is untouched)."""
    ref = import_reference()
    path = os.path.join(REFERENCE_ROOT, "models", "map2d_head.py")
    with open(path, "r") as f:
        src = f.read()
    src = src.replace('.to("cuda")', '.to("cpu")')
    m = types.ModuleType("ref_map2d_head")
    exec(compile(src, path, "exec"), m.__dict__)
    return ref, m
