"""Shared helpers for the test-suite (test infrastructure)."""
import os

import torch

from stcat_b200.config import get_default_cfg
from stcat_b200 import synthetic
from stcat_b200.param_spec import synthetic_params

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["b1_T8_res224_L8", "b2_ragged_T5_3", "b3_ragged_T4_1_6", "b1_T12_res320_L16", "b2_ragged_T4_6_mdetr"]


def load_golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, f"{name}.pt"), weights_only=False)


def cfg_for(spec, dropout=0.0):
    cfg = get_default_cfg()
    cfg.merge_from_list(["INPUT.MAX_VIDEO_LEN", spec["max_video_len"], "MODEL.STCAT.DROPOUT", dropout,
                         "MODEL.STCAT.FROM_SCRATCH", bool(spec.get("from_scratch", True))])
    return cfg


def case_inputs(spec):
    return synthetic.make_inputs(spec["durations"], spec["H"], spec["W"], spec["L"], seed=spec["seed"],
                                 ragged=spec["ragged"])


def case_params(cfg, spec):
    return synthetic_params(cfg, seed=spec["seed"])


def rel_err(a, b):
    """max-abs error relative to max |reference| (the measure used in SURVEY.md 7.3-1)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
