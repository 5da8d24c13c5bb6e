#!/usr/bin/env python
"""Diagnosis: is the bf16-mode forward loss bit-reproducible run to run, single- vs multi-stream decoder?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import load_golden, cfg_for, case_inputs, case_params
from stcat_b200 import ops, synthetic, decoder as dec
from stcat_b200.loss import STGLossPlan
from stcat_b200.nested import NestedTensor
from stcat_b200.pipeline import STCATHotPath

fx = load_golden("b2_ragged_T5_3"); spec = fx["spec"]; cfg = cfg_for(spec)
inp = case_inputs(spec); tg = synthetic.make_targets(spec["durations"], seed=spec["seed"])
for precision in ("fp32", "bf16"):
    ops.set_precision(precision)
    for streams in (False, True, False, True):
        vals = []
        for rep in range(4):
            ops.clear_weight_cache()
            m = STCATHotPath(cfg).load_flat_params(case_params(cfg, spec)).cuda().eval()
            dec.set_multi_stream(streams)
            with torch.no_grad():
                vis = inp["vis_features"].cuda(); txt = inp["text_memory"].cuda()
                out = m(NestedTensor(vis, inp["vis_mask"].cuda(), inp["durations"]), inp["vis_pos"].cuda(), (inp["text_mask"].cuda(), txt, None))
                total, _ = STGLossPlan(cfg, tg["boxes"], tg["actioness"], spec["durations"], "cuda")(out)
            torch.cuda.synchronize()
            vals.append((float(total), float(out["_memory_cache"]["encoded_memory"].double().sum()), float(out["_hs"].double().sum()), float(out["_time_hs"].double().sum())))
        print(precision, "streams" if streams else "single ", " | ".join(f"{v[0]:.6f} enc={v[1]:.5f} hs={v[2]:.5f} ths={v[3]:.5f}" for v in vals), flush=True)
dec.set_multi_stream(True)
