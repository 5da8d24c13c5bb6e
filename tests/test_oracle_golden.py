"""Pins oracle/stcat_oracle.py against outputs of the UNMODIFIED reference (tests/golden/*.pt, made by
oracle/make_golden.py from /root/reference).  CPU only."""
import pytest
import torch

from oracle import stcat_oracle as O
from helpers import GOLDEN_CASES, load_golden, cfg_for, case_inputs, case_params, rel_err

TOL = 2e-5  # fp32 oracle vs fp32 reference: same arithmetic, different summation order


@pytest.fixture(scope="module", params=GOLDEN_CASES)
def case(request):
    fx = load_golden(request.param)
    spec = fx["spec"]
    cfg = cfg_for(spec)
    inp = case_inputs(spec)
    P = case_params(cfg, spec)
    return fx, spec, cfg, inp, P


def test_input_recipe_is_stable(case):
    fx, spec, cfg, inp, P = case
    for k, cs in fx["inputs_checksum"].items():
        t = inp[k].double()
        got = torch.tensor([t.sum(), t.abs().sum(), (t * t).sum()])
        assert torch.allclose(got, cs, rtol=1e-12, atol=1e-9), k


def test_forward_matches_reference(case):
    fx, spec, cfg, inp, P = case
    with torch.no_grad():
        out = O.hot_path_forward(P, cfg, inp["vis_features"], inp["vis_mask"], inp["durations"], inp["vis_pos"],
                                 inp["text_mask"], inp["text_memory"])
    c = out["_memory_cache"]
    assert torch.equal(c["mask"], fx["cache"]["mask"])
    for k in ("encoded_memory", "frames_cls", "videos_cls"):
        assert rel_err(c[k], fx["cache"][k]) < TOL, k
    assert rel_err(out["_hs"], fx["hs"]) < TOL
    assert rel_err(out["_reference"], fx["reference"]) < TOL
    assert rel_err(out["_time_hs"], fx["time_hs"]) < TOL
    assert rel_err(out["_weights_all"], fx["weights_all"]) < TOL
    for k in ("pred_boxes", "pred_sted", "pred_actioness", "weights"):
        assert out[k].shape == fx["out"][k].shape
        assert rel_err(out[k], fx["out"][k]) < TOL, k
    for a, g in zip(out["aux_outputs"], fx["aux"]):
        for k in g:
            assert rel_err(a[k], g[k]) < TOL, k


def test_loss_and_gradients_match_reference(case):
    fx, spec, cfg, inp, P = case
    cfg.merge_from_list(["SOLVER.GIOU_COEF", 3, "SOLVER.TEMP_COEF", 10, "SOLVER.EOS_COEF", 0.3])
    from stcat_b200 import synthetic

    tg = synthetic.make_targets(spec["durations"], seed=spec["seed"])
    P = {k: v.clone().requires_grad_(not k.endswith(".te")) for k, v in P.items()}
    vis = inp["vis_features"].clone().requires_grad_(True)
    txt = inp["text_memory"].clone().requires_grad_(True)
    out = O.hot_path_forward(P, cfg, vis, inp["vis_mask"], inp["durations"], inp["vis_pos"], inp["text_mask"], txt)
    total, losses = O.stg_loss(cfg, out, tg["boxes"], tg["actioness"], spec["durations"])
    for k, v in fx["loss"].items():
        assert abs(float(losses[k].detach()) - float(v)) <= 2e-5 * max(1.0, abs(float(v))), k
    assert abs(float(total.detach()) - float(fx["loss_total"])) <= 2e-5 * abs(float(fx["loss_total"]))
    total.backward()
    # Input gradients.  fp32 backward through this network is conditioned at the 1e-3 level: the
    # reference's own fp32 gradients differ from its fp64 gradients by up to 2.3e-3 (max-abs / max|ref|)
    # on these cases, so fp32-vs-fp32 and fp32-vs-fp64 are gated at 5e-3; the semantic pin is the
    # fp64-vs-fp64 test below (1e-6).
    assert rel_err(vis.grad, fx["grad64"]["vis_features"]) < 5e-3
    assert rel_err(txt.grad, fx["grad64"]["text_memory"]) < 5e-3
    assert rel_err(vis.grad, fx["grad"]["vis_features"]) < 5e-3
    assert rel_err(txt.grad, fx["grad"]["text_memory"]) < 5e-3
    for k, g in fx["grad_full"].items():
        assert rel_err(P[k].grad, g) < 5e-3, k
    for k, gn in fx["grad_norm"].items():
        if k.startswith("ground_decoder.decoder.bbox_embed."):
            continue  # alias of bbox_embed.* (same parameter object in the reference)
        if torch.isnan(gn):  # parameters the reference never uses (fusion, ca_qtime_proj)
            assert P[k].grad is None or float(P[k].grad.abs().max()) == 0.0, k
        else:
            got = float(P[k].grad.double().norm())
            assert abs(got - float(gn)) <= 5e-3 * float(gn) + 1e-5, (k, got, float(gn))  # 1e-6: grads that are 0 up to rounding


def test_post_process_matches_reference(case):
    fx, spec, cfg, inp, P = case
    post = fx["post"]
    boxes, steds, _ = O.post_process(fx["out"]["pred_sted"], fx["out"]["pred_boxes"], post["target_sizes"],
                                     post["frames_id"], spec["durations"])
    assert torch.allclose(boxes, post["boxes"], rtol=1e-6, atol=1e-5)
    assert steds == post["steds"]


def test_map2d_matches_reference():
    fx = load_golden("map2d_N16")
    from stcat_b200.config import CfgNode
    from stcat_b200.synthetic import fill_param

    c = CfgNode(fx["cfg"])
    d = 256
    P = {}
    for i in range(c.CONV_LAYERS):
        P[f"encoder.convs.{i}.weight"] = fill_param(f"encoder.convs.{i}.weight", (d, d, c.KERNAL_SIZE, c.KERNAL_SIZE), fx["seed"])
        P[f"encoder.convs.{i}.bias"] = fill_param(f"encoder.convs.{i}.bias", (d,), fx["seed"])
    P["predictor.weight"] = fill_param("predictor.weight", (1, d, 1, 1), fx["seed"])
    P["predictor.bias"] = fill_param("predictor.bias", (1,), fx["seed"])
    P = {"head." + k: v for k, v in P.items()}
    m = O.gen_2d_map(fx["x"].view(-1, 20, d), c.MAX_MAP_SIZE, c.POOLING_COUNTS)
    assert torch.equal(m, fx["map2d"])
    m2 = O.gen_2d_map(fx["x2"].view(-1, 12, d), c.MAX_MAP_SIZE, c.POOLING_COUNTS)
    assert torch.equal(m2, fx["map2d_2"])
    mask2d, _, _ = O.map2d_masks(c.MAX_MAP_SIZE, c.POOLING_COUNTS)
    assert torch.equal(mask2d, fx["mask2d"])
    s_eval = O.map2d_conv_head(P, "head", fx["x"], c, training=False)
    s_train = O.map2d_conv_head(P, "head", fx["x"], c, training=True)
    assert rel_err(s_eval, fx["scores_eval"]) < 1e-4
    assert rel_err(s_train, fx["scores_train"]) < 1e-4


def test_map2d_attn_head_matches_reference():
    """TEMP_HEAD 'attn' (the reference default): row / column attention over the map with the reference's mask quirks"""
    fx = load_golden("map2d_attn_N16")
    from stcat_b200.config import CfgNode
    from stcat_b200.synthetic import fill_param

    c = CfgNode(fx["cfg"])
    P = {"head." + k: fill_param(k, shape, fx["seed"]) for k, shape in fx["shapes"].items()}
    s_eval = O.map2d_attn_head(P, "head", fx["x"], c, training=False)
    s_train = O.map2d_attn_head(P, "head", fx["x"], c, training=True)
    assert rel_err(s_eval, fx["scores_eval"]) < 1e-4
    assert rel_err(s_train, fx["scores_train"]) < 1e-4


def test_fp64_oracle_matches_fp64_reference(case):
    """Semantic pin at 1e-6: oracle in float64 vs the reference run in float64 (forward and backward)."""
    fx, spec, cfg, inp, P = case
    cfg.merge_from_list(["SOLVER.GIOU_COEF", 3, "SOLVER.TEMP_COEF", 10, "SOLVER.EOS_COEF", 0.3])
    from stcat_b200 import synthetic

    dt = torch.float64
    tg = synthetic.make_targets(spec["durations"], seed=spec["seed"])
    P = {k: v.to(dt).requires_grad_(not k.endswith(".te")) for k, v in P.items()}
    vis = inp["vis_features"].to(dt).requires_grad_(True)
    txt = inp["text_memory"].to(dt).requires_grad_(True)
    out = O.hot_path_forward(P, cfg, vis, inp["vis_mask"], inp["durations"], inp["vis_pos"].to(dt), inp["text_mask"],
                             txt, O.Prec(dtype=dt))
    for k in ("pred_boxes", "pred_sted", "pred_actioness", "weights"):
        assert rel_err(out[k], fx["out64"][k]) < 1e-6, k
    assert rel_err(out["_memory_cache"]["encoded_memory"], fx["encoded_memory64"]) < 1e-6
    total, _ = O.stg_loss(cfg, out, tg["boxes"].to(dt), tg["actioness"].to(dt), spec["durations"])
    assert abs(float(total.detach()) - float(fx["loss_total64"])) < 1e-6 * abs(float(fx["loss_total64"]))
    total.backward()
    assert rel_err(vis.grad, fx["grad64"]["vis_features"]) < 1e-6
    assert rel_err(txt.grad, fx["grad64"]["text_memory"]) < 1e-6
    for k, gn in fx["grad_norm64"].items():
        if k.startswith("ground_decoder.decoder.bbox_embed.") or torch.isnan(gn):
            continue
        got = float(P[k].grad.norm())
        assert abs(got - float(gn)) <= 1e-6 * float(gn) + 1e-10, (k, got, float(gn))
