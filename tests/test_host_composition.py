"""CPU (-m "not gpu"): the host-side composition of stcat_b200 (layouts, index maps, hand-written
backward chains of the fused blocks) driven through the torch-CPU emulation of the C ABI
(tests/emu_backend.py) and compared with the oracle and the golden fixtures of the reference.
The CUDA kernels themselves are compared with the same oracle in the -m gpu tests."""
import pytest
import torch

from oracle import stcat_oracle as O
from helpers import GOLDEN_CASES, load_golden, cfg_for, case_inputs, case_params, rel_err
from stcat_b200 import ops, synthetic
from stcat_b200.nested import NestedTensor
from stcat_b200.pipeline import STCATHotPath
from stcat_b200.param_spec import hot_path_spec
from emu_backend import EmuBackend


@pytest.fixture(autouse=True)
def emu():
    ops.set_backend(EmuBackend())
    ops.set_precision("fp32")
    yield
    ops.set_backend(None)
    ops.set_precision("fp32")
    ops.clear_weight_cache()


def build(cfg, P):
    m = STCATHotPath(cfg)
    m.load_flat_params(P)
    return m


def run_model(m, inp, grad=False):
    vis = inp["vis_features"].clone().requires_grad_(grad)
    txt = inp["text_memory"].clone().requires_grad_(grad)
    videos = NestedTensor(vis, inp["vis_mask"].clone(), inp["durations"])
    out = m(videos, inp["vis_pos"], (inp["text_mask"], txt, None))
    return out, vis, txt


def test_state_dict_matches_contract():
    cfg = cfg_for({"max_video_len": 20})
    m = STCATHotPath(cfg)
    sd = m.state_dict()
    spec = hot_path_spec(cfg)
    assert set(sd.keys()) == set(spec.keys())
    for k, shape in spec.items():
        assert tuple(sd[k].shape) == tuple(shape), k
    # the alias of pipeline.py:50
    assert m.ground_decoder.decoder.bbox_embed is m.bbox_embed


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_forward_matches_golden(name):
    fx = load_golden(name)
    spec = fx["spec"]
    cfg = cfg_for(spec)
    inp = case_inputs(spec)
    m = build(cfg, case_params(cfg, spec)).eval()
    with torch.no_grad():
        out, _, _ = run_model(m, inp)
    TOL = 5e-5
    c = out["_memory_cache"]
    assert torch.equal(c["mask"], fx["cache"]["mask"])
    for k in ("encoded_memory", "frames_cls", "videos_cls"):
        assert c[k].shape == fx["cache"][k].shape, k
        assert rel_err(c[k], fx["cache"][k]) < TOL, k
    assert rel_err(out["_hs"], fx["hs"]) < TOL
    assert rel_err(out["_reference"], fx["reference"]) < TOL
    assert rel_err(out["_time_hs"], fx["time_hs"]) < TOL
    assert rel_err(out["_weights_all"], fx["weights_all"]) < TOL
    for k in ("pred_boxes", "pred_sted", "pred_actioness", "weights"):
        assert out[k].shape == fx["out"][k].shape, k
        assert rel_err(out[k], fx["out"][k]) < TOL, k
    for a, g in zip(out["aux_outputs"], fx["aux"]):
        for k in g:
            assert rel_err(a[k], g[k]) < TOL, k


@pytest.mark.parametrize("name", ["b1_T8_res224_L8", "b3_ragged_T4_1_6", "b2_ragged_T4_6_mdetr"])
def test_gradients_match_golden(name):
    fx = load_golden(name)
    spec = fx["spec"]
    cfg = cfg_for(spec)
    cfg.merge_from_list(["SOLVER.GIOU_COEF", 3, "SOLVER.TEMP_COEF", 10, "SOLVER.EOS_COEF", 0.3])
    inp = case_inputs(spec)
    m = build(cfg, case_params(cfg, spec)).eval()  # eval like the fixture: the 0.3 head dropout is identity
    out, vis, txt = run_model(m, inp, grad=True)
    tg = synthetic.make_targets(spec["durations"], seed=spec["seed"])
    total, losses = O.stg_loss(cfg, out, tg["boxes"], tg["actioness"], spec["durations"])
    assert abs(float(total.detach()) - float(fx["loss_total"])) <= 5e-5 * abs(float(fx["loss_total"]))
    total.backward()
    assert rel_err(vis.grad, fx["grad64"]["vis_features"]) < 5e-3
    assert rel_err(txt.grad, fx["grad64"]["text_memory"]) < 5e-3
    named = dict(m.named_parameters())
    for k, g in fx["grad_full"].items():
        assert rel_err(named[k].grad, g) < 5e-3, k
    for k, gn in fx["grad_norm"].items():
        if k.startswith("ground_decoder.decoder.bbox_embed."):
            continue
        p = named[k]
        if torch.isnan(gn):
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
        else:
            assert p.grad is not None, k
            got = float(p.grad.double().norm())
            assert abs(got - float(gn)) <= 5e-3 * float(gn) + 2e-6, (k, got, float(gn))  # abs floor: fp32 noise on ~0 grads


def test_bf16_mode_tracks_rounded_oracle():
    """bf16 operand mode vs the oracle with bf16-rounded matmul operands (SURVEY.md 7.3-1 policy)."""
    fx = load_golden("b1_T8_res224_L8")
    spec = fx["spec"]
    cfg = cfg_for(spec)
    inp = case_inputs(spec)
    P = case_params(cfg, spec)
    m = build(cfg, P).eval()
    ops.set_precision("bf16")
    with torch.no_grad():
        out, _, _ = run_model(m, inp)
        ref = O.hot_path_forward(P, cfg, inp["vis_features"], inp["vis_mask"], inp["durations"], inp["vis_pos"],
                                 inp["text_mask"], inp["text_memory"], prec=O.Prec(round_operands="bf16"))
    for k in ("pred_boxes", "pred_sted", "pred_actioness", "weights"):
        assert rel_err(out[k], ref[k]) < 2e-2, k  # loose: rounding points differ slightly


def test_mdetr_branch_bf16_mode_forward_and_backward():
    """MODEL.STCAT.FROM_SCRATCH False (cross_attn_image branch of the box decoder) in bf16 operand mode: tracks the oracle with
    bf16-rounded operands, and the backward chain runs through the operand-dtype plumbing of that branch."""
    fx = load_golden("b2_ragged_T4_6_mdetr")
    spec = fx["spec"]
    cfg = cfg_for(spec)
    assert not cfg.MODEL.STCAT.FROM_SCRATCH
    inp = case_inputs(spec)
    P = case_params(cfg, spec)
    m = build(cfg, P).eval()
    assert "ground_decoder.decoder.layers.0.cross_attn_image.in_proj_weight" in m.state_dict()
    assert not any(".decoder.layers.0.cross_attn." in k for k in m.state_dict())
    ops.set_precision("bf16")
    out, vis, txt = run_model(m, inp, grad=True)
    with torch.no_grad():
        ref = O.hot_path_forward(P, cfg, inp["vis_features"], inp["vis_mask"], inp["durations"], inp["vis_pos"],
                                 inp["text_mask"], inp["text_memory"], prec=O.Prec(round_operands="bf16"))
    for k in ("pred_boxes", "pred_sted", "pred_actioness", "weights"):
        assert rel_err(out[k], ref[k]) < 2e-2, k
    (out["pred_boxes"].sum() + out["pred_sted"].sum()).backward()
    named = dict(m.named_parameters())
    for k in ("ground_decoder.decoder.layers.0.cross_attn_image.in_proj_weight", "ground_decoder.decoder.layers.3.ca_qtime_proj.weight",
              "ground_decoder.decoder.layers.0.ca_kpos_proj.weight"):
        assert named[k].grad is not None and torch.isfinite(named[k].grad).all() and float(named[k].grad.abs().max()) > 0, k
    assert torch.isfinite(vis.grad).all() and torch.isfinite(txt.grad).all()


@pytest.mark.parametrize("name", ["b1_T8_res224_L8", "b2_ragged_T5_3"])
def test_device_loss_matches_oracle_loss(name):
    from stcat_b200.loss import STGLossPlan, loss_weight_dict

    fx = load_golden(name)
    spec = fx["spec"]
    cfg = cfg_for(spec)
    cfg.merge_from_list(["SOLVER.GIOU_COEF", 3, "SOLVER.TEMP_COEF", 10, "SOLVER.EOS_COEF", 0.3])
    inp = case_inputs(spec)
    m = build(cfg, case_params(cfg, spec)).eval()
    with torch.no_grad():
        out, _, _ = run_model(m, inp)
    tg = synthetic.make_targets(spec["durations"], seed=spec["seed"])
    plan = STGLossPlan(cfg, tg["boxes"], tg["actioness"], spec["durations"], "cpu")
    total, named = plan(out)
    ref_total, ref_named = O.stg_loss(cfg, out, tg["boxes"], tg["actioness"], spec["durations"])
    assert loss_weight_dict(cfg) == O.loss_weight_dict(cfg)
    assert set(named) == set(ref_named)
    for k in ref_named:
        assert abs(float(named[k]) - float(ref_named[k])) <= 1e-5 * max(1.0, abs(float(ref_named[k]))), k
    assert abs(float(total) - float(ref_total)) <= 1e-5 * abs(float(ref_total))
    assert abs(float(total) - float(fx["loss_total"])) <= 5e-5 * abs(float(fx["loss_total"]))


def test_grad_fusion_matches_autograd_accumulation():
    """ops.set_grad_fusion(True): wgrad kernels accumulate straight into pre-allocated .grad views of one flat
    buffer; the result must equal the ordinary autograd accumulation."""
    fx = load_golden("b2_ragged_T5_3")
    spec = fx["spec"]
    cfg = cfg_for(spec)
    inp = case_inputs(spec)
    tg = synthetic.make_targets(spec["durations"], seed=spec["seed"])
    grads = []
    for fused in (False, True):
        m = build(cfg, case_params(cfg, spec)).eval()
        params = list(dict.fromkeys(m.parameters()))
        if fused:
            flat = torch.zeros(sum(p.numel() for p in params))
            o = 0
            for p in params:
                p.grad = flat[o:o + p.numel()].view_as(p)
                o += p.numel()
        ops.set_grad_fusion(fused)
        try:
            out, vis, txt = run_model(m, inp, grad=True)
            total, _ = O.stg_loss(cfg, out, tg["boxes"], tg["actioness"], spec["durations"])
            total.backward()
        finally:
            ops.set_grad_fusion(False)
        grads.append({k: (None if p.grad is None else p.grad.clone()) for k, p in m.named_parameters()})
        grads[-1]["__vis"] = vis.grad.clone()
    for k, g in grads[0].items():
        g2 = grads[1][k]
        if g is None:
            assert g2 is None or float(g2.abs().max()) == 0.0, k
        else:
            assert rel_err(g2, g) < 1e-5 or float((g2 - g).abs().max()) < 1e-7, k


def _map2d_head(fx, extra):
    from stcat_b200.config import get_default_cfg
    from stcat_b200.map2d import TempPredictionHead
    from stcat_b200.synthetic import fill_param

    cfg = get_default_cfg()
    c = fx["cfg"]
    cfg.merge_from_list(["MODEL.STCAT.MAX_MAP_SIZE", c["MAX_MAP_SIZE"], "MODEL.STCAT.POOLING_COUNTS", c["POOLING_COUNTS"],
                         "MODEL.STCAT.DROPOUT", 0.0] + extra)
    head = TempPredictionHead(cfg)
    sd = {k: fill_param(k, tuple(v.shape), fx["seed"]) for k, v in head.state_dict().items()}
    head.load_state_dict(sd)
    return head


@pytest.mark.parametrize("variant", ["conv", "attn"])
def test_map2d_head_matches_reference_fixture(variant):
    """stcat_b200.map2d.TempPredictionHead through the emulated C ABI vs the scores of the unmodified
    reference head (tests/golden/map2d*_N16.pt), train-mode logits and eval-mode sigmoid * mask; same state-dict names."""
    if variant == "conv":
        fx = load_golden("map2d_N16")
        c = fx["cfg"]
        head = _map2d_head(fx, ["MODEL.STCAT.TEMP_HEAD", "conv", "MODEL.STCAT.KERNAL_SIZE", c["KERNAL_SIZE"],
                                "MODEL.STCAT.CONV_LAYERS", c["CONV_LAYERS"]])
        assert set(head.state_dict()) == {f"encoder.convs.{i}.{p}" for i in range(c["CONV_LAYERS"]) for p in ("weight", "bias")} | \
            {"predictor.weight", "predictor.bias"}
        with torch.no_grad():
            m = head.map_maker(fx["x"].view(-1, 20, 256))
            assert torch.equal(m, fx["map2d"])
            assert torch.equal(head.map_maker(fx["x2"].view(-1, 12, 256)), fx["map2d_2"])  # T < N: adaptive max-pool upsampling
    else:
        fx = load_golden("map2d_attn_N16")
        c = fx["cfg"]
        head = _map2d_head(fx, ["MODEL.STCAT.TEMP_HEAD", "attn", "MODEL.STCAT.TEMP_PRED_LAYERS", c["TEMP_PRED_LAYERS"]])
        assert {k: tuple(v.shape) for k, v in head.state_dict().items()} == fx["shapes"]
    with torch.no_grad():
        head.train()
        assert rel_err(head(fx["x"]), fx["scores_train"]) < 1e-4
        head.eval()
        assert rel_err(head(fx["x"]), fx["scores_eval"]) < 1e-4


@pytest.mark.parametrize("variant", ["conv", "attn"])
def test_map2d_head_gradients_match_oracle(variant):
    """backward of the head (parameters and clip features) vs autograd through the oracle restatement"""
    from stcat_b200.config import CfgNode

    fx = load_golden("map2d_N16" if variant == "conv" else "map2d_attn_N16")
    c = fx["cfg"]
    extra = (["MODEL.STCAT.TEMP_HEAD", "conv", "MODEL.STCAT.KERNAL_SIZE", c["KERNAL_SIZE"], "MODEL.STCAT.CONV_LAYERS", c["CONV_LAYERS"]]
             if variant == "conv" else ["MODEL.STCAT.TEMP_HEAD", "attn", "MODEL.STCAT.TEMP_PRED_LAYERS", c["TEMP_PRED_LAYERS"]])
    head = _map2d_head(fx, extra).train()
    x = fx["x"].clone().requires_grad_(True)
    g = torch.randn(fx["scores_train"].shape, generator=torch.Generator().manual_seed(3))
    (head(x) * g).sum().backward()
    P = {"head." + k: v.detach().clone().requires_grad_(True) for k, v in head.state_dict().items()}
    xo = fx["x"].clone().requires_grad_(True)
    node = CfgNode(dict(c, HEADS=8, HIDDEN=256, FFN_DIM=2048))
    fn = O.map2d_conv_head if variant == "conv" else O.map2d_attn_head
    (fn(P, "head", xo, node, training=True) * g).sum().backward()
    assert rel_err(x.grad, xo.grad) < 1e-4
    for k, prm in head.named_parameters():
        assert rel_err(prm.grad, P["head." + k].grad) < 1e-4, k


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_post_process_matches_reference_fixture(name):
    """stcat_b200.pipeline.PostProcess (mirror of models/post_processor.py:17-55) against the reference's own PostProcess output
    stored in the fixtures: scaled / clamped xyxy boxes and the (start, end) frame ids of the best T x T cell."""
    from stcat_b200.pipeline import PostProcess

    fx = load_golden(name)
    post = fx["post"]
    outputs = {"pred_boxes": fx["out"]["pred_boxes"], "pred_sted": fx["out"]["pred_sted"]}
    boxes, steds = PostProcess()(outputs, post["target_sizes"], post["frames_id"], fx["spec"]["durations"])
    assert torch.allclose(boxes, post["boxes"], rtol=1e-6, atol=1e-5)
    assert steds == post["steds"]


@pytest.mark.parametrize("name", ["b1_T8_res224_L8", "b2_ragged_T5_3"])
@pytest.mark.parametrize("fuse_grads", [False, True])
def test_fused_glue_matches_the_cat_slice_composition(name, fuse_grads):
    """ops.token_assembly / mem_operands / template (csrc/assembly.cu, bf16 mode) against the torch.cat / slice / Linear
    composition they replace: every output, every parameter gradient and both input gradients."""
    from stcat_b200 import encoder as E

    fx = load_golden(name)
    spec = fx["spec"]
    cfg = cfg_for(spec)
    inp = case_inputs(spec)
    tg = synthetic.make_targets(spec["durations"], seed=spec["seed"])
    ops.set_precision("bf16")
    res = []
    for fused_glue, sink in ((False, False), (True, False), (True, True)):
        E.set_fused_glue(fused_glue)
        ops.set_grad_sink(sink)
        ops.clear_weight_cache()
        m = build(cfg, case_params(cfg, spec)).eval()
        params = list(dict.fromkeys(m.parameters()))
        if fuse_grads:
            flat = torch.zeros(sum(p.numel() for p in params))
            o = 0
            for p in params:
                p.grad = flat[o:o + p.numel()].view_as(p)
                o += p.numel()
        ops.set_grad_fusion(fuse_grads)
        try:
            be = ops.get_backend()
            n0 = be.launches
            out, vis, txt = run_model(m, inp, grad=True)
            total, _ = O.stg_loss(cfg, out, tg["boxes"], tg["actioness"], spec["durations"])
            total.backward()
            launches = be.launches - n0
        finally:
            ops.set_grad_fusion(False)
            E.set_fused_glue(True)
            ops.set_grad_sink(True)
        g = {k: (None if p.grad is None else p.grad.clone()) for k, p in m.named_parameters()}
        g["__vis"], g["__txt"] = vis.grad.clone(), txt.grad.clone()
        res.append((out, g, launches))
    (o0, g0, _) = res[0]
    # without the gradient sink the fused glue only reorders fp32 sums; with it (ops._GradSink) the 24 bf16 data gradients of
    # the encoder memory are summed with one rounding per contribution inside the GEMM epilogue instead of a bf16 add per
    # contribution: same terms, different bf16 rounding noise on everything upstream of the decoder (the encoder's gradients)
    for (o1, g1, _), tol in ((res[1], 2e-4), (res[2], 2e-2)):
        for k in ("pred_boxes", "pred_sted", "pred_actioness", "weights"):
            assert rel_err(o1[k], o0[k]) < 1e-5, k
        for k, g in g0.items():
            g2 = g1[k]
            if g is None or float(g.abs().max()) == 0.0:
                assert g2 is None or float(g2.abs().max()) == 0.0, k
            else:
                assert g2 is not None, k
                assert rel_err(g2, g) < tol or float((g2 - g).abs().max()) < 1e-6, (k, tol, rel_err(g2, g))


def test_gradient_sink_with_a_partial_backward_pass():
    """ops._GradSink (bf16 mode): a backward pass that reaches only SOME consumers of the encoder memory (here: the box
    decoder's outputs only; the time decoder's memory-side projections never run their backward) must still deliver the
    gradient of the consumers that did run -- the sink does not count consumers -- and a second, full backward pass through a
    fresh forward must not see leftovers of the first."""
    from stcat_b200 import encoder as E

    fx = load_golden("b2_ragged_T5_3")
    spec = fx["spec"]
    cfg = cfg_for(spec)
    inp = case_inputs(spec)
    ops.set_precision("bf16")
    res = {}
    for sink in (False, True):
        ops.set_grad_sink(sink)
        ops.clear_weight_cache()
        try:
            m = build(cfg, case_params(cfg, spec)).eval()
            out, vis, txt = run_model(m, inp, grad=True)
            out["pred_boxes"].square().sum().backward()  # partial: nothing of the time decoder
            g_partial = vis.grad.clone()
            assert float(g_partial.abs().max()) > 0
            vis.grad = None
            out2, vis2, _ = run_model(m, inp, grad=True)
            (out2["pred_boxes"].square().sum() + out2["pred_sted"].square().sum()).backward()
            res[sink] = (g_partial, vis2.grad.clone())
        finally:
            ops.set_grad_sink(True)
    for a, b in zip(res[False], res[True]):
        assert rel_err(b, a) < 2e-2, rel_err(b, a)  # same terms; bf16 rounding of the accumulated memory gradient differs
    assert rel_err(res[True][0], res[True][1]) > 1e-3  # the two passes really are different gradients


@pytest.mark.parametrize("name", ["b1_T8_res224_L8", "b1_T12_res320_L16"])
def test_early_in_projection_is_bit_identical(name):
    """ops.qkv_early (bf16 mode, one un-padded video): projecting the next spatial layer's q/k/v before the temporal layer and
    re-projecting only the patched frame-CLS rows gives exactly the tensors of the one-shot projection -- outputs and every
    gradient are bit-identical on the emulation backend (same products, same accumulation per row)."""
    from stcat_b200 import encoder as E

    fx = load_golden(name)
    spec = fx["spec"]
    cfg = cfg_for(spec)
    inp = case_inputs(spec)
    assert len(spec["durations"]) == 1
    ops.set_precision("bf16")
    res = []
    old = (E._EARLY_QKV, E._CLS_KERNELS)
    try:
        for early, cls_k in ((False, False), (False, True), (True, True)):
            E._EARLY_QKV, E._CLS_KERNELS = early, cls_k
            ops.clear_weight_cache()
            m = build(cfg, case_params(cfg, spec)).eval()
            be = ops.get_backend()
            n0 = be.launches
            out, vis, txt = run_model(m, inp, grad=True)
            (out["pred_boxes"].square().sum() + out["pred_sted"].square().sum()).backward()
            g = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
            res.append((out, g, vis.grad.clone(), be.launches - n0))
    finally:
        E._EARLY_QKV, E._CLS_KERNELS = old
    for o1, g1, v1, _ in res[1:]:
        for k in ("pred_boxes", "pred_sted", "pred_actioness", "weights"):
            assert torch.equal(o1[k], res[0][0][k]), k
        assert torch.equal(v1, res[0][2])
        assert set(g1) == set(res[0][1])
        for k, g in g1.items():
            assert torch.equal(g, res[0][1][k]), k
