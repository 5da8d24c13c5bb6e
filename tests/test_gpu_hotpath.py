"""-m gpu: the full hot path (encoder -> decoder -> heads -> loss -> backward) through the C ABI on a
B200, against (a) the golden fixtures produced by the UNMODIFIED reference and (b) the CPU oracle on the
same seeded inputs.

Tolerances (north_star: 1e-3 relative):
  * fp32 kernels vs fp32 reference / oracle: gated at 1e-3, measured ~1e-5 (summation order only);
  * gradients: 1e-2 -- the reference's own fp32 gradients differ from its fp64 gradients by up to
    2.3e-3 on these cases (tests/test_oracle_golden.py) and a different fp32 summation order moves them by
    as much again (measured 5.3e-3 worst case), so fp32-vs-fp64 gradients cannot be gated at 1e-3;
    the forward outputs and the loss are;
  * bf16 tensor-core mode vs the oracle with bf16-rounded matmul operands: 2e-2 on the outputs
    (rounding points of intermediate stores differ; DESIGN.md "precision policy"), and reported
    against the fp32 oracle.
"""
import pytest
import torch

from oracle import stcat_oracle as O
from helpers import GOLDEN_CASES, load_golden, cfg_for, case_inputs, case_params, rel_err
from stcat_b200 import ops, synthetic
from stcat_b200.nested import NestedTensor

pytestmark = pytest.mark.gpu

TOL = 1e-3


@pytest.fixture(autouse=True)
def cuda_backend():
    ops.set_backend(None)  # the real CudaBackend (created lazily; raises if the .so is missing)
    ops.set_precision("fp32")
    ops.clear_weight_cache()
    yield
    ops.set_precision("fp32")
    ops.clear_weight_cache()


def build(cfg, P):
    from stcat_b200.pipeline import STCATHotPath

    return STCATHotPath(cfg).load_flat_params(P).cuda()


def run_model(m, inp, grad=False):
    vis = inp["vis_features"].cuda().requires_grad_(grad)
    txt = inp["text_memory"].cuda().requires_grad_(grad)
    videos = NestedTensor(vis, inp["vis_mask"].cuda(), inp["durations"])
    out = m(videos, inp["vis_pos"].cuda(), (inp["text_mask"].cuda(), txt, None))
    return out, vis, txt


def test_backend_is_native():
    be = ops.get_backend()
    assert be.name == "cuda"
    assert be.lib.stcat_device_arch() >= 100


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_forward_matches_reference_golden(name):
    fx = load_golden(name)
    spec = fx["spec"]
    cfg = cfg_for(spec)
    inp = case_inputs(spec)
    m = build(cfg, case_params(cfg, spec)).eval()
    before = ops.get_backend().launches
    with torch.no_grad():
        out, _, _ = run_model(m, inp)
    assert ops.get_backend().launches > before
    c = out["_memory_cache"]
    assert torch.equal(c["mask"].cpu(), fx["cache"]["mask"])
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


@pytest.mark.parametrize("name", ["b1_T8_res224_L8", "b3_ragged_T4_1_6", "b1_T12_res320_L16"])
def test_loss_and_gradients_match_reference_golden(name):
    from stcat_b200.loss import STGLossPlan

    fx = load_golden(name)
    spec = fx["spec"]
    cfg = cfg_for(spec)
    cfg.merge_from_list(["SOLVER.GIOU_COEF", 3, "SOLVER.TEMP_COEF", 10, "SOLVER.EOS_COEF", 0.3])
    inp = case_inputs(spec)
    m = build(cfg, case_params(cfg, spec)).eval()  # eval like the fixture (0.3 head dropout = identity)
    out, vis, txt = run_model(m, inp, grad=True)
    tg = synthetic.make_targets(spec["durations"], seed=spec["seed"])
    plan = STGLossPlan(cfg, tg["boxes"], tg["actioness"], spec["durations"], "cuda")
    total, named = plan(out)
    for k, v in fx["loss"].items():
        assert abs(float(named[k]) - float(v)) <= 1e-4 * max(1.0, abs(float(v))), k
    assert abs(float(total) - float(fx["loss_total"])) <= 1e-4 * abs(float(fx["loss_total"]))
    total.backward()
    GT = 1e-2  # measured on B200: <= 5.3e-3 (fp32 summation-order noise amplified by the backward's conditioning)
    assert rel_err(vis.grad, fx["grad64"]["vis_features"]) < GT
    assert rel_err(txt.grad, fx["grad64"]["text_memory"]) < GT
    named_p = dict(m.named_parameters())
    for k, g in fx["grad_full"].items():
        assert rel_err(named_p[k].grad, g) < GT, k
    for k, gn in fx["grad_norm"].items():
        if k.startswith("ground_decoder.decoder.bbox_embed."):
            continue
        p = named_p[k]
        if torch.isnan(gn):
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
        else:
            got = float(p.grad.double().norm())
            assert abs(got - float(gn)) <= 1e-2 * float(gn) + 1e-5, (k, got, float(gn))


def test_forward_matches_oracle_midsize():
    """T=16, res=320 (H=W=10), L=12: larger than any fixture, checked against the CPU oracle directly."""
    spec = {"durations": [16], "H": 10, "W": 10, "L": 12, "seed": 3, "ragged": False, "max_video_len": 32}
    cfg = cfg_for(spec)
    inp = case_inputs(spec)
    P = case_params(cfg, spec)
    m = build(cfg, P).eval()
    with torch.no_grad():
        out, _, _ = run_model(m, inp)
        ref = O.hot_path_forward(P, cfg, inp["vis_features"], inp["vis_mask"], inp["durations"], inp["vis_pos"],
                                 inp["text_mask"], inp["text_memory"])
    for k in ("pred_boxes", "pred_sted", "pred_actioness", "weights"):
        assert rel_err(out[k], ref[k]) < TOL, k
    assert rel_err(out["_memory_cache"]["encoded_memory"], ref["_memory_cache"]["encoded_memory"]) < TOL


@pytest.mark.parametrize("name", ["b1_T8_res224_L8", "b2_ragged_T5_3"])
def test_bf16_mode_vs_rounded_oracle(name):
    fx = load_golden(name)
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
    worst = {}
    for k in ("pred_boxes", "pred_sted", "pred_actioness", "weights"):
        worst[k] = (rel_err(out[k], ref[k]), rel_err(out[k], fx["out"][k]))
        assert worst[k][0] < 2e-2, (k, worst[k])
    print("bf16 mode: (vs bf16-rounded oracle, vs fp32 reference)", worst)


def test_post_process_on_device():
    from stcat_b200.pipeline import PostProcess

    fx = load_golden("b2_ragged_T5_3")
    spec = fx["spec"]
    post = fx["post"]
    outputs = {"pred_boxes": fx["out"]["pred_boxes"].cuda(), "pred_sted": fx["out"]["pred_sted"].cuda()}
    boxes, steds = PostProcess()(outputs, post["target_sizes"].cuda(), post["frames_id"], spec["durations"])
    assert torch.allclose(boxes.cpu(), post["boxes"], rtol=1e-6, atol=1e-5)
    assert steds == post["steds"]


def test_cpu_tensors_are_rejected():
    """no CPU fallback: the product path must fail loudly off-device"""
    from stcat_b200.cabi import StcatError

    spec = load_golden("b1_T8_res224_L8")["spec"]
    cfg = cfg_for(spec)
    from stcat_b200.pipeline import STCATHotPath

    m = STCATHotPath(cfg).eval()
    inp = case_inputs(spec)
    videos = NestedTensor(inp["vis_features"], inp["vis_mask"], inp["durations"])
    with pytest.raises(StcatError):
        with torch.no_grad():
            m(videos, inp["vis_pos"], (inp["text_mask"], inp["text_memory"], None))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_grad_fusion_and_single_stream_agree(precision):
    """flat-buffer gradient accumulation (bench.py's mode) and the single-stream decoder give the same gradients
    as the default path (multi-stream decoder, autograd accumulation)"""
    from stcat_b200 import decoder as dec
    from stcat_b200.loss import STGLossPlan

    fx = load_golden("b2_ragged_T5_3")
    spec = fx["spec"]
    cfg = cfg_for(spec)
    inp = case_inputs(spec)
    tg = synthetic.make_targets(spec["durations"], seed=spec["seed"])
    ops.set_precision(precision)
    results = []
    configs = ((False, False), (False, True), (True, True), (False, True))
    for fused, streams in configs:
        ops.clear_weight_cache()
        m = build(cfg, case_params(cfg, spec)).eval()
        params = list(dict.fromkeys(m.parameters()))
        if fused:
            flat = torch.zeros(sum(p.numel() for p in params), device="cuda")
            o = 0
            for p in params:
                p.grad = flat[o:o + p.numel()].view_as(p)
                o += p.numel()
        ops.set_grad_fusion(fused)
        dec.set_multi_stream(streams)
        try:
            out, vis, txt = run_model(m, inp, grad=True)
            total, _ = STGLossPlan(cfg, tg["boxes"], tg["actioness"], spec["durations"], "cuda")(out)
            total.backward()
            torch.cuda.synchronize()
        finally:
            ops.set_grad_fusion(False)
            dec.set_multi_stream(True)
        g = {k: (None if p.grad is None else p.grad.clone()) for k, p in m.named_parameters()}
        g["__vis"] = vis.grad.clone()
        g["__loss"] = total.detach().reshape(1)
        results.append(g)
    # fp32: summation-order noise only.  bf16: the backward of this network amplifies rounding noise by ~5e4 (the
    # reference's own fp32 gradients sit 2e-3 from its fp64 gradients), so run-to-run differences in atomic /
    # reduce-add order move bf16-mode gradients by a few percent; the forward (loss) is deterministic and gated tight.
    tol = 1e-4 if precision == "fp32" else 1.5e-1
    for other in results[1:]:
        assert abs(float(other["__loss"]) - float(results[0]["__loss"])) <= 1e-5 * abs(float(results[0]["__loss"]))
    for ci, other in enumerate(results[1:], 1):
        for k, g0 in results[0].items():
            g1 = other[k]
            if g0 is None:
                assert g1 is None or float(g1.abs().max()) == 0.0, (configs[ci], k)
            else:
                assert rel_err(g1, g0) < tol or float((g1 - g0).abs().max()) < 1e-6, (configs[ci], k, rel_err(g1, g0))


@pytest.mark.parametrize("name", ["b1_T8_res224_L8", "b2_ragged_T5_3"])
def test_fused_glue_matches_the_cat_slice_composition_on_device(name):
    """bf16 mode: ops.token_assembly / mem_operands / template (csrc/assembly.cu) against the torch.cat / slice / Linear
    composition they replace (STCAT_FUSED_GLUE=0): loss, parameter gradients and the input gradients.  Tolerances as in
    test_grad_fusion_and_single_stream_agree: the forward differs by fp32 summation order inside the template generator only."""
    from stcat_b200 import encoder as enc
    from stcat_b200.loss import STGLossPlan

    fx = load_golden(name)
    spec = fx["spec"]
    cfg = cfg_for(spec)
    inp = case_inputs(spec)
    tg = synthetic.make_targets(spec["durations"], seed=spec["seed"])
    ops.set_precision("bf16")
    results = []
    for fused in (False, True):
        enc.set_fused_glue(fused)
        ops.clear_weight_cache()
        try:
            m = build(cfg, case_params(cfg, spec)).eval()
            out, vis, txt = run_model(m, inp, grad=True)
            total, _ = STGLossPlan(cfg, tg["boxes"], tg["actioness"], spec["durations"], "cuda")(out)
            total.backward()
            torch.cuda.synchronize()
        finally:
            enc.set_fused_glue(True)
        g = {k: (None if p.grad is None else p.grad.clone()) for k, p in m.named_parameters()}
        g["__vis"], g["__txt"] = vis.grad.clone(), txt.grad.clone()
        g["__loss"] = total.detach().reshape(1)
        g["__boxes"] = out["pred_boxes"].detach().clone()
        results.append(g)
    a, b_ = results
    assert abs(float(a["__loss"]) - float(b_["__loss"])) <= 2e-3 * abs(float(a["__loss"]))
    assert rel_err(b_["__boxes"], a["__boxes"]) < 3e-3
    for k, g0 in a.items():
        g1 = b_[k]
        if g0 is None:
            assert g1 is None or float(g1.abs().max()) == 0.0, k
        else:
            assert g1 is not None, k
            assert rel_err(g1, g0) < 1.5e-1 or float((g1 - g0).abs().max()) < 1e-6, (k, rel_err(g1, g0))


@pytest.mark.parametrize("durs,aux", [([12], True), ([7, 12, 3], True), ([9, 4], False)])
def test_fused_loss_kernel_matches_torch_restatement(durs, aux):
    """stcat_stg_loss (values + gradients of all layers in one launch) vs the torch restatement of VideoSTGLoss on the same
    random predictions: boxes L1 + GIoU, start/end KL, guided attention, actioness BCE, ragged durations."""
    from stcat_b200.loss import STGLossPlan

    cfg = cfg_for({"max_video_len": 20})
    cfg.merge_from_list(["SOLVER.GIOU_COEF", 3, "SOLVER.TEMP_COEF", 10, "SOLVER.EOS_COEF", 0.3, "SOLVER.USE_AUX_LOSS", aux])
    b, t, nl = len(durs), max(durs), 6
    tg = synthetic.make_targets(durs, seed=3)
    plan = STGLossPlan(cfg, tg["boxes"], tg["actioness"], durs, "cuda")
    gen = torch.Generator().manual_seed(5)
    mk = lambda *shape: torch.randn(*shape, generator=gen)
    leaves = {"_coord_all": torch.sigmoid(mk(nl, b * t, 4)), "_sted_all": 2 * mk(nl, b, t, 2), "_act_all": mk(nl, b, t, 1),
              "_weights_all": torch.softmax(mk(nl, b, t, t), -1)}
    res = []
    for fused in (True, False):
        out = {k: v.clone().cuda().requires_grad_(True) for k, v in leaves.items()}
        out["_hs"] = torch.zeros(nl, 1, device="cuda")
        total, named = plan(out) if fused else plan.torch_restatement(out)
        total.backward()
        res.append((total.detach(), {k: v.detach() for k, v in named.items()}, {k: out[k].grad for k in leaves}))
    (tf, nf, gf), (tt, nt, gt) = res
    assert abs(float(tf) - float(tt)) <= 2e-5 * abs(float(tt))
    assert set(nf) == set(nt)
    for k in nt:
        assert abs(float(nf[k]) - float(nt[k])) <= 2e-5 * max(1.0, abs(float(nt[k]))), k
    for k in leaves:
        ref = gt[k] if gt[k] is not None else torch.zeros_like(gf[k])
        assert rel_err(gf[k], ref) < 1e-4, k


def test_train_mode_dropout_matches_host_emulation():
    """Train mode with the reference's default MODEL.STCAT.DROPOUT 0.1 (+ the 0.3 of the temporal heads): the dropout masks
    are a counter-based stream (stcat_dropout), so the CUDA path and the torch emulation of the C ABI (tests/emu_backend.py,
    CPU) draw bit-identical masks for the same seed and must agree on the loss (1e-3) and on the gradients (1e-2, see the
    module docstring) -- which pins every dropout site's forward scaling and its backward."""
    from emu_backend import EmuBackend
    from stcat_b200.loss import STGLossPlan
    from stcat_b200.pipeline import STCATHotPath

    spec = load_golden("b2_ragged_T5_3")["spec"]
    cfg = cfg_for(spec, dropout=0.1)
    inp = case_inputs(spec)
    tg = synthetic.make_targets(spec["durations"], seed=spec["seed"])
    P = case_params(cfg, spec)

    def run(dev, seed):
        ops.clear_weight_cache()
        m = STCATHotPath(cfg).load_flat_params(P).to(dev).train()
        ops.set_dropout_seed(seed)
        mv = lambda x: x.to(dev)
        vis = mv(inp["vis_features"]).requires_grad_(True)
        out = m(NestedTensor(vis, mv(inp["vis_mask"]), inp["durations"]), mv(inp["vis_pos"]),
                (mv(inp["text_mask"]), mv(inp["text_memory"]), None))
        plan = STGLossPlan(cfg, tg["boxes"], tg["actioness"], spec["durations"], dev)
        total = plan(out)[0] if dev == "cuda" else plan.torch_restatement(out)[0]
        total.backward()
        g = {k: p.grad.detach().cpu() for k, p in m.named_parameters() if p.grad is not None}
        g["__vis"] = vis.grad.detach().cpu()
        return float(total), g

    l_gpu, g_gpu = run("cuda", 5)
    l_gpu2, _ = run("cuda", 5)
    l_gpu3, _ = run("cuda", 6)
    assert l_gpu == l_gpu2 and l_gpu != l_gpu3  # the seed fixes the masks
    ops.set_backend(EmuBackend())
    try:
        l_cpu, g_cpu = run("cpu", 5)
    finally:
        ops.set_backend(None)
    assert abs(l_gpu - l_cpu) < TOL * abs(l_cpu), (l_gpu, l_cpu)
    assert g_gpu.keys() == g_cpu.keys()
    for k in g_cpu:
        assert rel_err(g_gpu[k], g_cpu[k]) < 1e-2 or float((g_gpu[k] - g_cpu[k]).abs().max()) < 1e-6, (k, rel_err(g_gpu[k], g_cpu[k]))


@pytest.mark.parametrize("variant", ["conv", "attn"])
def test_map2d_head_matches_reference_fixture(variant):
    """The 2-D temporal proposal head (map2d_head.py) on the B200 through the C ABI (stcat_map2d_pool, GEMMs
    over the im2col of the map / packed in-projections, attention over map rows and columns, LayerNorm, FFN) against the
    scores of the unmodified reference head: train-mode logits and eval-mode sigmoid * mask, exact-fp32 mode, 1e-3."""
    from stcat_b200.config import get_default_cfg
    from stcat_b200.map2d import TempPredictionHead
    from stcat_b200.synthetic import fill_param

    fx = load_golden("map2d_N16" if variant == "conv" else "map2d_attn_N16")
    c = fx["cfg"]
    cfg = get_default_cfg()
    extra = (["MODEL.STCAT.TEMP_HEAD", "conv", "MODEL.STCAT.KERNAL_SIZE", c["KERNAL_SIZE"], "MODEL.STCAT.CONV_LAYERS", c["CONV_LAYERS"]]
             if variant == "conv" else ["MODEL.STCAT.TEMP_HEAD", "attn", "MODEL.STCAT.TEMP_PRED_LAYERS", c["TEMP_PRED_LAYERS"]])
    cfg.merge_from_list(["MODEL.STCAT.MAX_MAP_SIZE", c["MAX_MAP_SIZE"], "MODEL.STCAT.POOLING_COUNTS", c["POOLING_COUNTS"],
                         "MODEL.STCAT.DROPOUT", 0.0] + extra)
    head = TempPredictionHead(cfg)
    head.load_state_dict({k: fill_param(k, tuple(v.shape), fx["seed"]) for k, v in head.state_dict().items()})
    head = head.cuda()
    before = ops.get_backend().launches
    with torch.no_grad():
        head.train()
        assert rel_err(head(fx["x"].cuda()), fx["scores_train"]) < TOL
        head.eval()
        assert rel_err(head(fx["x"].cuda()), fx["scores_eval"]) < TOL
    assert ops.get_backend().launches > before
