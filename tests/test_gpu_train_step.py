"""-m gpu: stcat_b200.train.GraphedStep -- the training step as one replayed CUDA graph -- and the device-resident dropout step
counter (stcat_set_dropout_step) that gives every replay fresh masks."""
import pytest
import torch

from helpers import cfg_for
from stcat_b200 import ops

pytestmark = pytest.mark.gpu


def _setup(dropout, with_opt, lr=None, train=True):
    from stcat_b200 import synthetic
    from stcat_b200.dp import FlatGrads, hot_path_groups
    from stcat_b200.loss import STGLossPlan
    from stcat_b200.nested import NestedTensor
    from stcat_b200.optim import make_optimizer
    from stcat_b200.param_spec import synthetic_params
    from stcat_b200.pipeline import STCATHotPath

    cfg = cfg_for({"max_video_len": 16})
    cfg.merge_from_list(["MODEL.STCAT.DROPOUT", float(dropout)])
    if lr is not None:
        cfg.merge_from_list(["SOLVER.BASE_LR", lr, "SOLVER.TEMP_LR", lr, "SOLVER.MAX_GRAD_NORM", 1.0])
    T = 6
    model = STCATHotPath(cfg).load_flat_params(synthetic_params(cfg, seed=0)).cuda().train(train)
    inp = synthetic.make_inputs([T], 8, 8, 6, seed=3)
    tg = synthetic.make_targets([T], seed=3)
    plan = STGLossPlan(cfg, tg["boxes"], tg["actioness"], [T], "cuda")
    grads = FlatGrads(model, hot_path_groups(model))
    ops.set_grad_fusion(True)
    opt = make_optimizer(cfg, model, grads) if with_opt else None
    vmask, tmask = inp["vis_mask"].cuda(), inp["text_mask"].cuda()

    def fn(vis, pos, txt):
        grads.zero()
        out = model(NestedTensor(vis, vmask, [T]), pos, (tmask, txt, None))
        total, _ = plan(out)
        total.backward()
        if opt is not None:
            opt.step()
        return total

    ex = {"vis": inp["vis_features"].cuda().requires_grad_(True), "pos": inp["vis_pos"].cuda(), "txt": inp["text_memory"].cuda().requires_grad_(True)}
    return fn, ex, grads, model


@pytest.fixture(autouse=True)
def bf16():
    ops.set_backend(None)
    ops.set_precision("bf16")
    ops.clear_weight_cache()
    yield
    ops.set_grad_fusion(False)
    ops.set_shadow_provider(None)
    ops.get_backend().set_dropout_step(None)
    ops.set_precision("fp32")
    ops.clear_weight_cache()


def test_replay_without_dropout_is_bit_reproducible_in_the_loss():
    from stcat_b200.train import GraphedStep

    fn, ex, grads, _ = _setup(0.0, False)
    # the heads' hard-wired 0.3 dropout (pipeline.py:42-47) is active in train mode: freeze the counter to compare replays
    gs = GraphedStep(fn, ex, warmup=2)
    assert gs.graph is not None
    c0 = int(gs.counter)
    a = float(gs.replay()); g1 = grads.buf.clone()
    gs.counter.fill_(c0)
    b = float(gs.replay())
    assert abs(a - b) <= 1e-5 * abs(a)
    # run-to-run noise of the bf16 path (fp32 atomics + a bf16 rounding flip downstream), measured against the gradient scale
    assert float((grads.buf - g1).abs().max()) <= 1e-2 * float(g1.abs().max())
    gs.close()


def test_replay_draws_fresh_dropout_masks_and_is_reproducible_per_counter():
    from stcat_b200.train import GraphedStep

    ops.set_dropout_seed(1234)
    fn, ex, grads, _ = _setup(0.1, False)
    gs = GraphedStep(fn, ex, warmup=2)
    assert gs.graph is not None
    c0 = int(gs.counter)
    l1 = float(gs.replay())
    l2 = float(gs.replay())
    assert int(gs.counter) == c0 + 2
    assert l1 != l2                      # fresh masks on every replay (kernel arguments are frozen in the graph)
    gs.counter.fill_(c0)
    assert float(gs.replay()) == l1      # ... a pure function of the counter
    new = {k: v.detach() * 0.5 for k, v in ex.items()}
    l3 = float(gs(**new))                # new inputs are copied into the static tensors
    assert l3 != l1 and l3 == l3
    gs.close()


def test_replayed_step_with_fused_optimizer_trains():
    from stcat_b200.train import GraphedStep

    fn, ex, grads, model = _setup(0.0, True, lr=3e-4)
    w0 = model.ground_decoder.decoder.layers[0].linear1.weight.detach().clone()
    gs = GraphedStep(fn, ex, warmup=1)
    gs.counter.fill_(0)
    losses = []
    for _ in range(16):
        gs.counter.fill_(0)  # same head-dropout masks every step: the loss sequence reflects the weight updates only
        losses.append(float(gs.replay()))
    assert all(l == l for l in losses)
    assert not torch.equal(model.ground_decoder.decoder.layers[0].linear1.weight, w0)
    assert min(losses[-4:]) < losses[0], losses  # the same clip, 16 AdamW steps at lr 3e-4: the loss goes down
    gs.close()


def test_leaf_streams_give_the_same_gradients():
    """ops.set_leaf_streams: weight / bias gradients issued on side streams forked from the backward chain -- same gradients as
    the in-order backward, eagerly and under graph replay."""
    from stcat_b200.train import GraphedStep

    fn, ex, grads, _ = _setup(0.0, False, train=False)  # eval mode: no dropout anywhere, eager and replayed steps comparable
    ops.get_backend().set_dropout_step(None)
    try:
        ops.set_dropout_seed(7)
        fn(**ex)
        torch.cuda.synchronize()
        ref = grads.buf.clone()
        ops.set_leaf_streams(True)
        ops.set_dropout_seed(7)
        fn(**ex)
        assert len(ops.leaf_streams()) >= 1  # pending side streams with weight-gradient work
        ops.join_leaf_streams()
        assert not ops.leaf_streams()
        torch.cuda.synchronize()
        scale = float(ref.abs().max())
        # run-to-run noise of the bf16 path: fp32 atomics (head-averaged attention weights in the forward, split-K weight
        # gradients) change last bits, a bf16 rounding flip downstream of them moves single small elements by ~2e-3 of the
        # largest gradient; a mis-ordered leaf stream would show up at O(1)
        assert float((grads.buf - ref).abs().max()) <= 1e-2 * scale
        def fn_joined(**kw):
            total = fn(**kw)
            ops.join_leaf_streams()
            return total
        gs = GraphedStep(fn_joined, ex, warmup=1)
        gs.counter.fill_(0)
        a = float(gs.replay())
        torch.cuda.synchronize()
        g1 = grads.buf.clone()
        gs.counter.fill_(0)
        b = float(gs.replay())
        torch.cuda.synchronize()
        assert abs(a - b) <= 1e-5 * abs(a) and float((grads.buf - g1).abs().max()) <= 1e-2 * scale
        assert float((g1 - ref).abs().max()) <= 1e-2 * scale  # the replayed step equals the eager in-order step
        gs.close()
    finally:
        ops.set_leaf_streams(False)
