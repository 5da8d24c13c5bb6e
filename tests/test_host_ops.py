"""CPU (-m "not gpu"): the hand-written autograd functions of stcat_b200/ops.py in isolation, driven through the torch
emulation of the C ABI (tests/emu_backend.py) and compared with plain torch autograd: grouped multi-term Linears,
the frame-CLS row exchange, the flat gradient buffer layout, and the reference arm's JSON line."""
import json
import os
import subprocess
import sys

import pytest
import torch

from emu_backend import EmuBackend
from helpers import rel_err
from stcat_b200 import ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(autouse=True)
def emu():
    ops.set_backend(EmuBackend())
    ops.set_precision("fp32")
    yield
    ops.set_grad_fusion(False)
    ops.set_backend(None)
    ops.set_precision("fp32")
    ops.clear_weight_cache()


def _leaf(*shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).requires_grad_(True)


@pytest.mark.parametrize("fused", [False, True])
def test_linear_group_matches_torch_autograd(fused):
    """q = Wa x0 + Wb x1 + Wc x2, k = Wd x0 + We x2 (ReLU), v = rows [4:12] of a packed weight applied to x1: outputs and all
    gradients (inputs read by several terms, packed-weight row slices, shared biases) vs torch autograd."""
    M, d = 9, 16
    xs = [_leaf(M, d, seed=i) for i in range(3)]
    ws = [_leaf(d, d, seed=10 + i) for i in range(5)]
    bs = [_leaf(d, seed=20 + i) for i in range(5)]
    wp, bp = _leaf(12, d, seed=30), _leaf(12, seed=31)
    gq, gk, gv = (torch.randn(M, d, generator=torch.Generator().manual_seed(40 + i)) for i in range(2)), None, None
    gq, gk = list(gq)
    gv = torch.randn(M, 8, generator=torch.Generator().manual_seed(43))

    def ref():
        q = xs[0] @ ws[0].t() + bs[0] + xs[1] @ ws[1].t() + bs[1] + xs[2] @ ws[2].t() + bs[2]
        k = (xs[0] @ ws[3].t() + bs[3] + xs[2] @ ws[4].t() + bs[4]).relu()
        v = xs[1] @ wp[4:12].t() + bp[4:12]
        return q, k, v

    q, k, v = ref()
    ((q * gq).sum() + (k * gk).sum() + (v * gv).sum()).backward()
    want = [t.grad.clone() for t in (*xs, *ws, *bs, wp, bp)]
    for t in (*xs, *ws, *bs, wp, bp):
        t.grad = None
    if fused:  # wgrad accumulates straight into pre-existing .grad buffers
        for t in (*ws, *bs, wp, bp):
            t.grad = torch.zeros_like(t)
        ops.set_grad_fusion(True)
    q2, k2, v2 = ops.linear_group(
        [(x, None) for x in xs],
        [{"terms": [(0, ws[0], bs[0]), (1, ws[1], bs[1]), (2, ws[2], bs[2])]},
         {"terms": [(0, ws[3], bs[3]), (2, ws[4], bs[4])], "relu": True},
         {"terms": [(1, wp, bp, (4, 12))]}])
    for a, b in ((q2, q), (k2, k), (v2, v)):
        assert rel_err(a, b) < 1e-6
    ((q2 * gq).sum() + (k2 * gk).sum() + (v2 * gv).sum()).backward()
    for t, w_ in zip((*xs, *ws, *bs, wp, bp), want):
        assert t.grad is not None and rel_err(t.grad, w_) < 1e-5


def test_row_exchange_matches_inplace_autograd():
    """take_rows / put_rows (the frame-CLS exchange, modal_encoder.py:170-195) vs the same computation written with
    autograd's select + in-place row assignment."""
    n, S, d = 5, 7, 4
    x0 = _leaf(n, S, d, seed=1)
    w = _leaf(d, d, seed=2)
    gout = torch.randn(n, S, d, generator=torch.Generator().manual_seed(3))

    def tail(x):  # something that mixes all rows, so every gradient path is exercised
        return (x * x.sum(1, keepdim=True)).tanh()

    xa = x0 * 1.5
    rows = xa[:, 0, :]
    new = (rows @ w.t()).sin()
    xb = xa.clone()
    xb[:, 0, :] = new
    (tail(xb) * gout).sum().backward()
    want = (x0.grad.clone(), w.grad.clone())
    x0.grad = w.grad = None

    xa = x0 * 1.5
    base, rows = ops.take_rows(xa, 0)
    new = (rows @ w.t()).sin()
    xb = ops.put_rows(base, new, 0)
    (tail(xb) * gout).sum().backward()
    assert rel_err(x0.grad, want[0]) < 1e-6 and rel_err(w.grad, want[1]) < 1e-6


def test_flat_grads_layout():
    from stcat_b200.dp import FlatGrads

    m = torch.nn.ModuleDict({"a": torch.nn.Linear(3, 5), "b": torch.nn.Linear(5, 1), "c": torch.nn.Linear(7, 2)})
    groups = [("late", list(m["c"].parameters())), ("early", list(m["a"].parameters()))]
    fg = FlatGrads(m, groups)
    assert list(fg.ranges) == ["late", "early", "rest"]
    for p in m.parameters():
        assert p.grad.untyped_storage().data_ptr() == fg.buf.untyped_storage().data_ptr()
        assert (p.grad.data_ptr() - fg.buf.data_ptr()) % (4 * FlatGrads.ALIGN) == 0  # 256-byte aligned views (TMA reduce-add)
    lo, hi = fg.ranges["late"]
    assert lo == 0 and all(lo <= (p.grad.data_ptr() - fg.buf.data_ptr()) // 4 < hi for p in m["c"].parameters())
    m["a"].weight.grad.fill_(2.0)
    lo, hi = fg.ranges["early"]
    assert float(fg.buf[lo:hi].sum()) == 2.0 * m["a"].weight.numel() and float(fg.buf.sum()) == 2.0 * m["a"].weight.numel()
    fg.zero()
    assert float(m["a"].weight.grad.abs().sum()) == 0.0


def test_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the CPU arm: oracle port of the reference) on a tiny clip: exactly one JSON line on
    stdout with the contract's keys."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--T", "4", "--res", "96", "--L", "4"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "clips/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


# ---------------------------------------------------------------------------------------------------------------------
# train-mode dropout (counter-based masks regenerated in backward)
# ---------------------------------------------------------------------------------------------------------------------
def test_dropout_mask_statistics_and_backward_consistency():
    ops.set_dropout_seed(7)
    x = torch.ones(200, 256, requires_grad=True)
    y = ops.dropout(x, 0.3)
    kept = (y != 0).float().mean().item()
    assert abs(kept - 0.7) < 0.01 and abs(y.mean().item() - 1.0) < 0.02
    assert torch.allclose(y[y != 0], torch.full_like(y[y != 0], 1 / 0.7), rtol=1e-6)
    g = torch.randn(200, 256, generator=torch.Generator().manual_seed(1))
    y.backward(g)
    assert torch.equal(x.grad != 0, y != 0)  # the backward mask is the forward mask
    assert torch.allclose(x.grad, g * y.detach(), rtol=1e-6)  # y = mask * scale on ones
    y2 = ops.dropout(x, 0.3)  # the next site draws a different part of the stream
    assert not torch.equal(y2 != 0, y != 0)
    ops.set_dropout_seed(7)
    assert torch.equal(ops.dropout(x, 0.3), y)  # same seed -> same mask
    assert ops.dropout(x, 0.0) is x


def test_attention_dropout_matches_explicit_mask():
    from emu_backend import drop_keep_scale

    B, H, Lq, Lk, p = 2, 8, 5, 7, 0.25
    E = H * 32
    mk = lambda L, s: _leaf(B * L, E, seed=s)
    q, k, v = mk(Lq, 1), mk(Lk, 2), mk(Lk, 3)
    go = torch.randn(B * Lq, E, generator=torch.Generator().manual_seed(4))
    gw = torch.randn(B, Lq, Lk, generator=torch.Generator().manual_seed(5))
    ops.set_dropout_seed(99)
    o, w = ops.attention(q, k, v, B, H, Lq, Lk, 32 ** -0.5, need_pavg=True, drop_p=p)
    ((o * go).sum() + (w * gw).sum()).backward()
    got = (o.detach(), w.detach(), q.grad.clone(), k.grad.clone(), v.grad.clone())
    q.grad = k.grad = v.grad = None
    keep, sc = drop_keep_scale(B * H * Lq * Lk, p, 99, 0)
    m = keep.reshape(B, H, Lq, Lk).float() * sc
    hd = lambda t, L: t.reshape(B, L, H, 32).permute(0, 2, 1, 3)
    pr = torch.softmax(hd(q, Lq) @ hd(k, Lk).transpose(-1, -2) * 32 ** -0.5, -1) * m
    o_ref = (pr @ hd(v, Lk)).permute(0, 2, 1, 3).reshape(B * Lq, E)
    w_ref = pr.mean(1)
    ((o_ref * go).sum() + (w_ref * gw).sum()).backward()
    for a, b in zip(got, (o_ref, w_ref, q.grad, k.grad, v.grad)):
        assert rel_err(a, b) < 1e-5


def _masks(sizes, p, seed):
    """the masks consecutive dropout sites of one forward draw (same reservation order as ops._drop_reserve)"""
    from emu_backend import drop_keep_scale

    out, off = [], 0
    for n in sizes:
        keep, sc = drop_keep_scale(n, p, seed, off)
        out.append(keep.float() * sc)
        off += n
    return out


def test_ffn_block_dropout_matches_explicit_masks():
    R, d, F_, p = 6, 256, 512, 0.2
    x = _leaf(R, d, seed=1)
    w1, b1 = _leaf(F_, d, seed=2), _leaf(F_, seed=3)
    w2, b2 = _leaf(d, F_, seed=4), _leaf(d, seed=5)
    gm, bt = _leaf(d, seed=6), _leaf(d, seed=7)
    with torch.no_grad():
        w1.mul_(d ** -0.5); w2.mul_(F_ ** -0.5)
    go = torch.randn(R, d, generator=torch.Generator().manual_seed(8))
    leaves = (x, w1, b1, w2, b2, gm, bt)
    ops.set_dropout_seed(21)
    y, _ = ops.ffn_block(x, None, w1, b1, w2, b2, gm, bt, drop_p=p)
    (y * go).sum().backward()
    got = [y.detach()] + [t.grad.clone() for t in leaves]
    for t in leaves:
        t.grad = None
    mh, mo = _masks([R * F_, R * d], p, 21)
    h = (x @ w1.t() + b1).relu() * mh.view(R, F_)
    yl = (h @ w2.t() + b2) * mo.view(R, d)
    yr = torch.nn.functional.layer_norm(x + yl, (d,), gm, bt, 1e-5)
    (yr * go).sum().backward()
    for a_, b_ in zip(got, [yr] + [t.grad for t in leaves]):
        assert rel_err(a_, b_) < 2e-5


def test_self_attn_block_dropout_matches_explicit_masks():
    B, L, H, d, p = 3, 5, 8, 256, 0.2
    R = B * L
    x, pos = _leaf(R, d, seed=1), _leaf(R, d, seed=2)
    wi, bi = _leaf(3 * d, d, seed=3), _leaf(3 * d, seed=4)
    wo, bo = _leaf(d, d, seed=5), _leaf(d, seed=6)
    gm, bt = _leaf(d, seed=7), _leaf(d, seed=8)
    with torch.no_grad():
        wi.mul_(d ** -0.5); wo.mul_(d ** -0.5)
    go = torch.randn(R, d, generator=torch.Generator().manual_seed(9))
    leaves = (x, pos, wi, bi, wo, bo, gm, bt)
    ops.set_dropout_seed(33)
    y, _ = ops.self_attn_block(x, None, pos, None, wi, bi, wo, bo, gm, bt, B, L, H, drop_p=p)
    (y * go).sum().backward()
    got = [y.detach()] + [t.grad.clone() for t in leaves]
    for t in leaves:
        t.grad = None
    ma, mo = _masks([B * H * L * L, R * d], p, 33)
    qk = x + pos
    q, k, v = qk @ wi[:d].t() + bi[:d], qk @ wi[d:2 * d].t() + bi[d:2 * d], x @ wi[2 * d:].t() + bi[2 * d:]
    hd = lambda t: t.reshape(B, L, H, 32).permute(0, 2, 1, 3)
    pr = torch.softmax(hd(q) @ hd(k).transpose(-1, -2) * 32 ** -0.5, -1) * ma.view(B, H, L, L)
    o = (pr @ hd(v)).permute(0, 2, 1, 3).reshape(R, d)
    a = (o @ wo.t() + bo) * mo.view(R, d)
    yr = torch.nn.functional.layer_norm(x + a, (d,), gm, bt, 1e-5)
    (yr * go).sum().backward()
    for a_, b_ in zip(got, [yr] + [t.grad for t in leaves]):
        assert rel_err(a_, b_) < 2e-5


def test_hot_path_train_mode_dropout():
    """The whole hot path in train mode with MODEL.STCAT.DROPOUT 0.1 (the reference's default) runs forward and backward,
    is reproducible for a fixed seed, changes with the seed, and eval mode ignores dropout."""
    from helpers import cfg_for
    from stcat_b200 import synthetic
    from stcat_b200.loss import STGLossPlan
    from stcat_b200.nested import NestedTensor
    from stcat_b200.param_spec import synthetic_params
    from stcat_b200.pipeline import STCATHotPath

    cfg = cfg_for({"max_video_len": 16}, dropout=0.1)
    model = STCATHotPath(cfg).load_flat_params(synthetic_params(cfg, seed=3))
    T = 4
    inp = synthetic.make_inputs([T], 2, 3, 3, seed=5)
    tg = synthetic.make_targets([T], seed=5)
    plan = STGLossPlan(cfg, tg["boxes"], tg["actioness"], [T], "cpu")

    def loss_of(vis, seed=11):
        ops.set_dropout_seed(seed)
        torch.manual_seed(0)  # the 0.3 dropout of the temporal heads draws from torch's stream
        out = model(NestedTensor(vis, inp["vis_mask"].clone(), [T]), inp["vis_pos"], (inp["text_mask"], inp["text_memory"], None))
        return plan.torch_restatement(out)[0]

    model.train()
    vis = inp["vis_features"].clone().requires_grad_(True)
    L0 = loss_of(vis)
    L0.backward()
    assert torch.isfinite(vis.grad).all() and float(vis.grad.abs().max()) > 0
    for name, prm in model.named_parameters():
        assert prm.grad is None or torch.isfinite(prm.grad).all(), name
    with torch.no_grad():
        assert float(loss_of(vis)) == float(L0)                 # same seed, same masks
        assert float(loss_of(vis, seed=12)) != float(L0)        # another seed, other masks
        model.eval()
        e1, e2 = float(loss_of(vis, seed=1)), float(loss_of(vis, seed=2))
        assert e1 == e2 and e1 != float(L0)  # eval mode: no dropout anywhere


def test_fused_adamw_matches_torch_adamw_clip_and_ema():
    """optim.FusedAdamW (one stcat_sumsq + one stcat_adamw_step per (lr, wd) run, emulated here) against the reference's
    sequence: clip_grad_norm_ -> torch.optim.AdamW.step -> update_ema, over several steps, two LR groups, a parameter that
    never receives a gradient (skipped by torch), a parameter optimised elsewhere that takes part in the norm, and the bf16
    shadows served to ops."""
    import copy

    from stcat_b200.dp import FlatGrads
    from stcat_b200.optim import FusedAdamW

    torch.manual_seed(0)
    net = torch.nn.ModuleDict({"a": torch.nn.Linear(40, 24), "b": torch.nn.Linear(24, 7), "never": torch.nn.Linear(5, 3)})
    outside = torch.nn.Parameter(torch.randn(11))
    ref = copy.deepcopy(net)
    ref_out = torch.nn.Parameter(outside.detach().clone())
    ref_ema = copy.deepcopy(ref)
    ema_model = copy.deepcopy(net)
    groups = lambda m: [{"params": list(m["a"].parameters()) + list(m["never"].parameters())},
                        {"params": list(m["b"].parameters()), "lr": 3e-3, "weight_decay": 0.0}]
    flat = FlatGrads(net)
    opt = FusedAdamW(flat, groups(net), lr=1e-3, weight_decay=1e-2, max_grad_norm=0.1, ema_decay=0.99,
                     frozen=list(net["never"].parameters()), extra_norm_params=[outside])
    opt.attach_ema(net, ema_model)
    ops.set_precision("bf16")
    try:
        opt.enable_shadows()
        topt = torch.optim.AdamW(groups(ref), lr=1e-3, weight_decay=1e-2)
        for step in range(4):
            x = torch.randn(16, 40, generator=torch.Generator().manual_seed(step))
            for m, o in ((net, outside), (ref, ref_out)):
                if m is net:
                    opt.zero_grad()
                    o.grad = None
                else:
                    topt.zero_grad(set_to_none=True)
                    o.grad = None
                loss = (m["b"](torch.relu(m["a"](x))) ** 2).sum() * 50 + (o ** 2).sum()
                loss.backward()
            opt.step()
            torch.nn.utils.clip_grad_norm_(list(ref.parameters()) + [ref_out], 0.1)
            topt.step()
            with torch.no_grad():
                for k, e in ref_ema.state_dict().items():
                    e.copy_(e * 0.99 + 0.01 * ref.state_dict()[k])
            if step == 1:  # the LR schedule writes param_groups[i]["lr"] (train_net.py:142)
                for o_ in (opt, topt):
                    o_.param_groups[0]["lr"] = 5e-4
        for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
            assert rel_err(p, q) < 2e-6, k
        for (k, p), (_, q) in zip(ema_model.named_parameters(), ref_ema.named_parameters()):
            assert rel_err(p, q) < 2e-6, k
        assert torch.equal(net["never"].weight, ref["never"].weight)  # untouched, like a grad-less parameter under torch
        w = net["a"].weight
        assert torch.equal(ops._operand(w, is_weight=True), w.detach().to(torch.bfloat16))  # shadow served, up to date
        assert torch.equal(ops._operand(w[3:9], is_weight=True), w[3:9].detach().to(torch.bfloat16))
        sd = opt.state_dict()
        assert rel_err(sd["state"][0]["exp_avg"], topt.state_dict()["state"][0]["exp_avg"]) < 2e-6
        opt.load_state_dict(sd)
        assert opt.state[net["a"].weight]["exp_avg"].data_ptr() == opt.m.data_ptr()
    finally:
        ops.set_shadow_provider(None)
        ops.set_precision("fp32")


def test_weight_cache_and_shadows_follow_the_masters():
    """(i) the bf16 cast cache must not serve a freed model's weights to a new Parameter allocated at the same
    address; (ii) FusedAdamW.step without shadows must invalidate it (raw-pointer update, no version bump); (iii) with
    shadows, writes that are not step() -- load_state_dict after the optimizer was built, the reference's resume order --
    must be seen; (iv) extra_norm_params' gradients are scaled by the clip coefficient."""
    from stcat_b200.dp import FlatGrads
    from stcat_b200.optim import FusedAdamW

    ops.set_precision("bf16")
    try:
        ops.clear_weight_cache()
        stale = 0
        for trial in range(20):  # (i)
            w = torch.nn.Parameter(torch.randn(64, 64))
            a = ops._operand(w.detach(), True)
            assert torch.equal(a, w.detach().to(torch.bfloat16))
            ptr = w.data_ptr()
            del w, a
            w2 = torch.nn.Parameter(torch.randn(64, 64))
            if w2.data_ptr() == ptr:
                stale += 1
            assert torch.equal(ops._operand(w2.detach(), True), w2.detach().to(torch.bfloat16))
            del w2
        torch.manual_seed(0)
        net = torch.nn.Linear(16, 8)
        outside = torch.nn.Parameter(torch.randn(5))
        flat = FlatGrads(net)
        opt = FusedAdamW(flat, list(net.parameters()), lr=1e-2, max_grad_norm=0.1, extra_norm_params=[outside])
        before = ops._operand(net.weight.detach(), True).clone()
        opt.zero_grad()
        ((net(torch.randn(4, 16)) ** 2).sum() * 10 + (outside ** 2).sum()).backward()
        g_out = outside.grad.clone()
        opt.step()  # (ii)
        now = ops._operand(net.weight.detach(), True)
        assert torch.equal(now, net.weight.detach().to(torch.bfloat16)) and not torch.equal(now, before)
        coef = float(opt.clip_coef())  # (iv)
        assert coef < 1.0 and torch.allclose(outside.grad, g_out * coef, rtol=1e-6)
        total = torch.sqrt(flat.buf.pow(2).sum() + g_out.pow(2).sum())
        assert abs(coef - 0.1 / (float(total) + 1e-6)) < 1e-6
        opt.enable_shadows()  # (iii)
        sd = {k: torch.randn_like(v) for k, v in net.state_dict().items()}
        net.load_state_dict(sd)
        assert torch.equal(ops._operand(net.weight.detach(), True), sd["weight"].to(torch.bfloat16))
        assert torch.equal(ops._operand(net.weight.detach()[2:5], True), sd["weight"][2:5].to(torch.bfloat16))
        with torch.no_grad():
            net.weight.mul_(2.0)
        opt.refresh_shadows()
        assert torch.equal(opt.shadow[:net.weight.numel()].view_as(net.weight), (sd["weight"] * 2).to(torch.bfloat16))
    finally:
        ops.set_shadow_provider(None)
        ops.set_precision("fp32")
        ops.clear_weight_cache()


def test_unused_parameters_are_exactly_the_gradless_ones():
    """optim.unused_parameters (the ``frozen`` list FusedAdamW needs) == the parameters the reference leaves without a
    gradient, for both FROM_SCRATCH settings."""
    from helpers import cfg_for
    from stcat_b200 import synthetic
    from stcat_b200.loss import STGLossPlan
    from stcat_b200.nested import NestedTensor
    from stcat_b200.optim import unused_parameters
    from stcat_b200.param_spec import synthetic_params
    from stcat_b200.pipeline import STCATHotPath

    for fs in (True, False):
        cfg = cfg_for({"max_video_len": 16, "from_scratch": fs})
        model = STCATHotPath(cfg).load_flat_params(synthetic_params(cfg, seed=0)).eval()
        T = 4
        inp = synthetic.make_inputs([T], 2, 3, 4, seed=1)
        tg = synthetic.make_targets([T], seed=1)
        plan = STGLossPlan(cfg, tg["boxes"], tg["actioness"], [T], "cpu")
        out = model(NestedTensor(inp["vis_features"], inp["vis_mask"], [T]), inp["vis_pos"], (inp["text_mask"], inp["text_memory"], None))
        total, _ = plan(out)
        total.backward()
        gradless = {k for k, p in model.named_parameters() if p.grad is None}
        listed = {k for k, p in model.named_parameters() if any(p is q for q in unused_parameters(model, cfg))}
        assert listed == gradless, (fs, sorted(listed ^ gradless))


def test_dropout_hash_restatement_matches_pure_python():
    """tests/emu_backend.drop_keep_scale (vectorised, int64 lanes) against a scalar restatement of csrc/common.cuh
    drop_bits24: 64-bit counter, folded to 32 bits, multiply-xorshift mixer, top 24 bits against p * 2^24."""
    from emu_backend import drop_keep_scale

    def bits24(seed, idx):
        z = (idx + seed * 0x9E3779B97F4A7C15) & ((1 << 64) - 1)
        x = (z & 0xFFFFFFFF) ^ (((z >> 32) * 0x9E3779B1) & 0xFFFFFFFF)
        x ^= x >> 16
        x = (x * 0x21F0AAAD) & 0xFFFFFFFF
        x ^= x >> 15
        x = (x * 0x735A2D97) & 0xFFFFFFFF
        x ^= x >> 15
        return x >> 8

    for p, seed, off in [(0.1, 0x1234567, 987654321), (0.3, (1 << 62) + 12345, (1 << 40) + 7), (0.5, 7, 0)]:
        keep, sc = drop_keep_scale(3000, p, seed, off)
        t = min(int(float(torch.tensor(p, dtype=torch.float32)) * 16777216.0), 16777215)
        assert torch.equal(keep, torch.tensor([bits24(seed, off + i) >= t for i in range(3000)]))
        assert abs(sc - 16777216.0 / (16777216.0 - t)) < 1e-12
    keep, _ = drop_keep_scale(2_000_000, 0.1, 99, 1 << 33)
    k = keep.float()
    assert abs(float(k.mean()) - 0.9) < 1e-3
    assert abs(float(((k[1:] - k.mean()) * (k[:-1] - k.mean())).mean() / k.var())) < 5e-3  # neighbours are uncorrelated


@pytest.mark.parametrize("block", ["ffn", "attn"])
def test_block_dropout_in_bf16_mode_uses_the_fused_layernorm_mask(block):
    """bf16 mode: the block-output dropout is folded into the LayerNorm kernels (stcat_layernorm_dropout_fwd/_bwd: mask on
    load, masked operand copy and bias column sums in the backward).  Same masks, same offsets as the unfused fp32 mode: the
    two modes agree to bf16 operand rounding, and no stand-alone dropout launch is left for that site."""
    R, d, p = 10, 256, 0.2
    B, L, H = 2, 5, 8
    go = torch.randn(R, d, generator=torch.Generator().manual_seed(8))

    def run(precision):
        ops.set_precision(precision)
        ops.clear_weight_cache()
        x, pos = _leaf(R, d, seed=1), _leaf(R, d, seed=2)
        gm, bt = _leaf(d, seed=6), _leaf(d, seed=7)
        if block == "ffn":
            w1, b1, w2, b2 = _leaf(512, d, seed=2), _leaf(512, seed=3), _leaf(d, 512, seed=4), _leaf(d, seed=5)
            with torch.no_grad():
                w1.mul_(d ** -0.5); w2.mul_(512 ** -0.5)
            leaves = (x, w1, b1, w2, b2, gm, bt)
            fn = lambda: ops.ffn_block(x, None, w1, b1, w2, b2, gm, bt, drop_p=p)
        else:
            wi, bi, wo, bo = _leaf(3 * d, d, seed=3), _leaf(3 * d, seed=4), _leaf(d, d, seed=5), _leaf(d, seed=6)
            with torch.no_grad():
                wi.mul_(d ** -0.5); wo.mul_(d ** -0.5)
            leaves = (x, wi, bi, wo, bo, gm, bt)
            fn = lambda: ops.self_attn_block(x, None, pos, None, wi, bi, wo, bo, gm, bt, B, L, H, drop_p=p)
        ops.set_dropout_seed(21)
        be = ops.get_backend()
        calls = []
        orig = be.dropout
        be.dropout = lambda *a, **k: (calls.append(a[0].shape), orig(*a, **k))[1]
        try:
            y, _ = fn()
            (y * go).sum().backward()
        finally:
            be.dropout = orig
        return [y.detach()] + [t.grad.clone() for t in leaves], calls

    try:
        ref, calls32 = run("fp32")
        got, calls16 = run("bf16")
    finally:
        ops.set_precision("fp32")
        ops.clear_weight_cache()
    # fp32 mode: stand-alone dropout on the [R, d] block output forward and on dz backward; bf16 mode: none of those
    assert sum(1 for s in calls32 if tuple(s) == (R, d)) == 2 and sum(1 for s in calls16 if tuple(s) == (R, d)) == 0
    for a_, b_ in zip(got, ref):
        assert rel_err(a_, b_) < 3e-2


@pytest.mark.parametrize("p", [0.0, 0.25])
def test_out_ln_equals_linear_dropout_layernorm(p):
    """ops.out_ln (out-projection + block-output dropout + residual + LayerNorm as one autograd node, the decoders' sites
    query_decoder.py:342-345, 430-432, 610-613, 652-654) == the composition of the stand-alone ops on the same mask offsets:
    outputs and every gradient."""
    R, d = 12, 256
    go = torch.randn(R, d, generator=torch.Generator().manual_seed(8))
    res = {}
    for fused in (True, False):
        o, tgt = _leaf(R, d, seed=1), _leaf(R, d, seed=2)
        w, b = _leaf(d, d, seed=3), _leaf(d, seed=4)
        gm, bt = _leaf(d, seed=5), _leaf(d, seed=6)
        with torch.no_grad():
            w.mul_(d ** -0.5)
        ops.set_dropout_seed(5)
        if fused:
            y, _ = ops.out_ln(o, tgt, w, b, gm, bt, 1e-5, p)
        else:
            y = ops.layer_norm(ops.dropout(ops.linear(o, w, b), p), tgt, gm, bt, 1e-5)
        (y * go).sum().backward()
        res[fused] = [y.detach()] + [t.grad.clone() for t in (o, tgt, w, b, gm, bt)]
    for a_, b_ in zip(res[True], res[False]):
        assert rel_err(a_, b_) < 1e-5


def test_cls_gather_scatter_equal_take_put_rows():
    """ops.cls_gather / cls_scatter (the frame-CLS exchange with the temporal layer for one un-padded video, two row-sized
    nodes) == take_rows + cat + slices + put_rows (modal_encoder.py:170-195): values and gradients."""
    n, S, d = 5, 7, 16
    g1 = torch.randn(n, S, d, generator=torch.Generator().manual_seed(1))
    g2 = torch.randn(1, d, generator=torch.Generator().manual_seed(2))
    W = torch.randn(d, d, generator=torch.Generator().manual_seed(3))
    res = {}
    for fused in (True, False):
        x = _leaf(n * S, d, seed=4)
        video = _leaf(1, d, seed=5)
        X = (x * 1.5).view(n, S, d)  # a non-leaf stream, like a layer output
        if fused:
            X3, Y, q_op, y_op = ops.cls_gather(X, video, 0)
            assert q_op is None and y_op is None  # operand copies only with ``pos`` in bf16 mode
            Y = torch.tanh(Y @ W)  # stands in for the temporal layer
            X3, vs = ops.cls_scatter(X3, Y, 0)
        else:
            X3, cls = ops.take_rows(X, 0)
            Y = torch.tanh(torch.cat([video, cls], 0) @ W)
            vs = Y[:1]
            X3 = ops.put_rows(X3, Y[1:], 0)
        ((X3 * g1).sum() + (vs * g2).sum()).backward()
        res[fused] = [X3.detach().clone(), vs.detach().clone(), x.grad.clone(), video.grad.clone()]
    for a_, b_ in zip(res[True], res[False]):
        assert rel_err(a_, b_) < 1e-6
    # bf16 mode: the same launches also write the temporal layer's GEMM operands and refresh the stream's operand copy
    ops.set_precision("bf16")
    try:
        x = _leaf(n * S, d, seed=4)
        video = _leaf(1, d, seed=5)
        pos = torch.randn(1 + n, d, generator=torch.Generator().manual_seed(6))
        X = (x * 1.5).view(n, S, d)
        X_op = X.detach().to(torch.bfloat16).clone()
        X3, Y, q_op, y_op = ops.cls_gather(X, video, 0, pos=pos)
        assert torch.equal(Y.detach(), torch.cat([video.detach(), X.detach()[:, 0, :]], 0))
        assert torch.equal(q_op, (Y.detach() + pos).to(torch.bfloat16)) and torch.equal(y_op, Y.detach().to(torch.bfloat16))
        Y2 = torch.tanh(Y @ W)
        X3, vs = ops.cls_scatter(X3, Y2, 0, x_op=X_op)
        assert torch.equal(X_op[:, 0, :], Y2.detach()[1:].to(torch.bfloat16)) and torch.equal(X_op[:, 1:, :], X.detach()[:, 1:, :].to(torch.bfloat16))
        ((X3 * g1).sum() + (vs * g2).sum()).backward()
        assert rel_err(x.grad, res[True][2]) < 1e-6 and rel_err(video.grad, res[True][3]) < 1e-6
    finally:
        ops.set_precision("fp32")


# ---- the layout-glue / anchor-chain autograd nodes (csrc/assembly.cu) against plain torch autograd -------------------------
def _bf(t):
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize("durs", [[6], [3, 2, 4]])
def test_token_assembly_matches_torch_autograd(durs):
    """ops.token_assembly == the reference's cat / transpose / expand assembly (modal_encoder.py:40-72): X, POS, the two operand
    copies, and the gradients of vis / text / frame_cls (the text token is read by every frame of its video)."""
    ops.set_precision("bf16")
    n, b, d, H, W, L = sum(durs), len(durs), 16, 3, 2, 4
    vis, txt, cls = _leaf(n, d, H, W, seed=1), _leaf(L, b, d, seed=2), _leaf(1, d, seed=3)
    vpos, lpos = torch.randn(n, d, H, W, generator=torch.Generator().manual_seed(4)), torch.randn(1, d, generator=torch.Generator().manual_seed(5))
    f2v = torch.tensor([j for j, t in enumerate(durs) for _ in range(t)])
    vid_start = torch.tensor([0] + [sum(durs[: j + 1]) for j in range(b)])
    gX = torch.randn(n, 1 + H * W + L, d, generator=torch.Generator().manual_seed(6))
    Xr = torch.cat([cls.expand(n, 1, d), vis.flatten(2).transpose(1, 2), txt.transpose(0, 1).index_select(0, f2v)], 1)
    Pr = torch.cat([lpos.expand(n, 1, d), vpos.flatten(2).transpose(1, 2), torch.zeros(n, L, d)], 1)
    (Xr * gX).sum().backward()
    want = [t.grad.clone() for t in (vis, txt, cls)]
    for t in (vis, txt, cls):
        t.grad = None
    one = b == 1
    X, POS, qk, xo = ops.token_assembly(vis, vpos, txt, cls, lpos, None if one else f2v, None if one else vid_start)
    assert torch.equal(X, Xr.detach()) and torch.equal(POS, Pr)
    assert torch.equal(qk, (Xr.detach() + Pr).to(torch.bfloat16)) and torch.equal(xo, Xr.detach().to(torch.bfloat16))
    assert not POS.requires_grad and not qk.requires_grad
    (X * gX).sum().backward()
    for t, w in zip((vis, txt, cls), want):
        assert rel_err(t.grad, w) < 1e-6


def test_mem_operands_matches_torch_autograd():
    """ops.mem_operands: operands of rows 1.. and the CLS rows; backward = [g_cls ; g_mem + g_mempos] with bf16 or missing parts."""
    ops.set_precision("bf16")
    n, S, d = 4, 6, 8
    X = _leaf(n, S, d, seed=1)
    POS = torch.randn(n, S, d, generator=torch.Generator().manual_seed(2))
    mem_op, pos_op, mempos_op, cls = ops.mem_operands(X * 1.0, POS)
    assert torch.equal(mem_op, X.detach()[:, 1:].reshape(-1, d).to(torch.bfloat16))
    assert torch.equal(pos_op, POS[:, 1:].reshape(-1, d).to(torch.bfloat16))
    assert torch.equal(mempos_op, (X.detach() + POS)[:, 1:].reshape(-1, d).to(torch.bfloat16))
    assert torch.equal(cls, X.detach()[:, 0])
    g1 = torch.randn(n * (S - 1), d, generator=torch.Generator().manual_seed(3)).to(torch.bfloat16)
    g3 = torch.randn(n, d, generator=torch.Generator().manual_seed(4))
    ((mem_op.float() * g1.float()).sum() + (cls * g3).sum()).backward()  # mempos_op unused: its gradient is None
    want = torch.cat([g3[:, None, :], g1.float().view(n, S - 1, d)], 1)
    assert rel_err(X.grad, want) < 1e-6


@pytest.mark.parametrize("durs", [[5], [2, 3]])
@pytest.mark.parametrize("fused", [False, True])
def test_template_matches_torch_autograd(durs, fused):
    """ops.template == TemplateGenerator.forward + sigmoid (query_decoder.py:441-475, :105) with bf16-rounded GEMM operands: values
    and every gradient (video token, frame-CLS rows, 4 weights, 4 biases), with autograd accumulation and with grad fusion."""
    ops.set_precision("bf16")
    n, b, d, q = sum(durs), len(durs), 16, 4
    v, fc = _leaf(b, d, seed=1), _leaf(n, d, seed=2)
    Ws = [_leaf(d, d, seed=10 + i) for i in range(3)] + [_leaf(q, d, seed=13)]
    bs = [_leaf(d, seed=20 + i) for i in range(3)] + [_leaf(q, seed=23)]
    with torch.no_grad():
        for w in Ws:
            w.mul_(d ** -0.5)
            w.copy_(_bf(w))  # weights exactly representable: the reference below needs no weight rounding
    f2v = torch.tensor([j for j, t in enumerate(durs) for _ in range(t)])
    vid_start = torch.tensor([0] + [sum(durs[: j + 1]) for j in range(b)])
    ga = torch.randn(n, q, generator=torch.Generator().manual_seed(30))
    gt = torch.randn(n, d, generator=torch.Generator().manual_seed(31))

    class R(torch.autograd.Function):  # bf16 rounding of a GEMM operand, straight-through (the kernels round forward and backward operands)
        @staticmethod
        def forward(ctx, x):
            return _bf(x)

        @staticmethod
        def backward(ctx, g):
            return g

    lin = lambda x, W, bb: R.apply(x) @ W.t() + bb
    content = lin(v, Ws[0], bs[0])
    gamma, beta = torch.tanh(lin(v, Ws[1], bs[1])), torch.tanh(lin(v, Ws[2], bs[2]))
    mod = gamma.index_select(0, f2v) * fc + beta.index_select(0, f2v)
    anchor_r = torch.sigmoid(lin(mod, Ws[3], bs[3]))
    temp_r = content.index_select(0, f2v)
    ((anchor_r * ga).sum() + (temp_r * gt).sum()).backward()
    leaves = (v, fc, *Ws, *bs)
    want = [t.grad.clone() for t in leaves]
    for t in leaves:
        t.grad = None
    if fused:
        for t in (*Ws, *bs):
            t.grad = torch.zeros_like(t)
        ops.set_grad_fusion(True)
    one = b == 1
    anchor, temp = ops.template(v, fc, Ws[0], bs[0], Ws[1], bs[1], Ws[2], bs[2], Ws[3], bs[3], None if one else f2v, None if one else vid_start)
    assert rel_err(anchor, anchor_r) < 1e-5 and rel_err(temp, temp_r) < 1e-6
    ((anchor * ga).sum() + (temp * gt).sum()).backward()
    # the kernels round the incoming gradients of the Linears to bf16 as GEMM operands (like every Linear of the package); the
    # straight-through reference does not: 2^-8 relative per operand
    for t, w in zip(leaves, want):
        assert t.grad is not None and rel_err(t.grad, w) < 1.5e-2, rel_err(t.grad, w)


def test_box_mlp_head_and_mul_operand_match_torch_autograd():
    """ops.box_mlp_head (bbox_embed + refinement + sine of the detached result) and ops.mul_operand against torch autograd."""
    import math

    ops.set_precision("bf16")
    R_, d = 7, 16

    class Lyr:
        def __init__(self, o, i, seed):
            self.weight, self.bias = _leaf(o, i, seed=seed), _leaf(o, seed=seed + 50)
            with torch.no_grad():
                self.weight.mul_(i ** -0.5)
                self.weight.copy_(_bf(self.weight))

    layers = [Lyr(d, d, 1), Lyr(d, d, 2), Lyr(4, d, 3)]
    x = _leaf(R_, d, seed=4)
    anchor = torch.rand(R_, 4, generator=torch.Generator().manual_seed(5)).requires_grad_(True)
    g = torch.randn(R_, 4, generator=torch.Generator().manual_seed(6))
    h = _bf(x)
    for lyr in layers[:-1]:
        h = _bf((h @ lyr.weight.t() + lyr.bias).relu())
    delta = h @ layers[-1].weight.t() + layers[-1].bias
    a = anchor.clamp(0, 1)
    ref = torch.sigmoid(delta + torch.log(a.clamp(min=1e-3) / (1 - a).clamp(min=1e-3)))
    out, sine, sine_op = ops.box_mlp_head(layers, x, None, anchor, want_sine=True)
    assert rel_err(out, ref) < 1e-5
    k = torch.arange(128, dtype=torch.float32)
    p = (out.detach() * (2 * math.pi))[..., None] / (10000 ** (2 * torch.div(k, 2, rounding_mode="floor") / 128))
    e = torch.stack((p[..., 0::2].sin(), p[..., 1::2].cos()), dim=-1).flatten(-2)
    assert rel_err(sine, torch.cat((e[..., 1, :], e[..., 0, :], e[..., 2, :], e[..., 3, :]), dim=-1)) < 1e-5
    assert not sine.requires_grad and torch.equal(sine_op, sine.to(torch.bfloat16))
    (out * g).sum().backward()
    got = [t.grad.clone() for t in (x, anchor, *[l.weight for l in layers], *[l.bias for l in layers])]
    # reference gradients: straight-through roundings, plain autograd
    for t in (x, anchor, *[l.weight for l in layers], *[l.bias for l in layers]):
        t.grad = None

    class ST(torch.autograd.Function):
        @staticmethod
        def forward(ctx, t):
            return _bf(t)

        @staticmethod
        def backward(ctx, gg):
            return gg

    h = ST.apply(x)
    for lyr in layers[:-1]:
        h = ST.apply((h @ lyr.weight.t() + lyr.bias).relu())
    delta = h @ layers[-1].weight.t() + layers[-1].bias
    a = anchor.clamp(0, 1)
    (torch.sigmoid(delta + torch.log(a.clamp(min=1e-3) / (1 - a).clamp(min=1e-3))) * g).sum().backward()
    want = [t.grad.clone() for t in (x, anchor, *[l.weight for l in layers], *[l.bias for l in layers])]
    for a_, b_ in zip(got, want):
        assert rel_err(a_, b_) < 1.5e-2, rel_err(a_, b_)
    # mul_operand: product, its operand copy, the operand copy of a second tensor; gradient of the differentiable factor
    s_ = torch.randn(R_, 2 * d, generator=torch.Generator().manual_seed(7))
    sc, qp = _leaf(R_, d, seed=8), torch.randn(R_, d, generator=torch.Generator().manual_seed(9))
    prod, prod_op, qp_op = ops.mul_operand(s_, sc, qp)
    assert torch.equal(prod, s_[:, :d] * sc.detach()) and torch.equal(prod_op, prod.to(torch.bfloat16)) and torch.equal(qp_op, qp.to(torch.bfloat16))
    gp = torch.randn(R_, d, generator=torch.Generator().manual_seed(10))
    (prod * gp).sum().backward()
    assert rel_err(sc.grad, gp * s_[:, :d]) < 1e-6
