"""-m gpu: every C-ABI entry point against a CPU (oracle / torch float64) reference of the same op.

Tolerances are written next to each check.  fp32 kernels: 2e-5 relative (summation order only)."""
import math

import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu

TOL32 = 2e-5


@pytest.fixture(scope="module")
def be():
    from stcat_b200.cabi import CudaBackend

    b = CudaBackend()
    assert b.lib.stcat_device_arch() >= 100, "expected a Blackwell (sm_100) device"
    return b


def g(*shape, seed=0, scale=1.0):
    gen = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=gen) * scale


@pytest.mark.parametrize("M,N,K", [(1, 4, 256), (64, 256, 256), (213 * 3, 96, 70), (500, 2048, 256), (129, 256, 2048)])
def test_linear_fp32(be, M, N, K):
    x, w, b = g(M, K, seed=1), g(N, K, seed=2, scale=K ** -0.5), g(N, seed=3)
    dy = g(M, N, seed=4)
    xd, wd, bd, dyd = x.cuda(), w.cuda(), b.cuda(), dy.cuda()
    ref = x.double() @ w.double().t() + b.double()
    y = torch.empty(M, N, device="cuda")
    be.linear_fwd(xd, wd, bd, y)
    assert rel_err(y, ref) < TOL32
    be.linear_fwd(xd, wd, bd, y, relu=True)
    assert rel_err(y, ref.relu()) < TOL32
    y2 = torch.ones(M, N, device="cuda")
    be.linear_fwd(xd, wd, None, y2, accumulate=True)
    assert rel_err(y2, ref - b.double() + 1) < TOL32
    dx = torch.empty(M, K, device="cuda")
    be.linear_bwd_data(dyd, wd, dx)
    assert rel_err(dx, dy.double() @ w.double()) < TOL32
    dw = torch.empty(N, K, device="cuda")
    db = torch.empty(N, device="cuda")
    be.linear_bwd_weight(dyd, xd, dw, db)
    assert rel_err(dw, dy.double().t() @ x.double()) < TOL32
    assert rel_err(db, dy.double().sum(0)) < TOL32
    be.linear_bwd_weight(dyd, xd, dw, db, accumulate=True)
    assert rel_err(dw, 2 * (dy.double().t() @ x.double())) < TOL32
    assert rel_err(db, 2 * dy.double().sum(0)) < TOL32


def test_linear_strided_views(be):
    """column slices of a wider buffer (packed in_proj outputs) and row slices of a packed weight"""
    M, K = 77, 256
    x = g(M, K, seed=5).cuda()
    w = g(768, K, seed=6, scale=1 / 16).cuda()
    buf = torch.zeros(M, 768, device="cuda")
    be.linear_fwd(x, w[:512], None, buf[:, :512])
    be.linear_fwd(x, w[512:], None, buf[:, 512:])
    assert rel_err(buf, x.cpu().double() @ w.cpu().double().t()) < TOL32
    dx = torch.empty(M, K, device="cuda")
    be.linear_bwd_data(buf[:, 256:512], w[256:512], dx)
    assert rel_err(dx, buf[:, 256:512].cpu().double() @ w[256:512].cpu().double()) < TOL32


@pytest.mark.parametrize("rows", [1, 7, 1000, 13632])
def test_layernorm(be, rows):
    d = 256
    x, r = g(rows, d, seed=1), g(rows, d, seed=2)
    gm, bt = 1 + 0.1 * g(d, seed=3), 0.1 * g(d, seed=4)
    dy = g(rows, d, seed=5)
    xd, rd = x.double().requires_grad_(True), r.double().requires_grad_(True)
    gd, bd = gm.double().requires_grad_(True), bt.double().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xd + rd, (d,), gd, bd, 1e-5)
    ref.backward(dy.double())
    y = torch.empty(rows, d, device="cuda")
    yb = torch.empty(rows, d, device="cuda", dtype=torch.bfloat16)
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    be.layernorm_fwd(x.cuda(), r.cuda(), gm.cuda(), bt.cuda(), y, yb, mean, rstd)
    assert rel_err(y, ref) < TOL32
    assert torch.equal(yb, y.to(torch.bfloat16))
    dz = torch.empty(rows, d, device="cuda")
    dg, dbt = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    be.layernorm_bwd(dy.cuda(), x.cuda(), r.cuda(), gm.cuda(), mean, rstd, dz, dg, dbt)
    assert rel_err(dz, xd.grad) < TOL32
    assert rel_err(dg, gd.grad) < 1e-4  # atomics over many rows
    assert rel_err(dbt, bd.grad) < 1e-4
    # optional extras: bf16 operand copy of dz and the bias gradient of the Linear in front (colsum of dz), accumulated
    dz2 = torch.empty(rows, d, device="cuda")
    dzb = torch.empty(rows, d, device="cuda", dtype=torch.bfloat16)
    dlb = torch.ones(d, device="cuda")
    dg2, dbt2 = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    be.layernorm_bwd(dy.cuda(), x.cuda(), r.cuda(), gm.cuda(), mean, rstd, dz2, dg2, dbt2, dz_bf16=dzb, dbias=dlb)
    assert torch.equal(dz2, dz) and torch.equal(dzb, dz.to(torch.bfloat16))
    assert rel_err(dlb - 1, xd.grad.sum(0)) < 1e-4
    # no residual
    be.layernorm_fwd(x.cuda(), None, gm.cuda(), bt.cuda(), y, None, mean, rstd)
    assert rel_err(y, torch.nn.functional.layer_norm(x.double(), (d,), gm.double(), bt.double(), 1e-5)) < TOL32


@pytest.mark.parametrize("rows", [1, 65, 1000])
def test_layernorm_with_fused_dropout(be, rows):
    """stcat_layernorm_dropout_fwd/_bwd: y = LayerNorm(drop(x) + res) with the counter-based mask (element index row * d + col)
    against float64 with the explicit keep mask; the backward's operand copy and bias column sums carry mask(dz), dz itself
    (the residual branch's gradient) does not."""
    from emu_backend import drop_keep_scale

    d, p, seed, off = 256, 0.1, 4321, (1 << 35) + 17
    x, r = g(rows, d, seed=1), g(rows, d, seed=2)
    gm, bt = 1 + 0.1 * g(d, seed=3), 0.1 * g(d, seed=4)
    dy = g(rows, d, seed=5)
    keep, sc = drop_keep_scale(rows * d, p, seed, off)
    mk = (keep.double() * sc).view(rows, d)
    xd, rd = x.double().requires_grad_(True), r.double().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xd * mk + rd, (d,), gm.double(), bt.double(), 1e-5)
    ref.backward(dy.double())
    y = torch.empty(rows, d, device="cuda")
    yb = torch.empty(rows, d, device="cuda", dtype=torch.bfloat16)
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    drop = (p, seed, off)
    be.layernorm_fwd(x.cuda(), r.cuda(), gm.cuda(), bt.cuda(), y, yb, mean, rstd, drop=drop)
    assert rel_err(y, ref) < TOL32 and torch.equal(yb, y.to(torch.bfloat16))
    dz = torch.empty(rows, d, device="cuda")
    dzb = torch.empty(rows, d, device="cuda", dtype=torch.bfloat16)
    dlb = torch.zeros(d, device="cuda")
    dg, dbt = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    be.layernorm_bwd(dy.cuda(), x.cuda(), r.cuda(), gm.cuda(), mean, rstd, dz, dg, dbt, dz_bf16=dzb, dbias=dlb, drop=drop)
    assert rel_err(dz, rd.grad) < TOL32              # residual branch: unmasked
    assert rel_err(dzb, xd.grad) < 4e-3              # gradient w.r.t. x: masked, one bf16 rounding
    assert torch.equal(dzb == 0, (dz * mk.float().cuda()).to(torch.bfloat16) == 0)
    assert rel_err(dlb, xd.grad.sum(0)) < 1e-4


def attn_ref(q1, q2, k1, k2, v, mask, B, H, Lq, Lk, scale):
    """float64 batch-major reference with the same contract as stcat_attention_fwd"""
    def heads(t, L):
        return t.view(B, L, H, 32).permute(0, 2, 1, 3)

    s = heads(q1, Lq) @ heads(k1, Lk).transpose(-1, -2)
    if q2 is not None:
        s = s + heads(q2, Lq) @ heads(k2, Lk).transpose(-1, -2)
    s = s * scale
    if mask is not None:
        s = s.masked_fill(mask.bool()[:, None, None, :], float("-inf"))
    p = torch.softmax(s, -1)
    o = (p @ heads(v, Lk)).permute(0, 2, 1, 3).reshape(B * Lq, H * 32)
    return o, p.mean(1), torch.logsumexp(s, -1)


@pytest.mark.parametrize("B,H,Lq,Lk,two,use_mask,use_pavg", [
    (3, 8, 50, 77, False, True, False),
    (2, 8, 213, 213, False, False, False),
    (5, 8, 1, 212, True, True, False),
    (2, 8, 65, 65, False, True, True),
    (2, 4, 17, 130, True, True, True),
    (1, 8, 64, 64, False, False, False),    # decoder self-attention over T = 64 queries (short-sequence kernel)
    (1, 8, 128, 128, False, True, True),    # short-sequence kernel at its size limit, with the weights output
    (2, 8, 5, 3, False, True, True),
    (2, 8, 150, 140, False, True, True),    # generic SIMT path (too long for the short-sequence kernel)
])
def test_attention_fp32(be, B, H, Lq, Lk, two, use_mask, use_pavg):
    E = H * 32
    scale = (64 if two else 32) ** -0.5
    t = lambda L, s: g(B * L, E, seed=s)
    q1, k1, v = t(Lq, 1), t(Lk, 2), t(Lk, 3)
    q2, k2 = (t(Lq, 4), t(Lk, 5)) if two else (None, None)
    mask = None
    if use_mask:
        mask = torch.zeros(B, Lk, dtype=torch.uint8)
        for b in range(B):
            mask[b, Lk - 1 - 3 * b:] = 1
            mask[b, 0] = 0
    d_o = t(Lq, 6)
    dpavg = g(B, Lq, Lk, seed=7) if use_pavg else None
    leaves = [x.double().requires_grad_(True) if x is not None else None for x in (q1, q2, k1, k2, v)]
    o_ref, pavg_ref, lse_ref = attn_ref(*leaves, mask, B, H, Lq, Lk, scale)
    loss = (o_ref * d_o.double()).sum()
    if use_pavg:
        loss = loss + (pavg_ref * dpavg.double()).sum()
    loss.backward()

    c = lambda x: None if x is None else x.cuda()
    o = torch.empty(B * Lq, E, device="cuda")
    lse = torch.empty(B, H, Lq, device="cuda")
    pavg = torch.zeros(B, Lq, Lk, device="cuda") if use_pavg else None
    be.attention_fwd(c(q1), c(q2), c(k1), c(k2), c(v), o, c(mask), lse, pavg, B, H, Lq, Lk, scale)
    assert rel_err(o, o_ref) < TOL32
    assert rel_err(lse, lse_ref) < TOL32
    if use_pavg:
        assert rel_err(pavg, pavg_ref) < TOL32
    delta = torch.empty(B, H, Lq, device="cuda")
    e = lambda L: torch.empty(B * L, E, device="cuda")
    dq1, dk1, dv = e(Lq), e(Lk), e(Lk)
    dq2, dk2 = (e(Lq), e(Lk)) if two else (None, None)
    be.attention_bwd(c(q1), c(q2), c(k1), c(k2), c(v), c(d_o), c(mask), lse, c(dpavg), delta, dq1, dq2, dk1, dk2, dv,
                     B, H, Lq, Lk, scale)
    tol = 5e-5
    assert rel_err(dq1, leaves[0].grad) < tol
    assert rel_err(dk1, leaves[2].grad) < tol
    assert rel_err(dv, leaves[4].grad) < tol
    if two:
        assert rel_err(dq2, leaves[1].grad) < tol
        assert rel_err(dk2, leaves[3].grad) < tol


@pytest.mark.parametrize("B,H,Lq,Lk,use_mask,use_pavg", [
    (1, 8, 65, 65, False, True), (2, 8, 64, 64, True, False), (2, 4, 50, 77, True, True), (1, 8, 128, 128, True, True),
    (2, 8, 5, 3, True, True), (1, 8, 17, 100, False, False),
])
def test_short_sequence_attention_bf16(be, B, H, Lq, Lk, use_mask, use_pavg):
    """bf16 operands through the short-sequence tensor-core kernels (mma.sync; P / dS rounded to bf16 for the second
    products like every bf16 attention kernel; softmax statistics, weights output and delta in fp32)"""
    E = H * 32
    tb = lambda L, s: g(B * L, E, seed=s).to(torch.bfloat16)
    q, k, v, d_o = tb(Lq, 1), tb(Lk, 2), tb(Lk, 3), tb(Lq, 6)
    mask = None
    if use_mask:
        mask = torch.zeros(B, Lk, dtype=torch.uint8)
        for b in range(B):
            mask[b, Lk - 1 - b:] = 1
            mask[b, 0] = 0
    dpavg = g(B, Lq, Lk, seed=7) if use_pavg else None
    leaves = [x.double().requires_grad_(True) for x in (q, k, v)]
    o_ref, pavg_ref, lse_ref = attn_ref(leaves[0], None, leaves[1], None, leaves[2], mask, B, H, Lq, Lk, 32 ** -0.5)
    loss = (o_ref * d_o.double()).sum()
    if use_pavg:
        loss = loss + (pavg_ref * dpavg.double()).sum()
    loss.backward()
    c = lambda x: None if x is None else x.cuda()
    o = torch.empty(B * Lq, E, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, H, Lq, device="cuda")
    pavg = torch.zeros(B, Lq, Lk, device="cuda") if use_pavg else None
    be.attention_fwd(q.cuda(), None, k.cuda(), None, v.cuda(), o, c(mask), lse, pavg, B, H, Lq, Lk, 32 ** -0.5)
    assert rel_err(o, o_ref) < 6e-3 and rel_err(lse, lse_ref) < TOL32
    if use_pavg:
        assert rel_err(pavg, pavg_ref) < TOL32
    e = lambda L: torch.empty(B * L, E, device="cuda", dtype=torch.bfloat16)
    dq, dk, dv = e(Lq), e(Lk), e(Lk)
    be.attention_bwd(q.cuda(), None, k.cuda(), None, v.cuda(), d_o.cuda(), c(mask), lse, c(dpavg), torch.empty(B, H, Lq, device="cuda"),
                     dq, None, dk, None, dv, B, H, Lq, Lk, 32 ** -0.5)
    for got, leaf in ((dq, leaves[0]), (dk, leaves[1]), (dv, leaves[2])):
        assert rel_err(got, leaf.grad) < 1e-2


def test_anchor_glue_kernels(be):
    """anchor sine embedding and box refinement (one kernel each way) vs the torch restatement of net_utils.py:29-63"""
    from stcat_b200.decoder import anchor_sine_embed, inverse_sigmoid, anchor_sine_embed_op, box_refine
    from stcat_b200 import ops

    gen = torch.Generator().manual_seed(11)
    a = torch.rand(77, 4, generator=gen).clamp(1e-4, 1 - 1e-4)
    a[0] = torch.tensor([0.0, 1.0, 5e-4, 1 - 5e-4])  # the clamps of the logit
    dy = torch.randn(77, 512, generator=gen)
    ar = a.clone().double().requires_grad_(True)
    ref = anchor_sine_embed(ar)
    ref.backward(dy.double())
    for prec in ("fp32", "bf16"):
        ops.set_precision(prec)
        ac = a.clone().cuda().requires_grad_(True)
        out, out_op = anchor_sine_embed_op(ac)
        out.backward(dy.cuda())
        assert rel_err(out, ref) < 2e-6 and rel_err(ac.grad, ar.grad) < 2e-5
        assert (out_op is None) == (prec == "fp32")
        if out_op is not None:
            assert torch.equal(out_op, out.to(torch.bfloat16))
    ops.set_precision("fp32")
    d = torch.randn(77, 4, generator=gen)
    g = torch.randn(77, 4, generator=gen)
    dr, ar2 = d.clone().double().requires_grad_(True), a.clone().double().requires_grad_(True)
    r = torch.sigmoid(dr + inverse_sigmoid(ar2))
    r.backward(g.double())
    dc, ac = d.clone().cuda().requires_grad_(True), a.clone().cuda().requires_grad_(True)
    o = box_refine(dc, ac)
    o.backward(g.cuda())
    assert rel_err(o, r) < 2e-6 and rel_err(dc.grad, dr.grad) < 2e-5
    interior = (a > 1e-3) & (a < 1 - 1e-3)
    assert rel_err(ac.grad.cpu()[interior], ar2.grad[interior]) < 2e-5


def test_elementwise(be):
    a, b = g(1001, 256, seed=1), g(1001, 256, seed=2)
    out = torch.empty(1001, 256, device="cuda")
    ob = torch.empty(1001, 256, device="cuda", dtype=torch.bfloat16)
    be.add(a.cuda(), b.cuda(), out, ob)
    assert torch.equal(out.cpu(), a + b)
    assert torch.equal(ob.cpu(), (a + b).to(torch.bfloat16))
    y = g(333, 2048, seed=3).relu()
    dy = g(333, 2048, seed=4)
    dyd = dy.cuda()
    be.relu_bwd(y.cuda(), dyd)
    assert torch.equal(dyd.cpu(), dy * (y > 0))
    x = g(300, 70, seed=5)
    o1 = torch.empty(300, 70, device="cuda", dtype=torch.bfloat16)
    o2 = torch.empty(70, 300, device="cuda", dtype=torch.bfloat16)
    be.cast_bf16(x.cuda(), o1)
    be.cast_bf16(x.cuda(), o2, transpose=True)
    assert torch.equal(o1.cpu(), x.to(torch.bfloat16))
    assert torch.equal(o2.cpu(), x.t().to(torch.bfloat16))


@pytest.mark.parametrize("b,t,durs", [(1, 64, [64]), (3, 20, [20, 7, 1]), (2, 300, [300, 123])])
def test_sted_score(be, b, t, durs):
    from oracle import stcat_oracle as O

    sted = g(b, t, 2, seed=9, scale=3.0)
    boxes = torch.rand(b * t, 4)
    sizes = torch.ones(b * t, 2)
    frames = [list(range(t)) for _ in range(b)]
    _, steds, score_ref = O.post_process(sted, boxes, sizes, frames, durs)
    score = torch.empty(b, t, t, device="cuda")
    best = torch.empty(b, dtype=torch.int32, device="cuda")
    be.sted_score(sted.cuda(), torch.tensor(durs, dtype=torch.int32, device="cuda"), score, best)
    valid = score_ref > -1e31
    assert torch.allclose(score.cpu()[valid], score_ref[valid], rtol=1e-5, atol=1e-5)
    assert bool((score.cpu()[~valid] < -1e31).all())
    for i in range(b):
        idx = int(best[i])
        if durs[i] > 1:  # with a single valid frame every cell is masked; the reference argmax is then arbitrary (0)
            assert [idx // t, idx % t + 1] == steds[i]


def test_map2d_pool(be):
    from oracle import stcat_oracle as O

    N, counts, d, B = 32, [7, 4, 4], 256, 3
    x = g(B, N, d, seed=11)
    ref = O.gen_2d_map(x, N, counts)
    mask2d, _, _ = O.map2d_masks(N, counts)
    out = torch.empty(B, d, N, N, device="cuda")
    be.map2d_pool(x.cuda(), mask2d.to(torch.uint8).cuda(), out)
    assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_dropout_mask_is_the_counter_based_stream(be, dtype):
    """stcat_dropout keeps element i iff drop_bits24(seed, offset + i) (csrc/common.cuh) reaches p * 2^24: bit-for-bit the
    mask tests/emu_backend.drop_keep_scale builds on the host (which is checked against a pure-python restatement on the CPU)."""
    from emu_backend import drop_keep_scale

    n, p, seed, off = 100003, 0.1, 0x1234567, 987654321
    keep, sc = drop_keep_scale(n, p, seed, off)
    x = g(n, seed=1).to(dtype)
    out = torch.empty(n, device="cuda", dtype=dtype)
    be.dropout(x.cuda(), out, p, seed, off)
    want = (x.float() * keep.float() * sc).to(dtype)
    assert torch.equal(out.cpu(), want)
    assert abs(float(keep.float().mean()) - (1 - p)) < 5e-3
    xin = x.cuda()
    be.dropout(xin, xin, p, seed, off)  # in place
    assert torch.equal(xin.cpu(), want)
    be.dropout(x.cuda(), out, 0.0, seed, off)  # p = 0: identity
    assert torch.equal(out.cpu(), x)


@pytest.mark.parametrize("B,H,Lq,Lk,two,use_mask,use_pavg,bf16", [
    (2, 8, 50, 77, False, True, True, False),
    (2, 8, 213, 213, False, True, False, False),   # the encoder's spatial attention shape
    (3, 8, 1, 212, True, True, False, False),      # time-aligned single-query cross-attention
    (1, 8, 64, 64, False, False, True, False),     # decoder self-attention / time decoder with the weights output
    (2, 4, 17, 130, True, True, True, False),
    (2, 8, 65, 65, False, True, True, True),
    (1, 8, 213, 213, False, False, False, True),
])
def test_attention_with_dropout(be, B, H, Lq, Lk, two, use_mask, use_pavg, bf16):
    """stcat_attention_dropout_fwd/_bwd against float64 attention with the explicit keep mask multiplied into the
    probabilities (torch's multi_head_attention_forward: dropout after the softmax, the returned weights are the dropped ones)."""
    from emu_backend import drop_keep_scale

    p, seed, off = 0.1, 77, 1 << 33
    E = H * 32
    dt = torch.bfloat16 if bf16 else torch.float32
    scale = (64 if two else 32) ** -0.5
    t = lambda L, s: g(B * L, E, seed=s).to(dt)
    q1, k1, v = t(Lq, 1), t(Lk, 2), t(Lk, 3)
    q2, k2 = (t(Lq, 4), t(Lk, 5)) if two else (None, None)
    mask = None
    if use_mask:
        mask = torch.zeros(B, Lk, dtype=torch.uint8)
        for b in range(B):
            mask[b, Lk - 1 - 3 * b:] = 1
            mask[b, 0] = 0
    d_o = t(Lq, 6)
    dpavg = g(B, Lq, Lk, seed=7) if use_pavg else None
    keep, sc = drop_keep_scale(B * H * Lq * Lk, p, seed, off)
    km = (keep.double() * sc).view(B, H, Lq, Lk)
    leaves = [x.double().requires_grad_(True) if x is not None else None for x in (q1, q2, k1, k2, v)]
    hd = lambda x, L: x.view(B, L, H, 32).permute(0, 2, 1, 3)
    s = hd(leaves[0], Lq) @ hd(leaves[2], Lk).transpose(-1, -2)
    if two:
        s = s + hd(leaves[1], Lq) @ hd(leaves[3], Lk).transpose(-1, -2)
    s = s * scale
    if mask is not None:
        s = s.masked_fill(mask.bool()[:, None, None, :], float("-inf"))
    pr = torch.softmax(s, -1) * km
    o_ref = (pr @ hd(leaves[4], Lk)).permute(0, 2, 1, 3).reshape(B * Lq, E)
    loss = (o_ref * d_o.double()).sum()
    if use_pavg:
        loss = loss + (pr.mean(1) * dpavg.double()).sum()
    loss.backward()

    c = lambda x: None if x is None else x.cuda()
    o = torch.empty(B * Lq, E, device="cuda", dtype=dt)
    lse = torch.empty(B, H, Lq, device="cuda")
    pavg = torch.zeros(B, Lq, Lk, device="cuda") if use_pavg else None
    drop = (p, seed, off)
    be.attention_fwd(c(q1), c(q2), c(k1), c(k2), c(v), o, c(mask), lse, pavg, B, H, Lq, Lk, scale, drop=drop)
    tol_o, tol_g = (6e-3, 1e-2) if bf16 else (TOL32, 5e-5)
    assert rel_err(o, o_ref) < tol_o
    assert rel_err(lse, torch.logsumexp(s, -1)) < TOL32  # the statistics are those of the undropped softmax
    if use_pavg:
        assert rel_err(pavg, pr.mean(1)) < TOL32
    e = lambda L: torch.empty(B * L, E, device="cuda", dtype=dt)
    dq1, dk1, dv = e(Lq), e(Lk), e(Lk)
    dq2, dk2 = (e(Lq), e(Lk)) if two else (None, None)
    be.attention_bwd(c(q1), c(q2), c(k1), c(k2), c(v), c(d_o), c(mask), lse, c(dpavg), torch.empty(B, H, Lq, device="cuda"),
                     dq1, dq2, dk1, dk2, dv, B, H, Lq, Lk, scale, drop=drop)
    assert rel_err(dq1, leaves[0].grad) < tol_g
    assert rel_err(dk1, leaves[2].grad) < tol_g
    assert rel_err(dv, leaves[4].grad) < tol_g
    if two:
        assert rel_err(dq2, leaves[1].grad) < tol_g
        assert rel_err(dk2, leaves[3].grad) < tol_g


# ---- layout glue (csrc/assembly.cu): the CUDA kernels against the plain-torch restatement of tests/emu_backend.py ---------
def _emu():
    from emu_backend import EmuBackend

    return EmuBackend()


@pytest.mark.parametrize("n,H,W,L,durs", [(5, 7, 7, 8, None), (8, 14, 14, 16, None), (8, 5, 9, 3, [5, 3]), (4, 14, 23, 16, [1, 3])])
def test_token_assembly_and_backward(be, n, H, W, L, durs):
    d = 256
    b = 1 if durs is None else len(durs)
    vis, vpos, text = g(n, d, H, W, seed=1), g(n, d, H, W, seed=2), g(L, b, d, seed=3)
    cls, lpos = g(1, d, seed=4), g(1, d, seed=5)
    f2v = vid_start = None
    if durs is not None:
        f2v = torch.tensor([j for j, t in enumerate(durs) for _ in range(t)], dtype=torch.long)
        vid_start = torch.tensor([0] + [sum(durs[: j + 1]) for j in range(b)], dtype=torch.long)
    S = 1 + H * W + L
    mk = lambda dt: torch.empty(n, S, d, dtype=dt)
    X, POS, qk, xo = mk(torch.float32), mk(torch.float32), mk(torch.bfloat16), mk(torch.bfloat16)
    _emu().token_assembly(vis, vpos, text, f2v, cls, lpos, X, POS, qk, xo)
    c = lambda t: None if t is None else t.cuda()
    Xd, POSd, qkd, xod = (torch.empty_like(t, device="cuda") for t in (X, POS, qk, xo))
    be.token_assembly(c(vis), c(vpos), c(text), c(f2v), c(cls), c(lpos), Xd, POSd, qkd, xod)
    assert torch.equal(Xd.cpu(), X) and torch.equal(POSd.cpu(), POS)  # pure data movement: bit-exact
    assert torch.equal(qkd.cpu(), qk) and torch.equal(xod.cpu(), xo)  # one fp32 add + round-to-nearest-even: bit-exact
    dX = g(n, S, d, seed=6)
    dvis, dtext, dcls = torch.empty(n, d, H, W), torch.empty(L, b, d), torch.empty(1, d)
    _emu().token_assembly_bwd(dX, dvis, dtext, dcls, vid_start, H * W, L, b)
    dvd, dtd, dcd = (torch.full_like(t, float("nan"), device="cuda") for t in (dvis, dtext, dcls))
    be.token_assembly_bwd(c(dX), dvd, dtd, dcd, c(vid_start), H * W, L, b)
    assert torch.equal(dvd.cpu(), dvis)
    assert rel_err(dtd, dtext) < TOL32 and rel_err(dcd, dcls) < TOL32  # sums over frames: summation order only
    be.token_assembly_bwd(c(dX), None, dtd, None, c(vid_start), H * W, L, b)  # every output is optional
    assert rel_err(dtd, dtext) < TOL32


@pytest.mark.parametrize("n,S", [(3, 9), (64, 213), (5, 340)])
@pytest.mark.parametrize("gdt", [torch.bfloat16, torch.float32])
def test_mem_operands_and_backward(be, n, S, gdt):
    d = 256
    X, POS = g(n, S, d, seed=1), g(n, S, d, seed=2)
    M = S - 1
    ref = [torch.empty(n * M, d, dtype=torch.bfloat16) for _ in range(3)] + [torch.empty(n, d)]
    _emu().mem_operands(X, POS, *ref)
    got = [torch.empty_like(t, device="cuda") for t in ref]
    be.mem_operands(X.cuda(), POS.cuda(), *got)
    for a, r in zip(got, ref):
        assert torch.equal(a.cpu(), r)
    g1, g2, gc = g(n * M, d, seed=3).to(gdt), g(n * M, d, seed=4).to(gdt), g(n, d, seed=5)
    for use in ((True, True, True), (True, False, True), (False, True, False), (True, True, False)):
        a1, a2, ac = (t if u else None for t, u in zip((g1, g2, gc), use))
        dX = torch.empty(n, S, d)
        _emu().mem_operands_bwd(a1, a2, ac, dX)
        dXd = torch.full((n, S, d), float("nan"), device="cuda")
        cu = lambda t: None if t is None else t.cuda()
        be.mem_operands_bwd(cu(a1), cu(a2), cu(ac), dXd)
        assert torch.equal(dXd.cpu(), dX), use


@pytest.mark.parametrize("durs", [[64], [5, 3], [1, 7, 2]])
def test_template_generator_kernels(be, durs):
    d, q = 256, 4
    b, n = len(durs), sum(durs)
    bf = torch.bfloat16
    v, fc = g(b, d, seed=1), g(n, d, seed=2)
    W = [g(d, d, seed=10 + i, scale=d ** -0.5).to(bf) for i in range(3)] + [g(q, d, seed=13, scale=d ** -0.5).to(bf)]
    bias = [g(d, seed=20 + i, scale=0.1) for i in range(3)] + [g(q, seed=23, scale=0.1)]
    f2v = vid_start = None
    if b > 1:
        f2v = torch.tensor([j for j, t in enumerate(durs) for _ in range(t)], dtype=torch.long)
        vid_start = torch.tensor([0] + [sum(durs[: j + 1]) for j in range(b)], dtype=torch.long)

    def run(bk, dev):
        mv = lambda t: None if t is None else t.to(dev)
        e = lambda *s, dt=torch.float32: torch.empty(*s, dtype=dt, device=dev)
        content, gamma, beta, mod, anchor, temp = e(b, d), e(b, d), e(b, d), e(n, d, dt=bf), e(n, q), e(n, d)
        Wd, bd = [mv(w) for w in W], [mv(x) for x in bias]
        bk.template_fwd(mv(v), mv(fc), mv(f2v), Wd[0], bd[0], Wd[1], bd[1], Wd[2], bd[2], Wd[3], bd[3], content, gamma, beta, mod, anchor, temp)
        ga, gt = mv(g(n, q, seed=30)), mv(g(n, d, seed=31))
        dpq, dmod, dpre, dfc, dv = e(n, q, dt=bf), e(n, d), e(3, b, d), e(n, d), e(b, d)
        dW = [torch.ones(d, d, device=dev) for _ in range(3)] + [torch.ones(q, d, device=dev)]  # accumulated: start from 1
        db = [torch.ones(d, device=dev) for _ in range(3)] + [torch.ones(q, device=dev)]
        bk.template_bwd(ga, gt, anchor, mv(v), mv(fc), mv(f2v), mv(vid_start), gamma, beta, mod, Wd[0], Wd[1], Wd[2], Wd[3], dpq, dmod,
                        dpre, dfc, dv, dW[0], db[0], dW[1], db[1], dW[2], db[2], dW[3], db[3])
        return dict(content=content, gamma=gamma, beta=beta, mod=mod.float(), anchor=anchor, temp=temp, dfc=dfc, dv=dv,
                    **{f"dW{i}": t for i, t in enumerate(dW)}, **{f"db{i}": t for i, t in enumerate(db)})

    ref, got = run(_emu(), "cpu"), run(be, "cuda")
    for k in ref:
        # fp32 accumulation of bf16 products in another order; bf16-rounded intermediates (mod, dpq) may flip one ulp (2^-8)
        tol = 1e-2 if k in ("mod", "dfc", "dv") or k.startswith(("dW", "db")) else 1e-4
        assert rel_err(got[k], ref[k]) < tol, (k, rel_err(got[k], ref[k]))


@pytest.mark.parametrize("R", [1, 64, 201])
def test_box_head_and_mul_cast(be, R):
    """stcat_box_head_fwd (last bbox_embed Linear + anchor refinement + sine embedding) and stcat_mul_cast(_bwd) against the
    plain-torch restatement; the sine embedding of an anchor that differs by fp32 summation order in the box head is compared
    through the anchor (2 pi a / 10000^0 ... amplifies an anchor difference by at most 2 pi)."""
    K = 256
    bf = torch.bfloat16
    h = g(R, 2 * K, seed=1).to(bf)[:, :K]  # a row-major view with a leading dimension
    W, bias, anchor = g(4, K, seed=2, scale=K ** -0.5).to(bf), g(4, seed=3, scale=0.1), torch.rand(R, 4, generator=torch.Generator().manual_seed(4))
    anchor[0, 0], anchor[-1, 3] = 0.0, 1.0  # the clamped ends of inverse_sigmoid
    ref = [torch.empty(R, 4), torch.empty(R, 512), torch.empty(R, 512, dtype=bf)]
    _emu().box_head_fwd(h, W, bias, anchor, *ref)
    got = [torch.empty_like(t, device="cuda") for t in ref]
    be.box_head_fwd(h.cuda(), W.cuda(), bias.cuda(), anchor.cuda(), *got)
    assert rel_err(got[0], ref[0]) < 1e-5
    assert float((got[1].cpu() - ref[1]).abs().max()) < 1e-4
    assert float((got[2].float().cpu() - ref[2].float()).abs().max()) < 1e-2  # one bf16 ulp at |v| <= 1
    out_only = torch.empty(R, 4, device="cuda")
    be.box_head_fwd(h.cuda(), W.cuda(), bias.cuda(), anchor.cuda(), out_only, None, None)
    assert torch.equal(out_only, got[0])
    # backward: refinement gradient, bf16 operand copy, data gradient of the last Linear with the ReLU mask of h
    hr = h.float().relu().to(bf).contiguous()
    gg = g(R, 4, seed=9)
    rdd, rdh, rda = torch.empty(R, 4, dtype=bf), torch.empty(R, K, dtype=bf), torch.empty(R, 4)
    _emu().box_head_bwd(gg, ref[0], anchor, W, hr, rdd, rdh, rda)
    gdd, gdh, gda = (torch.empty_like(t, device="cuda") for t in (rdd, rdh, rda))
    be.box_head_bwd(gg.cuda(), ref[0].cuda(), anchor.cuda(), W.cuda(), hr.cuda(), gdd, gdh, gda)
    assert rel_err(gdd.float(), rdd.float()) < 1e-2 and rel_err(gda, rda) < 1e-5
    assert rel_err(gdh.float(), rdh.float()) < 1e-2  # bf16 outputs: a one-ulp flip of dd_op moves dh by 2^-8 relative
    assert torch.equal(gdh.cpu() == 0, rdh == 0) or float(((gdh.cpu() == 0) != (rdh == 0)).float().mean()) < 1e-3
    be.box_head_bwd(gg.cuda(), ref[0].cuda(), anchor.cuda(), W.cuda(), hr.cuda(), gdd, gdh, None)
    a, b = g(R, 512, seed=5), g(R, 256, seed=6)
    rf, rb = torch.empty(R, 256), torch.empty(R, 256, dtype=bf)
    _emu().mul_cast(a, b, rf, rb)
    gf, gb = torch.empty(R, 256, device="cuda"), torch.empty(R, 256, device="cuda", dtype=bf)
    be.mul_cast(a.cuda(), b.cuda(), gf, gb)
    assert torch.equal(gf.cpu(), rf) and torch.equal(gb.cpu(), rb)
    c_in, c_out = g(R, 256, seed=8), torch.empty(R, 256, device="cuda", dtype=bf)
    be.mul_cast(a.cuda(), b.cuda(), None, gb, c_in.cuda(), c_out)  # fp32 product optional; second operand copy in the same launch
    assert torch.equal(gb.cpu(), rb) and torch.equal(c_out.cpu(), c_in.to(bf))
    for gdt in (torch.float32, bf):
        gr = g(R, 256, seed=7).to(gdt)
        rd, gd = torch.empty(R, 256), torch.empty(R, 256, device="cuda")
        _emu().mul_cast_bwd(gr, a, rd)
        be.mul_cast_bwd(gr.cuda(), a.cuda(), gd)
        assert torch.equal(gd.cpu(), rd)


@pytest.mark.parametrize("n,S", [(5, 7), (64, 213)])
def test_cls_gather_scatter_kernels(be, n, S):
    d = 256
    bf = torch.bfloat16
    X, video, pos = g(n, S, d, seed=1), g(1, d, seed=2), g(1 + n, d, seed=3)
    ref = [torch.empty(1 + n, d), torch.empty(1 + n, d, dtype=bf), torch.empty(1 + n, d, dtype=bf)]
    _emu().cls_gather(X, video, pos, *ref, 0)
    got = [torch.empty_like(t, device="cuda") for t in ref]
    be.cls_gather(X.cuda(), video.cuda(), pos.cuda(), *got, 0)
    for a, r in zip(got, ref):
        assert torch.equal(a.cpu(), r)
    Y, P = g(1 + n, d, seed=4), g(n, S, d, seed=5)
    Xr, Xo, Qn = X.clone(), X.to(bf), g(n, S, d, seed=6).to(bf)
    Xd, Xod, Qnd = Xr.cuda(), Xo.cuda(), Qn.cuda()
    _emu().cls_scatter(Y, Xr, Xo, 0, Qn, P)
    be.cls_scatter(Y.cuda(), Xd, Xod, 0, Qnd, P.cuda())
    assert torch.equal(Xd.cpu(), Xr) and torch.equal(Xod.cpu(), Xo) and torch.equal(Qnd.cpu(), Qn)
    be.cls_scatter(Y.cuda(), Xd, None, 0)  # operand copies are optional
    assert torch.equal(Xd.cpu(), Xr)
