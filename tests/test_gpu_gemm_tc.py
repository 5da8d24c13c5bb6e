"""-m gpu: the tcgen05/TMA bf16 GEMM (stcat_linear_* with bf16 operands) against a float64 product of the
same bf16-rounded operands.  Tolerances: fp32 output 2e-5 relative to max|ref| (exact bf16 products,
fp32 accumulation order only); bf16 output 2^-8 (one rounding of the result)."""
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu

TOL_F32 = 2e-5
TOL_BF16 = 4e-3


@pytest.fixture(scope="module")
def be():
    from stcat_b200.cabi import CudaBackend

    return CudaBackend()


def g(*shape, seed=0, scale=1.0):
    gen = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=gen) * scale).to(torch.bfloat16)


SHAPES = [
    (128, 256, 64),       # exactly one tile, one k-block
    (128, 256, 256),
    (256, 512, 128),
    (64, 256, 256),       # M tail (decoder query side)
    (500, 96, 72),        # tails in all three dims
    (1000, 2048, 256),    # FFN linear1
    (1000, 256, 2048),    # FFN linear2 (long K)
    (13632, 256, 256),    # projections at T=64/res=448
    (4264, 768, 256),
    (13632, 2048, 256),   # FFN linear1 at full size
    (64, 256, 2048),      # decoder FFN linear2: 4 tiles x 32 k-blocks -> cluster split-K (8 CTAs per tile)
    (65, 256, 2048),      # temporal encoder layer (T + 1 rows)
    (300, 104, 1024),     # cluster split-K with row / column tails (clusters of 8)
    (13568, 256, 256),    # memory-side projections of the decoder (64 frames x 212 tokens): bf16 dgrad with accumulate
    (64, 2048, 256),      # bwd: dx[64, 256] = dy[64, 2048] . W: the FFN linear1 data gradient
    (129, 2048, 256),
    # few rows, contraction <= 768: the low-latency mma.sync kernel of the dependent chains (csrc/gemm_small.cu)
    (65, 256, 512),       # two 256-wide chunks (ref_point_head layer 1)
    (17, 104, 72),        # tails in every dimension
    (128, 2048, 256),     # 64 CTAs
    (100, 768, 200),
]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_fwd_bf16(be, M, N, K):
    x, w = g(M, K, seed=1), g(N, K, seed=2, scale=K ** -0.5)
    b = torch.randn(N, generator=torch.Generator().manual_seed(3))
    ref = x.double() @ w.double().t() + b.double()
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    y = torch.full((M, N), float("nan"), device="cuda")
    be.linear_fwd(xd, wd, bd, y)
    assert rel_err(y, ref) < TOL_F32
    yb = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    be.linear_fwd(xd, wd, bd, yb, relu=True)
    assert rel_err(yb, ref.relu()) < TOL_BF16
    y2 = torch.ones(M, N, device="cuda")
    be.linear_fwd(xd, wd, None, y2, accumulate=True)
    assert rel_err(y2, ref - b.double() + 1) < TOL_F32


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_bwd_bf16(be, M, N, K):
    x, w, dy = g(M, K, seed=1), g(N, K, seed=2, scale=K ** -0.5), g(M, N, seed=4)
    xd, wd, dyd = x.cuda(), w.cuda(), dy.cuda()
    dx = torch.full((M, K), float("nan"), device="cuda")
    be.linear_bwd_data(dyd, wd, dx)
    ref_dx = dy.double() @ w.double()
    assert rel_err(dx, ref_dx) < TOL_F32
    dxb = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)
    be.linear_bwd_data(dyd, wd, dxb)
    assert rel_err(dxb, ref_dx) < TOL_BF16
    be.linear_bwd_data(dyd, wd, dx, accumulate=True)
    assert rel_err(dx, 2 * ref_dx) < TOL_F32
    # bf16 output with accumulate: the gradient sink of the decoder's memory operands (ops._GradSink)
    be.linear_bwd_data(dyd, wd, dxb, accumulate=True)
    assert rel_err(dxb, 2 * ref_dx) < 2 * TOL_BF16
    dw = torch.full((N, K), float("nan"), device="cuda")
    db = torch.empty(N, device="cuda")
    be.linear_bwd_weight(dyd, xd, dw, db)
    ref_dw = dy.double().t() @ x.double()
    assert rel_err(dw, ref_dw) < TOL_F32
    assert rel_err(db, dy.double().sum(0)) < 1e-4
    be.linear_bwd_weight(dyd, xd, dw, db, accumulate=True)
    assert rel_err(dw, 2 * ref_dw) < TOL_F32


@pytest.mark.parametrize("M,N,K,odt", [(1000, 256, 2048, "bf16"), (13632, 256, 2048, "bf16"), (64, 256, 2048, "bf16"),
                                         (500, 256, 128, "f32"), (300, 64, 192, "f32")])
def test_bwd_data_fused_relu_mask_and_bias_grad(be, M, N, K, odt):
    """dx = (dy W) * (y > 0) and dbias += colsum(dx) in the dgrad epilogue (F.relu backward + bias gradient of the
    layer below, modal_encoder.py:239).  N = out features of the upper Linear, K = width of the ReLU layer."""
    w, dy = g(N, K, seed=2, scale=N ** -0.5), g(M, N, seed=4)
    y = g(M, K, seed=7).relu()  # forward activation of the ReLU layer (zeros where it was clamped)
    ref = (dy.double() @ w.double()) * (y.double() > 0)
    dt = torch.bfloat16 if odt == "bf16" else torch.float32
    dx = torch.full((M, K), float("nan"), device="cuda", dtype=dt)
    db = torch.ones(K, device="cuda")
    be.linear_bwd_data(dy.cuda(), w.cuda(), dx, relu_y=y.cuda(), dbias=db)
    assert rel_err(dx, ref) < (TOL_BF16 if odt == "bf16" else TOL_F32)
    assert bool(((dx != 0) <= (y.cuda() > 0)).all())  # exact zeros where the ReLU was inactive
    # the bias gradient is the column sum of what was STORED (bf16-rounded in bf16 mode), on top of the old content
    assert rel_err(db - 1, dx.double().sum(0)) < 2e-5
    # column sums alone (no mask)
    db2 = torch.zeros(K, device="cuda")
    dx2 = torch.empty(M, K, device="cuda", dtype=dt)
    be.linear_bwd_data(dy.cuda(), w.cuda(), dx2, dbias=db2)
    assert rel_err(dx2, dy.double() @ w.double()) < (TOL_BF16 if odt == "bf16" else TOL_F32)
    assert rel_err(db2, dx2.double().sum(0)) < 2e-5


@pytest.mark.parametrize("dt", ["bf16", "f32"])
@pytest.mark.parametrize("M", [64, 200, 1000])
def test_linear_group(be, M, dt):
    """stcat_linear_group: several multi-term Linear GEMMs in one launch, all three kinds (decoder query side:
    q = Wqc tgt + Wqt time + Wqp pos, k likewise, v; then the packed in-projection; and their backward)."""
    d = 256
    cvt = (lambda t: t) if dt == "bf16" else (lambda t: t.float())
    tol = TOL_F32 if dt == "bf16" else 2e-5
    xs = [cvt(g(M, d, seed=10 + i)).cuda() for i in range(3)]
    ws = [cvt(g(d, d, seed=20 + i, scale=1 / 16)).cuda() for i in range(7)]
    bs = [torch.randn(d, generator=torch.Generator().manual_seed(30 + i)).cuda() for i in range(7)]
    layout = [[(0, 0), (1, 1), (2, 2)], [(0, 3), (1, 4), (2, 5)], [(0, 6)]]  # (input, weight) per term
    odt = torch.bfloat16 if dt == "bf16" else torch.float32
    outs = [torch.full((M, d), float("nan"), device="cuda", dtype=odt) for _ in layout]
    be.linear_group(0, [dict(terms=[(xs[i], ws[w], bs[w]) for i, w in terms], out=o, relu=(j == 2)) for j, (terms, o) in enumerate(zip(layout, outs))])
    refs = []
    for j, terms in enumerate(layout):
        r = sum(xs[i].double() @ ws[w].double().t() + bs[w].double() for i, w in terms)
        refs.append(r.relu() if j == 2 else r)
        assert rel_err(outs[j], refs[j]) < (TOL_BF16 if dt == "bf16" else tol), j
    # packed in-projection rows as weight slices, strided output columns
    wp = cvt(g(3 * d, d, seed=40, scale=1 / 16)).cuda()
    qkv = torch.zeros(M, 3 * d, device="cuda", dtype=odt)
    be.linear_group(0, [dict(terms=[(outs[i], wp[i * d:(i + 1) * d], None)], out=qkv[:, i * d:(i + 1) * d]) for i in range(3)])
    for i in range(3):
        assert rel_err(qkv[:, i * d:(i + 1) * d], outs[i].double() @ wp[i * d:(i + 1) * d].double().t()) < (TOL_BF16 if dt == "bf16" else tol)
    # bwd_data: dx_i = sum over the terms reading x_i
    dys = [cvt(g(M, d, seed=50 + j)).cuda() for j in range(3)]
    dx = [torch.full((M, d), float("nan"), device="cuda") for _ in range(3)]
    jobs = []
    for i in range(3):
        terms = [(dys[j], ws[w], None) for j, tl in enumerate(layout) for ii, w in tl if ii == i]
        jobs.append(dict(terms=terms, out=dx[i]))
    be.linear_group(1, jobs)
    for i in range(3):
        ref = sum(dys[j].double() @ ws[w].double() for j, tl in enumerate(layout) for ii, w in tl if ii == i)
        assert rel_err(dx[i], ref) < tol, i
    # bwd_weight (+ bias column sums), accumulating into existing gradients
    dw = [torch.ones(d, d, device="cuda") for _ in range(7)]
    db = [torch.ones(d, device="cuda") for _ in range(7)]
    jobs = [dict(terms=[(dys[j], xs[i], None)], out=dw[w], accumulate=True, dbias=db[w]) for j, tl in enumerate(layout) for i, w in tl]
    be.linear_group(2, jobs)
    for j, tl in enumerate(layout):
        for i, w in tl:
            assert rel_err(dw[w] - 1, dys[j].double().t() @ xs[i].double()) < 5e-5, w
            assert rel_err(db[w] - 1, dys[j].double().sum(0)) < 1e-4, w


def test_strided_operands_bf16(be):
    """column slices of the packed qkv buffer (ld = 768) and row slices of the packed in_proj weight"""
    M, d = 777, 256
    x = g(M, d, seed=5).cuda()
    w = g(3 * d, d, seed=6, scale=1 / 16).cuda()
    buf = torch.zeros(M, 3 * d, device="cuda", dtype=torch.bfloat16)
    be.linear_fwd(x, w[: 2 * d], None, buf[:, : 2 * d])
    be.linear_fwd(x, w[2 * d:], None, buf[:, 2 * d:])
    ref = x.cpu().double() @ w.cpu().double().t()
    assert rel_err(buf, ref) < TOL_BF16
    dx = torch.empty(M, d, device="cuda")
    be.linear_bwd_data(buf[:, d:2 * d], w[d:2 * d], dx)
    assert rel_err(dx, buf[:, d:2 * d].cpu().double() @ w[d:2 * d].cpu().double()) < TOL_F32
    dw = torch.empty(2 * d, d, device="cuda")
    be.linear_bwd_weight(buf[:, : 2 * d], x, dw, None)
    assert rel_err(dw, buf[:, : 2 * d].cpu().double().t() @ x.cpu().double()) < TOL_F32


def test_tensor_core_kernel_is_what_runs(be):
    """the SIMT kernel must not silently serve a shape the tcgen05 kernel supports: compare timing classes"""
    M, N, K = 13632, 2048, 256
    x, w = g(M, K, seed=1).cuda(), g(N, K, seed=2, scale=1 / 16).cuda()
    y = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        be.linear_fwd(x, w, None, y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        be.linear_fwd(x, w, None, y)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    tflops = 2.0 * M * N * K / (us * 1e-6) / 1e12
    print(f"FFN linear1 fwd bf16: {us:.1f} us/launch, {tflops:.0f} TFLOP/s")
    assert tflops > 150, "tcgen05 path not active (SIMT fp32 peaks far below this)"


@pytest.mark.parametrize("M,N,K", [(13632, 2048, 256), (20000, 1000, 200), (40000, 264, 72), (19000, 512, 256)])
def test_weight_resident_variant(be, M, N, K):
    """K <= 256, bf16 output and >= 2 tiles per SM: the launch takes the weight-resident variant (gemm_tcgen05.cu BRES: the
    [256 x K] weight tile stays in shared memory, CTAs walk contiguous column-block-major tile ranges, half-width 64B-swizzled
    staging tiles).  Forward with bias + ReLU and the data gradient (MN-major weight tile), tails in every dimension."""
    x, w = g(M, K, seed=1), g(N, K, seed=2, scale=K ** -0.5)
    b = torch.randn(N, generator=torch.Generator().manual_seed(3))
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    ref = (xd.double() @ wd.double().t() + bd.double()).relu()
    yb = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    be.linear_fwd(xd, wd, bd, yb, relu=True)
    assert rel_err(yb, ref) < TOL_BF16
    # exact check of the tile bookkeeping: against the fp32-output launch (plain variant) rounded once
    yf = torch.empty(M, N, device="cuda")
    be.linear_fwd(xd, wd, bd, yf, relu=True)
    assert torch.equal(yb, yf.to(torch.bfloat16))
    # data gradient: dx[M, N] = dy[M, K] . W[K, N] with the roles swapped (contraction over the K <= 256 side)
    w2 = g(K, N, seed=5, scale=K ** -0.5).cuda()  # Linear(N -> K): weight [K, N]
    dxb = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    be.linear_bwd_data(xd, w2, dxb)
    dxf = torch.empty(M, N, device="cuda")
    be.linear_bwd_data(xd, w2, dxf)
    assert rel_err(dxf, xd.double() @ w2.double()) < TOL_F32
    assert torch.equal(dxb, dxf.to(torch.bfloat16))


def test_weight_resident_variant_grouped(be):
    """three independent [M, 256] x [256, 256]^T projections with bf16 outputs in one launch (the decoder's memory-side K / V
    projections): each CTA's contiguous range crosses job boundaries, i.e. the resident weight tile is replaced mid-range."""
    M, d = 13568, 256
    xs = [g(M, d, seed=20 + i).cuda() for i in range(2)]
    ws = [g(d, d, seed=30 + i, scale=1 / 16).cuda() for i in range(3)]
    bs = [torch.randn(d, generator=torch.Generator().manual_seed(40 + i)).cuda() for i in range(3)]
    outs = [torch.full((M, d), float("nan"), device="cuda", dtype=torch.bfloat16) for _ in range(3)]
    be.linear_group(0, [dict(terms=[(xs[i % 2], ws[i], bs[i])], out=outs[i]) for i in range(3)])
    for i in range(3):
        ref = xs[i % 2].double() @ ws[i].double().t() + bs[i].double()
        assert rel_err(outs[i], ref) < TOL_BF16, i


def test_cluster_split_k_is_bit_reproducible(be):
    """the cluster split-K sums the partial tiles in rank order: two launches give identical bits"""
    M, N, K = 64, 256, 2048
    x, w = g(M, K, seed=1).cuda(), g(N, K, seed=2, scale=K ** -0.5).cuda()
    y1, y2 = torch.empty(M, N, device="cuda"), torch.empty(M, N, device="cuda")
    be.linear_fwd(x, w, None, y1)
    for _ in range(5):
        be.linear_fwd(x, w, None, y2)
        assert torch.equal(y1, y2)


@pytest.mark.parametrize("M,N,K", [(13632, 2048, 256), (64, 2048, 256), (1000, 512, 256), (300, 264, 72)])
def test_dropout_in_the_epilogue_and_scaled_relu_backward(be, M, N, K):
    """stcat_linear_dropout_fwd: y = drop(relu(x W^T + b)) with the mask drawn in the GEMM epilogue == the plain launch followed
    by stcat_dropout, bit for bit (warp epilogue, weight-resident and skinny-tile variants).  stcat_linear_bwd_data_scaled:
    dx = alpha (dy W) where the DROPPED activation is > 0 == ReLU mask, then dropout, of the unscaled product."""
    p, seed, off = 0.1, 777, (1 << 34) + 5
    x, w = g(M, K, seed=1).cuda(), g(N, K, seed=2, scale=K ** -0.5).cuda()
    b = torch.randn(N, generator=torch.Generator().manual_seed(3)).cuda()
    y = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    be.linear_dropout_fwd(x, w, b, y, True, (p, seed, off))
    y0 = torch.empty(M, N, device="cuda")
    be.linear_fwd(x, w, b, y0, relu=True)          # fp32 result of the same accumulation
    be.dropout(y0, y0, p, seed, off)
    assert torch.equal(y, y0.to(torch.bfloat16))
    assert abs(float((y == 0).float().mean()) - float(((y0 == 0)).float().mean())) == 0
    if N % 64 == 0:
        # backward of the pair: h = y (dropped activation), dz [M, K2] with the second Linear's weight [K2, N]
        K2 = 256
        dz, w2 = g(M, K2, seed=4).cuda(), g(K2, N, seed=5, scale=K2 ** -0.5).cuda()
        alpha = 16777216.0 / (16777216.0 - int(float(torch.tensor(p)) * 16777216.0))
        dh = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
        db = torch.zeros(N, device="cuda")
        be.linear_bwd_data(dz, w2, dh, relu_y=y, dbias=db, alpha=alpha)
        ref = (dz.double() @ w2.double()) * (y.double() > 0) * alpha
        assert rel_err(dh, ref) < TOL_BF16
        assert rel_err(db, dh.double().sum(0)) < 1e-4
