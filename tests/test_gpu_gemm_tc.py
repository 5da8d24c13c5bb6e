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
