"""-m gpu: the fused optimizer-side step (csrc/optim.cu: stcat_sumsq, stcat_adamw_step) and optim.FusedAdamW on the B200
against their fp32 restatement / torch.optim.AdamW + clip_grad_norm_ + the reference's EMA (first run on hardware: round 2,
profiles/r2_a_validate_staged.log)."""
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def be():
    from stcat_b200.cabi import CudaBackend

    return CudaBackend()


@pytest.mark.parametrize("n,off", [(1 << 20, 0), (100003, 0), (4099, 1), (7, 3)])
def test_adamw_step_and_sumsq(be, n, off):
    """stcat_sumsq / stcat_adamw_step vs the fp32 restatement in tests/emu_backend.py (which is checked against
    torch.optim.AdamW + clip_grad_norm_ + the reference's EMA on CPU); aligned and misaligned ranges, vector tail."""
    from emu_backend import EmuBackend

    g = torch.Generator().manual_seed(n)
    mk = lambda s=1.0: torch.randn(n + off, generator=g) * s
    p, gr, m, v, ema = mk(), mk(0.1), mk(0.01), mk(0.01).abs(), mk()
    sl = lambda t: t[off:]
    dev = [t.cuda() for t in (p, gr, m, v, ema)]
    shadow = torch.zeros(n + off, dtype=torch.bfloat16, device="cuda")
    acc = torch.full((1,), 2.5, device="cuda")
    be.sumsq(sl(dev[1]), acc)
    ref_acc = torch.full((1,), 2.5)
    EmuBackend().sumsq(sl(gr).contiguous(), ref_acc)
    assert rel_err(acc, ref_acc) < 1e-5
    args = dict(lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-2, step=3, max_norm=0.1, ema_decay=0.999)
    be.adamw_step(sl(dev[0]), sl(dev[1]), sl(dev[2]), sl(dev[3]), sl(dev[4]), sl(shadow), total_sumsq=acc, **args)
    rp, rg, rm, rv, re = [sl(t).clone() for t in (p, gr, m, v, ema)]
    rs = torch.zeros(n, dtype=torch.bfloat16)
    EmuBackend().adamw_step(rp, rg, rm, rv, re, rs, total_sumsq=ref_acc, **args)
    for got, want in zip(dev[:1] + dev[2:], (rp, rm, rv, re)):
        assert rel_err(sl(got), want) < 2e-6
    assert torch.equal(sl(shadow).cpu(), sl(dev[0]).cpu().to(torch.bfloat16))
    if off:
        assert torch.equal(dev[0][:off].cpu(), p[:off])  # nothing outside the range is touched
    be.adamw_step(sl(dev[0]), sl(dev[1]), sl(dev[2]), sl(dev[3]), None, None, total_sumsq=None, **dict(args, max_norm=0.0))  # optional outputs
    assert torch.isfinite(dev[0]).all()
