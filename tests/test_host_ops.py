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
