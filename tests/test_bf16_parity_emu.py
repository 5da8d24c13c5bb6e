"""CPU: the per-layer bf16 parity harness (tests/layer_parity.py) driven through the torch emulation of the C ABI.

What it pins without a GPU: (i) the oracle's kernel-matched rounding model (oracle.BF16_KERNEL) lists the same rounding
points as the host composition allocates bf16 tensors for -- layer by layer the two agree to fp32 summation noise plus
isolated bf16 rounding flips; (ii) the harness itself (layout conversions, layer replay, gradient comparison) before it is
spent on GPU time; (iii) the end-to-end noise floor of a bf16 implementation of this network, measured on the oracle alone.
"""
import pytest
import torch

from oracle import stcat_oracle as O
from emu_backend import EmuBackend
from helpers import load_golden, cfg_for, case_inputs, case_params, rel_err
import layer_parity as LP
from stcat_b200 import ops


@pytest.fixture(autouse=True)
def emu_bf16():
    ops.set_backend(EmuBackend())
    ops.set_precision("bf16")
    ops.clear_weight_cache()
    yield
    ops.set_backend(None)
    ops.set_precision("fp32")
    ops.clear_weight_cache()


def _model(cfg, P):
    from stcat_b200.pipeline import STCATHotPath

    return STCATHotPath(cfg).load_flat_params(P).eval()


@pytest.mark.parametrize("name", ["b2_ragged_T5_3", "b2_ragged_T4_6_mdetr"])
def test_every_layer_matches_kernel_matched_oracle(name):
    fx = load_golden(name)
    spec = fx["spec"]
    cfg = cfg_for(spec)
    inp = case_inputs(spec)
    P = case_params(cfg, spec)
    m = _model(cfg, P)
    recs = []
    with LP.record_layers(recs), torch.no_grad():
        LP.run_full(m, inp, "cpu")
    assert len(recs) == 24
    names = LP.module_names(m)
    errs = LP.check_forward(recs, names, P, spec["durations"], from_scratch=bool(spec.get("from_scratch", True)))
    worst = max(v for e in errs.values() for v in e.values())
    assert worst < 2e-3, sorted(((max(e.values()), k) for k, e in errs.items()), reverse=True)[:4]
    back = LP.check_backward(recs, names, P, spec["durations"], from_scratch=bool(spec.get("from_scratch", True)),
                             select=lambda n: n.endswith((".0", ".3")))
    for layer, r in back.items():
        for k, (e, c, ok, l2) in r.items():
            assert ok, (layer, k, e, c, l2)


def test_bf16_noise_floor_of_the_oracle_itself():
    """Same rounding points, accumulation in fp32 vs fp64: the end-to-end outputs of the kernel-matched oracle move by more
    than 1e-3 (recorded: 2e-3 ... 1.5e-2).  This is the floor under any end-to-end bf16 comparison of this network and the
    reason the tight bf16 gate is per layer (tests/layer_parity.py)."""
    fx = load_golden("b1_T8_res224_L8")
    spec = fx["spec"]
    cfg = cfg_for(spec)
    inp = case_inputs(spec)
    P = case_params(cfg, spec)
    with torch.no_grad():
        a = O.hot_path_forward(P, cfg, inp["vis_features"], inp["vis_mask"], inp["durations"], inp["vis_pos"], inp["text_mask"],
                               inp["text_memory"], prec=O.BF16_KERNEL)
        P64 = {k: v.double() for k, v in P.items()}
        b = O.hot_path_forward(P64, cfg, inp["vis_features"].double(), inp["vis_mask"], inp["durations"], inp["vis_pos"].double(),
                               inp["text_mask"], inp["text_memory"].double(), prec=O.Prec(torch.float64, "bf16", True))
    floor = {k: rel_err(a[k], b[k]) for k in ("pred_boxes", "pred_sted", "pred_actioness", "weights")}
    print("bf16 end-to-end noise floor (oracle fp32 vs fp64 accumulation):", floor)
    assert max(floor.values()) > 1e-3
    assert max(floor.values()) < 5e-2
