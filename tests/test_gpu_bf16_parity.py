"""-m gpu: parity of the BENCHMARKED path (bf16 operands, tcgen05 kernels) on the B200, through the C ABI.

Gate (see tests/layer_parity.py for why it is per layer):
  * every encoder / decoder layer, on the real activations of a full forward pass, vs the kernel-matched oracle
    (oracle.BF16_KERNEL: bf16 rounding at the same points as the kernels): forward <= 2e-3 (max-abs relative to max);
  * the same layers' gradients (inputs and parameters, random upstream gradient) vs the oracle's fp64 gradients on
    bf16-rounded operands: cosine >= 0.999 and relative L2 <= 5e-2.  A few gradients are cancellation-dominated (the key-side
    position projections of the decoder's self attention: every row of dS sums to zero and the position operand is nearly the
    same in every row, so the signal cancels and the bf16 rounding noise of dS does not); for a tensor that misses the simple
    gate the test runs the same layers through the torch emulation of the C ABI (tests/emu_backend.py: identical rounding
    points, fp32 torch arithmetic) and requires the GPU's error against the oracle to be no worse than 2x the emulation's
    (+2e-3): a kernel defect shows up as an error far above what the rounding points alone produce;
  * end to end: outputs within 3e-2 of the kernel-matched oracle -- the oracle's own end-to-end noise floor when only its
    accumulation dtype changes is 3e-3 ... 1.5e-2 (tests/test_bf16_parity_emu.py) -- and reported against the fp32 reference.
Sizes: all golden fixtures (ragged batches, both cross-attention branches), T=16/res=320, and the benchmarked T=64/res=448/L=16.
"""
import pytest
import torch

from oracle import stcat_oracle as O
from helpers import GOLDEN_CASES, load_golden, cfg_for, case_inputs, case_params, rel_err
import layer_parity as LP
from stcat_b200 import ops

pytestmark = pytest.mark.gpu

# Per-layer forward gate, max |err| / max |ref|.  Every layer but one sits at <= 1.4e-3.  The first box-decoder layer is the
# exception by construction: its content query is all zeros, so its first LayerNorm normalises a small attention output to
# unit scale, and its memory-side key is the only two-term bf16 store (k = Wkc mem + Wkp pos).  A single bf16 rounding flip
# among the ~10^5 stored key elements (fp32 summation order: one tensor-core accumulator vs the oracle's two sgemm calls)
# moves one frame's output row by ~2e-3 of the layer's max -- the torch emulation of the same rounding points on the CPU shows
# the same 1.3e-3 ... 1.7e-3 for that layer and 2e-7 for every other one (tests/test_bf16_parity_emu.py).  Recorded on B200:
# 1.2e-3 ... 2.1e-3 depending on the fixture and on the summation order of the kernels of the day.
FWD_TOL = 2e-3
FWD_TOL_FIRST_BOX_LAYER = 3e-3
E2E_TOL = 3e-2


@pytest.fixture(autouse=True)
def bf16_cuda():
    ops.set_backend(None)
    ops.set_precision("bf16")
    ops.clear_weight_cache()
    yield
    ops.set_precision("fp32")
    ops.clear_weight_cache()


def _model(cfg, P):
    from stcat_b200.pipeline import STCATHotPath

    return STCATHotPath(cfg).load_flat_params(P).cuda().eval()


def _spec(name):
    if name == "mid_T16_res320":
        return {"durations": [16], "H": 10, "W": 10, "L": 12, "seed": 3, "ragged": False, "max_video_len": 32}
    if name == "bench_T64_res448":
        return {"durations": [64], "H": 14, "W": 14, "L": 16, "seed": 7, "ragged": False, "max_video_len": 200}
    if name == "nonsquare_T12_14x23":  # short side 448, long side 720 (datasets/build.py:21-22): 322 visual tokens, S = 339
        return {"durations": [12], "H": 14, "W": 23, "L": 16, "seed": 9, "ragged": False, "max_video_len": 32}
    return load_golden(name)["spec"]


def _emu_backward_errors(cfg, P, inp, spec, fs, layers):
    """the same per-layer gradient comparison for the torch emulation of the C ABI (CPU), for the given layers"""
    from emu_backend import EmuBackend
    from stcat_b200.pipeline import STCATHotPath

    ops.set_backend(EmuBackend())
    ops.clear_weight_cache()
    try:
        m = STCATHotPath(cfg).load_flat_params(P).eval()
        recs = []
        with LP.record_layers(recs), torch.no_grad():
            LP.run_full(m, inp, "cpu")
        return LP.check_backward(recs, LP.module_names(m), P, spec["durations"], select=lambda n: n in layers, from_scratch=fs)
    finally:
        ops.set_backend(None)
        ops.clear_weight_cache()


CASES = GOLDEN_CASES + ["mid_T16_res320", "nonsquare_T12_14x23", "bench_T64_res448"]


@pytest.mark.parametrize("name", CASES)
def test_every_layer_forward_and_backward(name):
    spec = _spec(name)
    cfg = cfg_for(spec)
    inp = case_inputs(spec)
    P = case_params(cfg, spec)
    m = _model(cfg, P)
    fs = bool(spec.get("from_scratch", True))
    recs = []
    with LP.record_layers(recs), torch.no_grad():
        out = LP.run_full(m, inp, "cuda")
    torch.cuda.synchronize()
    assert len(recs) == 24
    names = LP.module_names(m)
    errs = LP.check_forward(recs, names, P, spec["durations"], from_scratch=fs)
    ranked = sorted(((max(e.values()), k) for k, e in errs.items()), reverse=True)
    print(f"[{name}] per-layer forward, worst 3:", [(f"{e:.1e}", k) for e, k in ranked[:3]])
    first = "ground_decoder.decoder.layers.0"
    assert all(e < (FWD_TOL_FIRST_BOX_LAYER if k == first else FWD_TOL) for e, k in ranked), ranked[:4]
    big = name == "bench_T64_res448"
    sel = (lambda n: n.endswith((".0", ".5"))) if big else (lambda n: n.endswith((".0", ".2", ".5")))
    back = LP.check_backward(recs, names, P, spec["durations"], select=sel, from_scratch=fs)
    worst = sorted(((1 - c, e, layer, k) for layer, r in back.items() for k, (e, c, ok, _) in r.items()), reverse=True)
    print(f"[{name}] per-layer backward, lowest cosines:", [(f"{1 - a:.5f}", f"{e:.1e}", l.split('.', 1)[1], k) for a, e, l, k in worst[:3]])
    bad = [(layer, k, e, c) for layer, r in back.items() for k, (e, c, ok, _) in r.items() if not ok]
    if bad:
        emu = _emu_backward_errors(cfg, P, inp, spec, fs, {layer for layer, _, _, _ in bad})
        still = []
        for layer, k, e, c in bad:
            l2_gpu, l2_emu = back[layer][k][3], emu[layer][k][3]
            print(f"[{name}] {layer} {k}: rel-L2 vs oracle  gpu {l2_gpu:.3e}  torch emulation {l2_emu:.3e}")
            if not l2_gpu <= 2 * l2_emu + 2e-3:
                still.append((layer, k, l2_gpu, l2_emu))
        assert not still, still[:6]
    # end to end, for the record and against the noise floor
    with torch.no_grad():
        ref = O.hot_path_forward(P, cfg, inp["vis_features"], inp["vis_mask"], inp["durations"], inp["vis_pos"], inp["text_mask"],
                                 inp["text_memory"], prec=O.BF16_KERNEL)
        ref32 = O.hot_path_forward(P, cfg, inp["vis_features"], inp["vis_mask"], inp["durations"], inp["vis_pos"],
                                   inp["text_mask"], inp["text_memory"])
    e2e = {k: (rel_err(out[k], ref[k]), rel_err(out[k], ref32[k])) for k in ("pred_boxes", "pred_sted", "pred_actioness", "weights")}
    print(f"[{name}] end to end (vs kernel-matched oracle, vs fp32 oracle):", {k: (f"{a:.1e}", f"{b:.1e}") for k, (a, b) in e2e.items()})
    for k, (a, b) in e2e.items():
        assert a < E2E_TOL, (k, a)


def test_fast_kernels_serve_every_baseline_shape():
    """No BASELINE configuration may fall onto the generic SIMT attention (csrc/attention_simt.cu): spatial S up to 512 tokens
    (res 640: S = 417; 448 x 720: S = 339) runs the tcgen05 kernel, temporal / query sequences up to MAX_VIDEO_LEN + 1 = 301
    tokens run the mma.sync kernel.  The C ABI counts launches per kernel family (stcat_debug_attn_counts)."""
    be = ops.get_backend()
    bf = torch.bfloat16
    for (B, L) in [(64, 66), (64, 117), (48, 186), (64, 213), (12, 339), (8, 417), (4, 512), (1, 65), (1, 129), (2, 201), (1, 301)]:
        q, k, v = (torch.randn(B * L, 256, device="cuda").to(bf) for _ in range(3))
        o = torch.empty_like(q)
        lse = torch.empty(B, 8, L, device="cuda")
        c0 = be.attn_counts()
        be.attention_fwd(q, None, k, None, v, o, None, lse, None, B, 8, L, L, 32 ** -0.5)
        do = torch.randn_like(q)
        dq, dk, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
        delta = torch.empty(B, 8, L, device="cuda")
        be.attention_bwd(q, None, k, None, v, do, None, lse, None, delta, dq, None, dk, None, dv, B, 8, L, L, 32 ** -0.5, o=o)
        c1 = be.attn_counts()
        d = {kk: c1[kk] - c0[kk] for kk in c1}
        assert d["simt"] == 0 and d["small"] == 0, ((B, L), d)
        assert d["tc"] + d["mma"] == 2, ((B, L), d)
    torch.cuda.synchronize()
