#!/usr/bin/env python
"""Benchmark of the STCAT hot path on B200: clips/sec, fwd+bwd, T=64 / res=448 / L=16, one clip per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision bf16|fp32] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one synthetic clip per GPU: ground_encoder -> ground_decoder
-> prediction heads -> loss -> backward through all of it (gradients for every hot-path parameter and
for the visual / text inputs), plus -- for N > 1 -- the NCCL all-reduce of the flat gradient buffer.
Rank 0 prints ONE JSON line (see README / DESIGN.md "measurement").

`--impl reference` times the reference's own algorithm for the same step on the host CPU cores.  The
reference is pure Python/PyTorch and /root/reference does not exist on the GPU box, so this arm runs
`oracle/stcat_oracle.py` (the CPU restatement pinned against the reference's outputs, kind "port")
with all host threads.  Nothing else in this file touches `oracle/`.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL prints its version banner there) are sent to
# stderr, and the line is written to the saved original stdout.
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)

import torch  # noqa: E402

METRIC = "clips/sec (T=64, res=448) fwd+bwd"
UNIT = "clips/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("STCAT_PRECISION", "bf16"), choices=["bf16", "fp32"])
    ap.add_argument("--T", type=int, default=64)
    ap.add_argument("--res", type=int, default=448)
    ap.add_argument("--L", type=int, default=16)
    ap.add_argument("--dropout", type=float, default=0.0,
                    help="MODEL.STCAT.DROPOUT of the train-mode step (default 0 = the parity / headline configuration; the "
                         "reference's training default is 0.1, which routes attention through the dropout-capable SIMT kernels)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="do not replay the step from a CUDA graph")
    ap.add_argument("--no-optimizer", action="store_true",
                    help="time fwd + loss + bwd only (default: + the optimizer-side step: global-norm clip, AdamW, EMA, bf16 weight refresh)")
    ap.add_argument("--cpu-steps", type=int, default=2, help="timed CPU-baseline steps (bounded sample)")
    ap.add_argument("--no-prefetch", dest="e2e_prefetch", action="store_false",
                    help="e2e: copy each step's inputs on the compute stream instead of prefetching them under the previous step")
    ap.add_argument("--phases", action="store_true",
                    help="record external timing events at the phase boundaries inside the captured step and report the phase "
                         "durations of un-profiled graph replays (fused-glue path, 1 GPU)")
    ap.add_argument("--profile", default=None, help="write a per-kernel device-time table of 3 steps to this file")
    return ap.parse_args()


def workload(args):
    hw = math.ceil(args.res / 32)
    return {"T": args.T, "res": args.res, "H": hw, "W": hw, "L": args.L}


def make_cfg(T, dropout=0.0):
    from stcat_b200.config import get_default_cfg

    cfg = get_default_cfg()
    # the two shipped experiment files' loss coefficients (experiments/*.yaml) and dropout 0 (parity policy)
    cfg.merge_from_list(["INPUT.MAX_VIDEO_LEN", max(200, T), "MODEL.STCAT.DROPOUT", float(dropout), "SOLVER.GIOU_COEF", 3,
                         "SOLVER.TEMP_COEF", 10, "SOLVER.EOS_COEF", 0.3])
    return cfg


# ------------------------------------------------------------------------------------------------
# algorithmic work (SURVEY.md 8d / BASELINE.md 4): forward FLOPs, backward = 2x
# ------------------------------------------------------------------------------------------------
def flops_forward(w, d=256, F=2048, nl=6):
    T, HW, L = w["T"], w["H"] * w["W"], w["L"]
    S = 1 + HW + L
    Ns, M = T * S, T * (HW + L)
    enc_gemm = nl * Ns * (8 * d * d + 4 * d * F)
    enc_core = nl * 4 * T * S * S * d
    enc_attn_block = nl * (8 * Ns * d * d) + enc_core
    temporal = nl * ((T + 1) * (8 * d * d + 4 * d * F) + 4 * (T + 1) ** 2 * d)
    dec = nl * (6 * M * d * d + 2 * M * 768) + nl * (4 * M * d * d + 2 * M * 512)
    return {"encoder_gemm": enc_gemm, "encoder_attn_core": enc_core, "encoder_attn_block": enc_attn_block,
            "temporal": temporal, "decoder_memory_side": dec, "total": enc_gemm + enc_core + temporal + dec}


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks DURING the timed region")
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    _BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
             0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self._BITS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.01)

    def start(self):
        if self.nv is not None:
            self._stop.clear()
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm (test infrastructure used here ONLY as the
# timed CPU baseline / reference arm)
# ------------------------------------------------------------------------------------------------
def cpu_reference_step_fn(args):
    from oracle import stcat_oracle as O
    from stcat_b200 import synthetic
    from stcat_b200.param_spec import synthetic_params

    w = workload(args)
    if args.dropout:
        raise SystemExit("the CPU arm restates the reference at dropout 0 (the oracle has no dropout): drop --dropout")
    cfg = make_cfg(w["T"])
    torch.set_num_threads(os.cpu_count() or 1)
    P = synthetic_params(cfg, seed=0)
    P = {k: v.clone().requires_grad_(not k.endswith(".te")) for k, v in P.items()}
    inp = synthetic.make_inputs([w["T"]], w["H"], w["W"], w["L"], seed=42)
    tg = synthetic.make_targets([w["T"]], seed=42)
    # optimizer-side step as the reference does it (train_net.py:134-143, engine/optimizer.py:5-58): clip_grad_norm_,
    # torch.optim.AdamW with the temp_decoder LR group, EMA copy updated tensor by tensor
    with_opt = not getattr(args, "no_optimizer", False)
    trainable = {k: v for k, v in P.items() if v.requires_grad}
    temp = [v for k, v in trainable.items() if "ground_decoder.temp_decoder" in k]
    rest = [v for k, v in trainable.items() if "ground_decoder.temp_decoder" not in k]
    optim = torch.optim.AdamW([{"params": rest}, {"params": temp, "lr": float(cfg.SOLVER.TEMP_LR)}], lr=float(cfg.SOLVER.BASE_LR),
                              weight_decay=float(cfg.SOLVER.WEIGHT_DECAY)) if with_opt else None
    ema = {k: v.detach().clone() for k, v in trainable.items()} if with_opt else None

    def step():
        for v in P.values():
            v.grad = None
        vis = inp["vis_features"].clone().requires_grad_(True)
        txt = inp["text_memory"].clone().requires_grad_(True)
        out = O.hot_path_forward(P, cfg, vis, inp["vis_mask"], inp["durations"], inp["vis_pos"], inp["text_mask"], txt)
        total, _ = O.stg_loss(cfg, out, tg["boxes"], tg["actioness"], [w["T"]])
        total.backward()
        if with_opt:
            torch.nn.utils.clip_grad_norm_(list(trainable.values()), float(cfg.SOLVER.MAX_GRAD_NORM))
            optim.step()
            decay = float(cfg.MODEL.EMA_DECAY)
            with torch.no_grad():
                for k, e in ema.items():
                    e.copy_(e * decay + (1.0 - decay) * trainable[k].detach())
        return float(total.detach())

    return step


def time_cpu(args, steps, warmup):
    step = cpu_reference_step_fn(args)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return steps / dt, dt / steps


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload(args)
    v, sec = time_cpu(args, max(1, args.steps), max(0, args.warmup))
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"STCAT hot path fwd+bwd{'' if args.no_optimizer else '+optimizer step (clip, AdamW, EMA)'} (ground_encoder+ground_decoder+heads+loss), 1 clip/GPU, "
                               f"T={w['T']} res={w['res']} ({w['H']}x{w['W']} tokens) L={w['L']}, dropout 0, exact fp32 on host CPU cores",
                   "note": "reference algorithm on host CPU cores (oracle port of the pure-PyTorch reference; "
                           "/root/reference itself cannot travel to the GPU box)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} full-size clips fwd+bwd{'' if args.no_optimizer else '+optimizer step'} after {args.warmup} warm-up"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def build_b200(args, device):
    from stcat_b200 import ops, synthetic
    from stcat_b200.dp import FlatGrads, GradSync, hot_path_groups
    from stcat_b200.loss import STGLossPlan
    from stcat_b200.nested import NestedTensor
    from stcat_b200.param_spec import synthetic_params
    from stcat_b200.pipeline import STCATHotPath

    w = workload(args)
    rank = int(os.environ.get("RANK", "0"))
    cfg = make_cfg(w["T"], args.dropout)
    ops.set_precision(args.precision)
    model = STCATHotPath(cfg).load_flat_params(synthetic_params(cfg, seed=0)).to(device).train()
    inp = synthetic.make_inputs([w["T"]], w["H"], w["W"], w["L"], seed=42 + rank)
    tg = synthetic.make_targets([w["T"]], seed=42 + rank)
    plan = STGLossPlan(cfg, tg["boxes"], tg["actioness"], [w["T"]], device)
    host = {k: inp[k].pin_memory() for k in ("vis_features", "vis_pos", "text_memory")}
    dev = {k: v.to(device) for k, v in host.items()}
    vis_mask = inp["vis_mask"].to(device)
    text_mask = inp["text_mask"].to(device)
    # gradient ranges in backward-completion order (decoder + heads, encoder blocks 5..0, rest): each is all-reduced on
    # a side stream as soon as it is complete, overlapping the rest of the backward pass (stcat_b200/dp.py GradSync)
    grads = FlatGrads(model, hot_path_groups(model))
    sync = GradSync(grads, device)
    sync.prepare(model)
    if sync.active:
        from stcat_b200.decoder import _side_streams

        sync.extra_streams = _side_streams(device)
    ops.set_grad_fusion(True)  # wgrad kernels accumulate straight into the flat buffer (no per-parameter adds)
    # weight / bias gradients leave the dependent chain of the backward pass for side streams (ops.set_leaf_streams; measured
    # 5.85 -> 5.47 ms with the decoders' / temporal layers' small nodes only, 5.23 ms with every node; STCAT_LEAF_ROWS=0: off)
    leaf_rows = int(os.environ.get("STCAT_LEAF_ROWS", str(1 << 30)))
    ops.set_leaf_streams(leaf_rows > 0, leaf_rows)
    # optimizer-side step of the training loop (clips/s is over fwd + bwd + optimizer): the reference's
    # clip_grad_norm_ + AdamW (2 LR groups on the hot path) + EMA, fused (stcat_b200/optim.py), with the bf16 weight
    # shadows refreshed in the same pass
    opt = None
    if not getattr(args, "no_optimizer", False):
        from stcat_b200.optim import make_optimizer

        opt = make_optimizer(cfg, model, grads)

    zero_async = os.environ.get("STCAT_ZERO_ASYNC", "1") != "0"

    def fwd_bwd(vis, pos, txt):
        ops.mark_phase("step_start")
        if zero_async:
            grads.zero_async()  # the fill of the flat gradient buffer runs under the forward pass (joined before backward)
        else:
            grads.zero()
        vis.grad = None
        txt.grad = None
        out = model(NestedTensor(vis, vis_mask, [w["T"]]), pos, (text_mask, txt, None))
        ops.mark_phase("decoder_fwd_end")
        total, _ = plan(out)
        ops.mark_phase("loss_end")
        sync.begin_step()
        sync.attach(out)
        grads.wait_zero()
        total.backward()
        ops.join_leaf_streams()
        sync.finish()  # mean over ranks (GradSync averages, like the DDP wrapper it replaces)
        ops.mark_phase("backward_end")
        if opt is not None:
            opt.step()
        ops.mark_phase("step_end")
        return total

    return {"model": model, "cfg": cfg, "host": host, "dev": dev, "grads": grads, "fwd_bwd": fwd_bwd, "ops": ops, "w": w,
            "sync": sync, "opt": opt}


def kernel_breakdown(ctx, device):
    """One instrumented step: CUDA events around every C-ABI call, grouped by entry point."""
    be = ctx["ops"].get_backend()
    names = ["linear_fwd", "linear_bwd_data", "linear_bwd_weight", "layernorm_fwd", "layernorm_bwd", "attention_fwd",
             "attention_bwd", "add", "relu_bwd", "cast_bf16"]
    rec = []
    orig = {}

    def wrap(nm, fn):
        def inner(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            rec.append((nm, e0, e1))
            return r

        return inner

    for nm in names:
        orig[nm] = getattr(be, nm)
        setattr(be, nm, wrap(nm, orig[nm]))
    try:
        d = ctx["dev"]
        vis = d["vis_features"].clone().requires_grad_(True)
        txt = d["text_memory"].clone().requires_grad_(True)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(device)
        s0.record()
        ctx["fwd_bwd"](vis, d["vis_pos"], txt)
        s1.record()
        torch.cuda.synchronize(device)
    finally:
        for nm in names:
            setattr(be, nm, orig[nm])
    agg = {}
    for nm, e0, e1 in rec:
        a = agg.setdefault(nm, [0, 0.0])
        a[0] += 1
        a[1] += e0.elapsed_time(e1)
    total = s0.elapsed_time(s1)
    return {"step_ms_instrumented": total, "calls": {k: {"n": v[0], "ms": round(v[1], 4)} for k, v in agg.items()}}


def _graph_time_ms(fn, iters, device):
    """Device time per call of fn(i): `iters` calls captured into one CUDA graph, replayed, timed with CUDA events on the
    replay stream (no host launch overhead inside the timed region)."""
    torch.cuda.synchronize(device)
    side = torch.cuda.Stream(device)
    side.wait_stream(torch.cuda.current_stream())
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=side):
        for i in range(iters):
            fn(i)
    gr.replay()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gr.replay()
    e1.record()
    torch.cuda.synchronize(device)
    return e0.elapsed_time(e1) / iters


def encoder_attention_block(ctx, device, peaks):
    """The BASELINE metric's second half: the spatial encoder's attention block (q = k = x + pos add, packed QK / V
    in-projection, per-frame softmax(QK^T)V over S tokens, out-projection) forward at the step's shape, timed alone
    from a CUDA graph; algorithmic FLOPs 8 N_s d^2 + 4 T S^2 d (no padding FLOPs), against the measured bf16 peak."""
    be = ctx["ops"].get_backend()
    w = ctx["w"]
    T, S = w["T"], 1 + w["H"] * w["W"] + w["L"]
    R, d, H = T * S, 256, 8
    bf = torch.bfloat16
    nbuf = 3
    x = [torch.randn(R, d, device=device) for _ in range(nbuf)]
    pos = torch.randn(R, d, device=device)
    xo = [t.to(bf) for t in x]
    wi = (torch.randn(3 * d, d, device=device) / 16).to(bf)
    wo = (torch.randn(d, d, device=device) / 16).to(bf)
    bi, bo = torch.zeros(3 * d, device=device), torch.zeros(d, device=device)
    qk_in = torch.empty(R, d, device=device, dtype=bf)
    qkv = torch.empty(R, 3 * d, device=device, dtype=bf)
    o = torch.empty(R, d, device=device, dtype=bf)
    a = torch.empty(R, d, device=device)
    lse = torch.empty(T, H, S, device=device)
    scale = 32 ** -0.5

    def core(i):
        be.attention_fwd(qkv[:, :d], None, qkv[:, d:2 * d], None, qkv[:, 2 * d:], o, None, lse, None, T, H, S, S, scale)

    def block(i):
        be.add(x[i % nbuf], pos, None, qk_in)
        be.linear_group(0, [dict(terms=[(qk_in, wi[: 2 * d], bi[: 2 * d])], out=qkv[:, : 2 * d]),
                            dict(terms=[(xo[i % nbuf], wi[2 * d:], bi[2 * d:])], out=qkv[:, 2 * d:])])
        core(i)
        be.linear_fwd(o, wo, bo, a)

    for i in range(2):
        block(i)
    ms_block = _graph_time_ms(block, 12, device)
    ms_core = _graph_time_ms(core, 12, device)
    fl_block = 8.0 * R * d * d + 4.0 * T * S * S * d
    fl_core = 4.0 * T * S * S * d
    peak = peaks.get("bf16_tflops", 1590.0)
    return {"us_block": ms_block * 1e3, "us_core": ms_core * 1e3, "tflops_block": fl_block / ms_block / 1e9,
            "tflops_core": fl_core / ms_core / 1e9, "frac_of_bf16_peak_block": fl_block / ms_block / 1e9 / peak,
            "frac_of_bf16_peak_core": fl_core / ms_core / 1e9 / peak, "peak": peak,
            "shape": {"frames": T, "tokens_per_frame": S, "heads": H, "head_dim": 32},
            "note": "forward, per layer; block = x+pos add, QK/V in-projection (one grouped launch), attention core, out-projection"}


def dominant_kernel_roofline(ctx, device, peaks):
    """The dominant kernel of the step is the tcgen05 GEMM (45 % of the step's kernel time, profiles/); its largest
    launch is the FFN linear1 of the spatial encoder layers (FFN = 28.6 of 38.7 GF per layer).  Time it
    alone at the step's exact shape (M = T*S rows, N = 2048, K = 256, bias + ReLU, bf16 out): 20 launches over
    rotating buffers larger than L2, replayed from a CUDA graph, CUDA events on the replay stream."""
    be = ctx["ops"].get_backend()
    w = ctx["w"]
    M = w["T"] * (1 + w["H"] * w["W"] + w["L"])
    N, K = 2048, 256
    bf = ctx["ops"].get_precision() == "bf16"
    dt = torch.bfloat16 if bf else torch.float32
    nbuf = 4  # 4 x (7 + 56) MB of operands/outputs > 126 MB L2
    xs = [torch.randn(M, K, device=device).to(dt) for _ in range(nbuf)]
    ys = [torch.empty(M, N, device=device, dtype=dt) for _ in range(nbuf)]
    wt = (torch.randn(N, K, device=device) * K ** -0.5).to(dt)
    bias = torch.zeros(N, device=device)
    for i in range(3):
        be.linear_fwd(xs[i % nbuf], wt, bias, ys[i % nbuf], relu=True)
    ms = _graph_time_ms(lambda i: be.linear_fwd(xs[i % nbuf], wt, bias, ys[i % nbuf], relu=True), 20, device)
    flops = 2.0 * M * N * K
    achieved = flops / (ms * 1e-3) / 1e12
    peak = peaks.get("bf16_tflops", 1590.0)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("traffic_bytes_per_launch")
        except Exception:
            traffic = None
    return {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": "ncu --set full capture of this kernel at this shape, committed as "
            "profiles/dominant_kernel_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum per launch; not re-measured in this run)",
            "kernel": "stcat_linear_fwd (FFN linear1 + bias + ReLU)",
            "shape": {"M": M, "N": N, "K": K, "dtype": "bf16" if bf else "f32"}, "us_per_launch": ms * 1e3,
            "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst, kernel timed alone)" if "bf16_tflops" in peaks
            else "fallback 1590 (B200_PROFILING.md)"}


def run_b200_arm(args):
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the STCAT hot path has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # the gradient all-reduces run next to the backward pass: bound the SMs NCCL's resident CTAs take, and have the
        # persistent kernels leave exactly those free while a collective is in flight (stcat_b200/dp.py GradSync)
        os.environ.setdefault("NCCL_MAX_CTAS", "24")
        dist.init_process_group("nccl", device_id=device)
    if getattr(args, "phases", False):
        from stcat_b200 import ops as _ops

        _ops.enable_phase_marks(True)
    ctx = build_b200(args, device)
    be = ctx["ops"].get_backend()
    grads = ctx["grads"]
    d, h = ctx["dev"], ctx["host"]
    w = ctx["w"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    # static input tensors (graph-capturable): device-resident copies the step reads
    vis = d["vis_features"].clone().requires_grad_(True)
    txt = d["text_memory"].clone().requires_grad_(True)
    pos = d["vis_pos"]
    loss_out = torch.zeros((), device=device)

    # the step (fwd + loss + bwd [+ gradient all-reduce] + optimizer) as the package's replayable training step
    from stcat_b200.train import GraphedStep

    use_graph = not args.no_graph
    n_eager_warm = max(3, min(args.warmup, 3)) if use_graph else args.warmup
    l0 = be.launches
    ctx["fwd_bwd"](vis, pos, txt)  # one counted eager step: C-ABI calls (= kernels of this package) per step
    launches_per_step = be.launches - l0
    fn = lambda vis, pos, txt: ctx["fwd_bwd"](vis, pos, txt)
    try:
        gstep = GraphedStep(fn, {"vis": vis, "pos": pos, "txt": txt}, use_graph=use_graph, warmup=n_eager_warm)
    except Exception as e:  # capture is an optimisation of launch overhead, not of the math
        if rank == 0:
            import traceback

            print(f"[bench] CUDA-graph capture unavailable ({type(e).__name__}: {e}); running eager", file=sys.stderr)
            if os.environ.get("STCAT_BENCH_DEBUG"):
                traceback.print_exc()
        torch.cuda.synchronize(device)
        gstep = GraphedStep(fn, {"vis": vis, "pos": pos, "txt": txt}, use_graph=False, warmup=1)
    graph = gstep.graph
    vis, pos, txt = gstep.static["vis"], gstep.static["pos"], gstep.static["txt"]  # the step's static inputs
    loss_out = gstep.loss
    barrier()

    def step():
        gstep.replay()  # the NCCL all-reduces were captured on their side stream with the rest of the step

    for _ in range(max(0, args.warmup - n_eager_warm) + (2 if graph is not None else 0)):
        step()
    barrier()

    # ---- timed region 1: inputs resident in HBM ----
    sampler = ClockSampler(local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    sampler.stop()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * args.steps / (ms_total * 1e-3)

    # ---- timed region 2: end to end through the public API with HOST buffers ----
    h2d = sum(h[k].numel() * h[k].element_size() for k in ("vis_features", "vis_pos", "text_memory"))
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()

    # Every step's inputs cross PCIe from pinned host memory inside the timed region and every step's loss is read back
    # (the user reads it every step, train_net.py:129).  Like a data loader with pinned, non-blocking copies, the H2D of
    # step k+1 runs on a copy stream while step k computes (into staging buffers; a device-to-device hand-over into the
    # graph's static inputs starts each step), unless --no-prefetch.
    copy_stream = torch.cuda.Stream(device)
    staging = {"vis_features": torch.empty_like(vis), "vis_pos": torch.empty_like(pos), "text_memory": torch.empty_like(txt)}
    graph_in = {"vis_features": vis, "vis_pos": pos, "text_memory": txt}

    def h2d_async():
        copy_stream.wait_stream(torch.cuda.current_stream())  # staging is free once the previous hand-over is enqueued
        with torch.cuda.stream(copy_stream), torch.no_grad():
            for k, dst in staging.items():
                dst.copy_(h[k], non_blocking=True)
        return copy_stream.record_event()

    def run_e2e(n_steps):
        last = None
        ready = h2d_async() if args.e2e_prefetch else None  # the first step's copy is exposed (inside the timed region)
        for i in range(n_steps):
            with torch.no_grad():
                if args.e2e_prefetch:
                    torch.cuda.current_stream().wait_event(ready)
                    for k, dst in graph_in.items():
                        dst.copy_(staging[k], non_blocking=True)
                    if i + 1 < n_steps:
                        ready = h2d_async()
                else:
                    for k, dst in graph_in.items():
                        dst.copy_(h[k], non_blocking=True)
            step()
            loss_host.copy_(loss_out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            last = float(loss_host)
        return last

    run_e2e(3)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    last_loss = run_e2e(args.steps)
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    ms2 = max(e0.elapsed_time(e1), wall * 1e3)
    t = torch.tensor([ms2], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * args.steps / (float(t.item()) * 1e-3)

    phases = None
    if getattr(args, "phases", False) and rank == 0 and world == 1:
        # un-profiled replays; the external event nodes of the last replay hold the phase boundaries
        names = ["step_start", "encoder_fwd_end", "decoder_fwd_end", "loss_end", "decoder_bwd_end", "backward_end", "step_end"]
        evs = ctx["ops"].phase_events() or {}
        acc = {}
        reps = 10
        for _ in range(reps):
            step()
            torch.cuda.synchronize(device)
            have = [n_ for n_ in names if n_ in evs]
            for a_, b_ in zip(have[:-1], have[1:]):
                acc[f"{a_}->{b_}"] = acc.get(f"{a_}->{b_}", 0.0) + evs[a_].elapsed_time(evs[b_]) / reps
        phases = {k: round(v, 4) for k, v in acc.items()}
        sys.stderr.write("phases (ms, un-profiled graph replay): " + json.dumps(phases) + "\n")

    if args.profile and rank == 0 and world == 1:  # (the captured step holds collectives: never replay it on one rank alone)
        from torch.profiler import profile, ProfilerActivity

        torch.cuda.synchronize(device)
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for _ in range(3):
                step()
            torch.cuda.synchronize(device)
        if os.environ.get("STCAT_TRACE"):
            prof.export_chrome_trace(os.environ["STCAT_TRACE"])
        evs = [e for e in prof.key_averages() if getattr(e, "device_time_total", 0) > 0]
        evs.sort(key=lambda e: -e.device_time_total)
        tot = sum(e.device_time_total for e in evs)
        with open(args.profile, "w") as f:
            f.write(f"# 3 steps ({'CUDA graph replay' if graph is not None else 'eager'}), precision {args.precision}: "
                    f"sum of kernel device time {tot / 3e3:.3f} ms/step; timed step {ms_total / args.steps:.3f} ms\n")
            f.write("| kernel | launches/step | us/step | share |\n|---|---|---|---|\n")
            for e in evs[:60]:
                f.write(f"| `{e.key[:100]}` | {e.count / 3:.1f} | {e.device_time_total / 3:.1f} | {100 * e.device_time_total / tot:.1f}% |\n")

    line = None
    ctx["sync"].active = False  # what follows runs on rank 0 alone: no collectives in it
    if rank == 0:
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peaks = json.load(open(pk))
        fl = flops_forward(w)
        breakdown = kernel_breakdown(ctx, device)
        roof = dominant_kernel_roofline(ctx, device, peaks)
        gemm_family = None
        if args.precision == "bf16":
            try:  # time-weighted GEMM efficiency over the shapes of one spatial encoder layer (fwd + dgrad + wgrad)
                sys.path.insert(0, os.path.join(ROOT, "scripts"))
                import bench_gemm

                rows = bench_gemm.time_shapes(bench_gemm.ENCODER_LAYER_FAMILY, 10, True, ns=w["T"] * (1 + w["H"] * w["W"] + w["L"]))
                fam_us, fam_fl = sum(r["us"] for r in rows), sum(r["flops"] for r in rows)
                fam_pk = peaks.get("bf16_tflops", 1590.0)
                gemm_family = {"what": "the 13 GEMMs of one spatial encoder layer, forward + data gradients + weight gradients, each "
                                       "timed alone (CUDA-graph replay, rotating operands > L2)",
                               "gflop": fam_fl / 1e9, "us": fam_us, "tflops": fam_fl / fam_us / 1e6,
                               "frac_of_bf16_peak": fam_fl / fam_us / 1e6 / fam_pk,
                               "per_gemm_us": {r["name"]: round(r["us"], 2) for r in rows}}
            except Exception as e:
                gemm_family = {"error": f"{type(e).__name__}: {e}"}
        try:
            enc_attn = encoder_attention_block(ctx, device, peaks) if args.precision == "bf16" else None
        except Exception as e:  # a diagnostic, never fatal for the bench line
            enc_attn = {"error": f"{type(e).__name__}: {e}"}
        step_s = ms_total * 1e-3 / args.steps
        sustained = peaks.get("bf16_tflops_sustained", 1400.0)
        gemm_ms = sum(v["ms"] for k, v in breakdown["calls"].items() if k.startswith("linear"))
        attn_ms = sum(v["ms"] for k, v in breakdown["calls"].items() if k.startswith("attention"))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {
                "workload": f"STCAT hot path fwd+bwd{'' if args.no_optimizer else '+optimizer step (clip, AdamW, EMA)'} (ground_encoder+ground_decoder+heads+loss), 1 clip/GPU, "
                            f"T={w['T']} res={w['res']} ({w['H']}x{w['W']} tokens) L={w['L']}, dropout {args.dropout:g}, "
                            f"{'bf16 operands / fp32 accumulate+residual' if args.precision == 'bf16' else 'exact fp32'}",
                "parallelism": f"dp{world}", "cuda_graph": graph is not None,
                "l2": "per-step activations (>1 GB) and rotating GEMM buffers exceed the 126 MB L2; no explicit flush",
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "loss": last_loss,
                    "input_copy": "pinned host -> device every step inside the timed region; "
                                  + ("the copy of step k+1 overlaps step k on a copy stream (double-buffered staging)"
                                     if args.e2e_prefetch else "on the compute stream, not overlapped")},
            "gpu_launches": int(launches_per_step or 0) * args.steps,
            "gemm_family": gemm_family,
            "clocks": sampler.summary(),
            "roofline": roof,
            "step_flops": {"forward": fl["total"], "fwd_bwd": 3 * fl["total"],
                           "achieved_tflops": 3 * fl["total"] / step_s / 1e12,
                           "frac_of_sustained_bf16_peak": 3 * fl["total"] / step_s / 1e12 / sustained,
                           "encoder_attn_block_fwd": fl["encoder_attn_block"]},
            "encoder_attention": enc_attn,
            "kernel_breakdown": breakdown,
            **({"phases_ms": phases} if phases else {}),
            # shares of the device time of the instrumented C-ABI calls (the instrumented step itself is host-bound)
            "kernel_share": {"gemm": gemm_ms / max(sum(v["ms"] for v in breakdown["calls"].values()), 1e-9),
                             "attention": attn_ms / max(sum(v["ms"] for v in breakdown["calls"].values()), 1e-9)},
        }
        if world == 1 and not args.no_cpu_baseline and not args.dropout:  # the oracle (CPU arm) has no dropout
            try:
                v, sec = time_cpu(args, args.cpu_steps, 1)
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                        "sample": f"{args.cpu_steps} full-size clips (T={w['T']}, res={w['res']}) fwd+bwd"
                                                  f"{'' if args.no_optimizer else '+optimizer step'} "
                                                  f"after 1 warm-up, oracle port of the PyTorch reference, fp32"}
            except Exception as e:
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                        "sample": f"failed: {type(e).__name__}: {e}"}
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        # A communicator whose collectives live in a captured graph can block in ncclCommDestroy: release the graph first,
        # and never let process teardown hang the job (the line above is already printed).
        def _bail():
            time.sleep(30)
            os._exit(0)

        threading.Thread(target=_bail, daemon=True).start()
        dist.barrier()
        torch.cuda.synchronize(device)
        if graph is not None:
            graph.reset()
            del graph
        torch.cuda.synchronize(device)
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
