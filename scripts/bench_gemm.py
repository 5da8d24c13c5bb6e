#!/usr/bin/env python
"""Times the GEMM entry points of the C ABI alone at the shape classes of one bench step (T=64, res=448, L=16):
CUDA events over rotating operand sets larger than L2.  Prints one table; `--only NAME` restricts to one
row (used under `ncu -k regex:gemm_tc`).

    python scripts/bench_gemm.py [--only ffn1_fwd] [--iters 20]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from stcat_b200 import ops  # noqa: E402

NS, MM = 13632, 13568
#           name          kind    M    N     K     out   relu
SHAPES = [
    ("qk_fwd", "fwd", NS, 512, 256, "bf16", 0),
    ("v_fwd", "fwd", NS, 256, 256, "bf16", 0),
    ("out_fwd", "fwd", NS, 256, 256, "f32", 0),
    ("ffn1_fwd", "fwd", NS, 2048, 256, "bf16", 1),
    ("ffn2_fwd", "fwd", NS, 256, 2048, "f32", 0),
    ("mem_kv_fwd", "fwd", MM, 256, 256, "bf16", 0),
    ("ffn2_dgrad", "dgrad", NS, 256, 2048, "bf16", 0),   # dh[M,2048] = dz[M,256] . W2[256,2048]
    ("ffn1_dgrad", "dgrad", NS, 2048, 256, "f32", 0),    # dx[M,256] = dh[M,2048] . W1[2048,256]
    ("out_dgrad", "dgrad", NS, 256, 256, "bf16", 0),
    ("qk_dgrad", "dgrad", NS, 512, 256, "f32", 0),
    ("ffn1_wgrad", "wgrad", NS, 2048, 256, "f32", 0),    # dW1[2048,256] = dh^T x
    ("ffn2_wgrad", "wgrad", NS, 256, 2048, "f32", 0),    # dW2[256,2048] = dz^T h
    ("ffn1_wgrad_nob", "wgrad", NS, 2048, 256, "f32", 0),
    ("ffn2_wgrad_nob", "wgrad", NS, 256, 2048, "f32", 0),
    ("out_wgrad", "wgrad", NS, 256, 256, "f32", 0),
    ("out_wgrad_nob", "wgrad", NS, 256, 256, "f32", 0),
    ("qk_wgrad", "wgrad", NS, 512, 256, "f32", 0),
    ("dec_q_fwd", "fwd", 64, 256, 256, "f32", 0),
    ("dec_ffn1_fwd", "fwd", 64, 2048, 256, "bf16", 1),
    ("dec_ffn2_fwd", "fwd", 64, 256, 2048, "f32", 0),
    ("dec_q_wgrad", "wgrad", 64, 256, 256, "f32", 0),
    ("dec_q_wgrad_nob", "wgrad", 64, 256, 256, "f32", 0),
    ("dec_ffn1_dgrad", "dgrad", 64, 2048, 256, "f32", 0),
    ("dec_ffn2_dgrad", "dgrad", 64, 256, 2048, "bf16", 0),
    ("dec_ffn1_wgrad_nob", "wgrad", 64, 2048, 256, "f32", 0),
]


def time_shapes(names=None, iters=20, graph=True, ns=None, mm=None):
    """Device time of the listed rows (all when None).  ``ns`` / ``mm`` rescale the token counts (other clip shapes).
    Returns [{name, kind, M, N, K, out, us, tflops, gbs}]."""
    dev = torch.device("cuda", torch.cuda.current_device())
    be = ops.get_backend()
    bf = torch.bfloat16
    rows = []
    for name, kind, M, N, K, out, relu in SHAPES:
        if names is not None and name not in names:
            continue
        if ns is not None and M == NS:
            M = ns
        if mm is not None and M == MM:
            M = mm
        odt = bf if out == "bf16" else torch.float32
        nbuf = 4 if M > 1000 else 2
        if kind == "fwd":
            A = [torch.randn(M, K, device=dev).to(bf) for _ in range(nbuf)]
            W = (torch.randn(N, K, device=dev) * K ** -0.5).to(bf)
            b = torch.zeros(N, device=dev)
            C = [torch.empty(M, N, device=dev, dtype=odt) for _ in range(nbuf)]
            fn = lambda i: be.linear_fwd(A[i % nbuf], W, b, C[i % nbuf], relu=bool(relu))
            nbytes = 2 * (M * K + N * K) + C[0].element_size() * M * N
        elif kind == "dgrad":
            A = [torch.randn(M, N, device=dev).to(bf) for _ in range(nbuf)]
            W = (torch.randn(N, K, device=dev) * K ** -0.5).to(bf)
            C = [torch.empty(M, K, device=dev, dtype=odt) for _ in range(nbuf)]
            fn = lambda i: be.linear_bwd_data(A[i % nbuf], W, C[i % nbuf])
            nbytes = 2 * (M * N + N * K) + C[0].element_size() * M * K
        else:
            A = [torch.randn(M, N, device=dev).to(bf) for _ in range(nbuf)]
            X = [torch.randn(M, K, device=dev).to(bf) for _ in range(nbuf)]
            dW = torch.zeros(N, K, device=dev)
            db = torch.zeros(N, device=dev)
            fn = lambda i: be.linear_bwd_weight(A[i % nbuf], X[i % nbuf], dW, None if name.endswith("_nob") else db, accumulate=True)
            nbytes = 2 * (M * N + M * K) + 4 * N * K
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if not graph:
            e0.record()
            for i in range(iters):
                fn(i)
            e1.record()
        else:  # replay from a CUDA graph: device time, not the host's launch rate (tensor-map encode + ctypes)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=side):
                for i in range(iters):
                    fn(i)
            gr.replay()
            torch.cuda.synchronize()
            e0.record()
            gr.replay()
            e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / iters * 1e3
        fl = 2.0 * M * N * K
        rows.append({"name": name, "kind": kind, "M": M, "N": N, "K": K, "out": out, "us": us, "tflops": fl / us / 1e6,
                     "gbs": nbytes / us / 1e3, "flops": fl})
    return rows


# the GEMMs of one spatial encoder layer, forward + backward, as the step issues them (weight gradients without the fused
# bias column sum: the bias gradients come from the LayerNorm backward / the dgrad epilogue)
ENCODER_LAYER_FAMILY = ["qk_fwd", "v_fwd", "out_fwd", "ffn1_fwd", "ffn2_fwd", "ffn2_dgrad", "ffn1_dgrad", "out_dgrad", "qk_dgrad",
                        "ffn1_wgrad_nob", "ffn2_wgrad_nob", "out_wgrad_nob", "qk_wgrad"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    print(f"{'name':18s} {'kind':6s} {'M':>6s} {'N':>5s} {'K':>5s} {'out':>4s} {'us':>8s} {'TFLOP/s':>8s} {'GB/s(alg)':>9s}")
    for r in time_shapes([args.only] if args.only else None, args.iters, not args.no_graph):
        print(f"{r['name']:18s} {r['kind']:6s} {r['M']:6d} {r['N']:5d} {r['K']:5d} {r['out']:>4s} {r['us']:8.1f} {r['tflops']:8.1f} {r['gbs']:9.1f}", flush=True)
    fam = time_shapes(ENCODER_LAYER_FAMILY, args.iters, not args.no_graph) if not args.only else []
    if fam:
        us, fl = sum(r["us"] for r in fam), sum(r["flops"] for r in fam)
        print(f"encoder layer family (fwd + dgrad + wgrad, {len(fam)} GEMMs): {fl / 1e9:.1f} GFLOP in {us:.1f} us = {fl / us / 1e6:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
