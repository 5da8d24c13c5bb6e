#!/usr/bin/env python
"""Times the single-query attention kernels (the decoders' time-aligned cross attention) under CUDA-graph replay, on rotating
K / V sets larger than L2, against the HBM time of their algorithmic bytes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from stcat_b200 import ops

be = ops.get_backend()
dev = "cuda"
B, H, Lk = 64, 8, 212
E = H * 32
NSET = 10
bf = torch.bfloat16
for two in (True, False):
    scale = (64 if two else 32) ** -0.5
    mk = lambda L: torch.randn(B * L, E, device=dev).to(bf)
    sets = [dict(q1=mk(1), q2=mk(1) if two else None, k1=mk(Lk), k2=mk(Lk) if two else None, v=mk(Lk), g=mk(1),
                 dk1=mk(Lk), dk2=mk(Lk) if two else None, dv=mk(Lk)) for _ in range(NSET)]
    o, dq1, dq2 = mk(1), mk(1), (mk(1) if two else None)
    lse = torch.empty(B, H, 1, device=dev); delta = torch.empty(B, H, 1, device=dev)
    km = torch.zeros(B, Lk, dtype=torch.uint8, device=dev)
    def f(s):
        be.attention_fwd(s["q1"], s["q2"], s["k1"], s["k2"], s["v"], o, km, lse, None, B, H, 1, Lk, scale)
    def b(s):
        be.attention_bwd(s["q1"], s["q2"], s["k1"], s["k2"], s["v"], s["g"], km, lse, None, delta, dq1, dq2, s["dk1"], s["dk2"], s["dv"],
                         B, H, 1, Lk, scale, o=o)
    parts = 3 if two else 2
    for name, fn, nbytes in (("fwd", f, parts * B * Lk * E * 2), ("bwd", b, 2 * parts * B * Lk * E * 2)):
        fn(sets[0]); torch.cuda.synchronize()
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=side):
            for i in range(2 * NSET):
                fn(sets[i % NSET])
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / (2 * NSET) * 1e3
        print(f"single-query attention B={B} H={H} Lk={Lk} two_part={two} {name}: {us:7.2f} us, {nbytes / 1e6:.1f} MB -> {nbytes / us / 1e6:.2f} TB/s", flush=True)
