#!/usr/bin/env python
"""Times the short-sequence attention kernels (decoder / temporal self-attention shapes) under CUDA-graph replay."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from stcat_b200 import ops

be = ops.get_backend()
dev = "cuda"
for (B, H, L, mask, pavg, dt) in [(1, 8, 64, True, False, torch.bfloat16), (1, 8, 64, True, True, torch.bfloat16),
                                  (1, 8, 65, True, False, torch.bfloat16), (1, 8, 64, False, False, torch.float32)]:
    E = H * 32
    q, k, v, g = [torch.randn(B * L, E, device=dev).to(dt) for _ in range(4)]
    o = torch.empty_like(q); dq, dk, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
    lse = torch.empty(B, H, L, device=dev); delta = torch.empty(B, H, L, device=dev)
    km = torch.zeros(B, L, dtype=torch.uint8, device=dev) if mask else None
    pa = torch.zeros(B, L, L, device=dev) if pavg else None
    dpa = torch.randn(B, L, L, device=dev) if pavg else None
    f = lambda: be.attention_fwd(q, None, k, None, v, o, km, lse, pa, B, H, L, L, 32 ** -0.5)
    b = lambda: be.attention_bwd(q, None, k, None, v, g, km, lse, dpa, delta, dq, None, dk, None, dv, B, H, L, L, 32 ** -0.5)
    for name, fn in (("fwd", f), ("bwd", b)):
        fn(); torch.cuda.synchronize()
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=side):
            for _ in range(20):
                fn()
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        print(f"B={B} H={H} L={L} mask={mask} pavg={pavg} {str(dt)[6:]:9s} {name}: {e0.elapsed_time(e1) / 20 * 1e3:7.2f} us", flush=True)
