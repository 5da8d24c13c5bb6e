"""Phase timeline of the tcgen05 spatial-attention forward (stcat_debug_attn_trace): where one CTA's time per work item goes.

    python scripts/attn_fwd_timeline.py [T] [S]        (needs a B200)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from stcat_b200.cabi import CudaBackend

EV = ["qk_issue", "v_issue", "mma:qk_landed", "mma:tmem_free", "mma:p_full", "mma:v_landed", "mma:pv_issued", "sm:item_start",
      "sm:mask_done", "sm:s_full", "sm:pass1_end", "sm:pass2_end", "sm:o_full", "sm:item_end"]


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 213
    be = CudaBackend()
    H, E = 8, 256
    qkv = torch.randn(T * S, 3 * E, device="cuda").to(torch.bfloat16)
    o = torch.empty(T * S, E, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(T, H, S, device="cuda")
    run = lambda: be.attention_fwd(qkv[:, :E], None, qkv[:, E:2 * E], None, qkv[:, 2 * E:], o, None, lse, None, T, H, S, S, 32 ** -0.5)
    for _ in range(3):
        run()
    buf = torch.zeros(128, dtype=torch.int64, device="cuda")
    be._rc(be.lib.stcat_debug_attn_trace(buf.data_ptr()), "trace on")
    run()
    torch.cuda.synchronize()
    be._rc(be.lib.stcat_debug_attn_trace(None), "trace off")
    t = buf.cpu().view(8, 16)
    t0 = int(t[t > 0].min())
    print(f"T={T} S={S}: SM clocks relative to the first event (CTA 0); items alternate between the two buffer sets")
    print("item " + " ".join(f"{n:>14s}" for n in EV))
    for i in range(8):
        if int(t[i].max()) == 0:
            continue
        print(f"{i:4d} " + " ".join(f"{(int(t[i, e]) - t0) if int(t[i, e]) else -1:14d}" for e in range(len(EV))))
    print("per item: s_full->pass1_end, pass1->pass2_end, pass2_end->o_full (PV MMA), o_full->item_end (epilogue), item_end->next s_full (same set)")
    for i in range(8):
        if int(t[i].max()) == 0:
            continue
        nxt = int(t[i + 2, 9]) - int(t[i, 13]) if i + 2 < 8 and int(t[i + 2, 9]) else -1
        print(f"{i:4d} {int(t[i,10]-t[i,9]):8d} {int(t[i,11]-t[i,10]):8d} {int(t[i,12]-t[i,11]):8d} {int(t[i,13]-t[i,12]):8d} {nxt:8d}")


if __name__ == "__main__":
    main()
