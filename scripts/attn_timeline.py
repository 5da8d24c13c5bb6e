"""Phase timeline of the tcgen05 spatial-attention kernels (stcat_debug_attn_trace): where one CTA's time per work item goes.

    python scripts/attn_timeline.py [T] [S] [fwd|bwd]        (needs a B200; STCAT_ATTN_FWD_V2=1 selects the staged forward)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from stcat_b200.cabi import CudaBackend

EV = ["qk_issue", "v_issue", "mma:qk_landed", "mma:tmem_free", "mma:p_full", "mma:v_landed", "mma:pv_issued", "sm:item_start",
      "sm:mask_done", "sm:s_full", "sm:pass1_end", "sm:pass2_end", "sm:o_full", "sm:item_end"]


BEV = ["mma:sd_issue", "mma:pds_full", "mma:grad_issue", "th:ready", "th:sd_full", "th:tiles_free", "th:pds_written"]


def backward(be, T, S):
    """per (query tile, key tile) block of CTA 0's first items: score MMAs -> thread phase -> gradient MMAs"""
    H, E = 8, 256
    qkv = torch.randn(T * S, 3 * E, device="cuda").to(torch.bfloat16)
    o = torch.empty(T * S, E, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(T, H, S, device="cuda")
    q, k, v = qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:]
    be.attention_fwd(q, None, k, None, v, o, None, lse, None, T, H, S, S, 32 ** -0.5)
    d_o = torch.randn(T * S, E, device="cuda").to(torch.bfloat16)
    dqkv = torch.empty_like(qkv)
    delta = torch.empty(T, H, S, device="cuda")
    run = lambda: be.attention_bwd(q, None, k, None, v, d_o, None, lse, None, delta, dqkv[:, :E], None, dqkv[:, E:2 * E], None,
                                   dqkv[:, 2 * E:], T, H, S, S, 32 ** -0.5, o=o)
    for _ in range(3):
        run()
    buf = torch.zeros(128, dtype=torch.int64, device="cuda")
    be._rc(be.lib.stcat_debug_attn_trace(buf.data_ptr()), "trace on")
    run()
    torch.cuda.synchronize()
    be._rc(be.lib.stcat_debug_attn_trace(None), "trace off")
    t = buf.cpu().view(16, 8)
    t0 = int(t[t > 0].min())
    print(f"backward T={T} S={S}: SM clocks relative to the first event (CTA 0); 4 blocks (qt, kt) per work item")
    print("blk  " + " ".join(f"{n:>15s}" for n in BEV))
    for i in range(16):
        if int(t[i].max()) == 0:
            continue
        print(f"{i:4d} " + " ".join(f"{(int(t[i, e]) - t0) if int(t[i, e]) else -1:15d}" for e in range(len(BEV))))
    print("per block: sd_issue->sd_full (score MMAs), sd_full->pds_written (thread phase), pds_written->next sd_full")
    for i in range(15):
        if int(t[i].max()) == 0 or int(t[i + 1].max()) == 0:
            continue
        print(f"{i:4d} {int(t[i,4]-t[i,0]):8d} {int(t[i,6]-t[i,4]):8d} {int(t[i+1,4]-t[i,6]):8d}")


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 213
    be = CudaBackend()
    if len(sys.argv) > 3 and sys.argv[3] == "bwd":
        return backward(be, T, S)
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
