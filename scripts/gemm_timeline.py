"""Phase timeline of the tcgen05 GEMM (stcat_debug_gemm_trace): is a tile bounded by its loads, its MMAs or its epilogue?

    python scripts/gemm_timeline.py [M] [N] [K]        (needs a B200; default = FFN linear1 forward 13632 x 2048 x 256)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stcat_b200.cabi import CudaBackend

EV = ["tma:first_free", "tma:all_issued", "mma:acc_free", "mma:first_landed", "mma:all_issued", "epi:ready", "epi:acc_full",
      "epi:stores_issued"]


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 13632
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    K = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    be = CudaBackend()
    bf = torch.bfloat16
    x = torch.randn(M, K, device="cuda").to(bf)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(bf)
    b = torch.zeros(N, device="cuda")
    y = torch.empty(M, N, device="cuda", dtype=bf)
    run = lambda: be.linear_fwd(x, w, b, y, relu=True)
    for _ in range(3):
        run()
    buf = torch.zeros(64, dtype=torch.int64, device="cuda")
    be._rc(be.lib.stcat_debug_gemm_trace(buf.data_ptr()), "trace on")
    run()
    torch.cuda.synchronize()
    be._rc(be.lib.stcat_debug_gemm_trace(None), "trace off")
    t = buf.cpu().view(8, 8)
    t0 = int(t[t > 0].min())
    print(f"GEMM {M} x {N} x {K} (bias + ReLU, bf16 out): SM clocks relative to the first event, CTA 0, one row per 128 x 256 tile")
    print("tile " + " ".join(f"{n:>17s}" for n in EV))
    for i in range(8):
        if int(t[i].max()) == 0:
            continue
        print(f"{i:4d} " + " ".join(f"{(int(t[i, e]) - t0) if int(t[i, e]) else -1:17d}" for e in range(8)))
    print("per tile: loads issued->first landed, first landed->MMAs issued, MMAs issued->acc_full (epilogue sees it), acc_full->stores issued,"
          " tile period (acc_full to acc_full)")
    for i in range(8):
        if int(t[i].max()) == 0:
            continue
        per = int(t[i + 1, 6] - t[i, 6]) if i + 1 < 8 and int(t[i + 1, 6]) else -1
        print(f"{i:4d} {int(t[i,3]-t[i,0]):8d} {int(t[i,4]-t[i,3]):8d} {int(t[i,6]-t[i,4]):8d} {int(t[i,7]-t[i,6]):8d} {per:8d}")


if __name__ == "__main__":
    main()
