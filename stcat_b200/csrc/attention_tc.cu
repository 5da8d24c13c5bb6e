// tcgen05 self-attention core for the spatial encoder (reference modal_encoder.py:161-168 -> torch
// nn.MultiheadAttention bmm/softmax/bmm, functional.py:6630-6665): B = T frames x H = 8 heads of
// softmax(Q K^T * scale + key mask) V with head dim 32 and S = 1 + HW + L <= 512 tokens per frame.
//
// Forward.  Work item = (frame, head, 128-query tile).  Per item
//   TMA    : Q tile [128 x 32], K [S x 32], V [S x 32] bf16 head slices of the packed qkv buffer via 3-D tensor maps
//            {cols, S, frames} (64B swizzle, 256-row boxes).  Rows >= S of a frame are out of bounds -> zero-filled,
//            so a tile never sees its neighbour frame and padded keys contribute exactly 0.
//   MMA    : S = Q K^T        tcgen05.mma 128 x Npad x 16 (x2 for dh = 32) -> up to 256 fp32 TMEM columns per MMA
//   softmax: tcgen05.ld of the score row in 32-column chunks, masked max, exp2, row sum, P (un-normalised, relative to
//            the row max) as bf16 into a 128B-swizzled K-major smem tile; the columns of a row are split over the 2 (4)
//            softmax warpgroups of the item, which exchange their partial max / sum through shared memory
//   MMA    : O = P V          tcgen05.mma 128 x 32 x 16, ceil(S/16) steps, V as MN-major operand (no transpose)
//   epilog : tcgen05.ld O, * 1/rowsum, bf16, straight to global; lse for the backward
// S <= 256: two buffer sets (smem Q/K/V/P + 256 TMEM columns each) ping-pong, two softmax warpgroups (256 threads, a
// (row, column half) each) per set, so the tensor pipe and TMA run under the softmax of the other item.  exp2 on the
// MUFU pipe is the bound (head dim 32 gives only 64 MMA FLOP per exponential): 16 softmax warps = 4 per SM sub-partition
// keep that pipe fed (the round-1 kernel had 2 per sub-partition, one per set, and ran at a third of the MUFU rate).
// 256 < S <= 512 (BIG): one item at a time uses both sets' buffers as one (K, V: 512 rows, P: 128 x 512, S: all 512
// TMEM columns, two score MMAs) and all four softmax warpgroups (a (row, column quarter) each).
//
// Backward: see attn_tc_bwd_kernel below.
#include "tc_common.cuh"
#include <math.h>
#include <stdlib.h>
#include <type_traits>

namespace stcat {
namespace tc {

constexpr int AT_DH = 32;
constexpr int AT_QT = 128;          // query rows per item
constexpr int AT_KBOX = 256;        // key rows per TMA box / per buffer set
constexpr int AT_KMAX = 512;        // max keys (= max S)
constexpr int AT_Q_BYTES = AT_QT * AT_DH * 2;      // 8 KB
constexpr int AT_KV_BYTES = AT_KBOX * AT_DH * 2;   // 16 KB
constexpr int AT_P_BYTES = AT_QT * AT_KBOX * 2;    // 64 KB
constexpr int AT_SET_BYTES = AT_Q_BYTES + 2 * AT_KV_BYTES + AT_P_BYTES;  // 104 KB
constexpr int AT_XCH_BYTES = 2 * 4 * AT_QT * 4;    // partial row max / row sum of the 4 softmax warpgroups
constexpr int AT_FWD_SMEM = 2 * AT_SET_BYTES + 1024 + 256 + AT_XCH_BYTES;
constexpr int AT_FWD_THREADS = 128 + 512;          // 4 control warps + 16 softmax warps

struct AttnFwdParams {
    __nv_bfloat16* o;
    int64_t ldo;
    const uint8_t* key_mask;  // [B, S] or null
    float* lse;               // [B, H, S]
    int B, H, S, nqt;
    float scale;
    DropArgs drop;            // dropout on the probabilities (DROP instantiation only)
    long long* trace;         // diagnostics (stcat_debug_attn_trace): SM clock at the phase boundaries of CTA 0's first 8 items
};

// two finite floats -> packed bf16x2 with integer ops (round half away from zero: +0x8000 on the bit pattern grows the
// magnitude of either sign; differs from round-to-nearest-even only on exact ties)
__device__ __forceinline__ uint32_t pack_prob_bf16x2(float lo, float hi) {
    return __byte_perm(__float_as_uint(lo) + 0x8000u, __float_as_uint(hi) + 0x8000u, 0x7632);
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void named_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// SEP (S <= 224, not BIG): the output accumulator has its own TMEM columns (S: set * 224, O: 448 + set * 32), so the score
// MMA of the set's next item is issued as soon as the current item's P is written (before its PV MMA), and the softmax
// threads read an item's O out one item later (between the row-max pass and the exp pass of the set's next item), when it
// has long been complete: neither the PV MMA nor the epilogue sits on the softmax threads' critical path, and the two
// sets' exp passes (the MUFU-bound part) interleave instead of idling the pipe together.
template <bool DROP, bool BIG, bool SEP>
__global__ void __launch_bounds__(AT_FWD_THREADS, 1)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AttnFwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = base + 2 * AT_SET_BYTES;
    // per set: 0 qk_full (Q, K landed), 1 v_free (PV MMA done with V / P), 2 s_full (score MMA done: S readable, Q / K free),
    //          3 p_full (P written), 4 o_full, 5 tmem_free (O read out), 6 v_full (V landed).  BIG uses set 0's only.
    auto bar = [&](int set, int which) { return bar_base + 8u * (set * 7 + which); };
    const uint32_t tmem_slot = bar_base + 8u * 14;
    const uint32_t xch = bar_base + 256;  // float [2][4][128]: partial row max, partial row sum per softmax warpgroup
    auto sQ = [&](int set) { return base + set * AT_SET_BYTES; };
    auto sK = [&](int set) { return base + set * AT_SET_BYTES + AT_Q_BYTES; };
    auto sV = [&](int set) { return base + set * AT_SET_BYTES + AT_Q_BYTES + AT_KV_BYTES; };
    auto sP = [&](int set) { return base + set * AT_SET_BYTES + AT_Q_BYTES + 2 * AT_KV_BYTES; };
    static_assert(!(BIG && SEP), "SEP needs two buffer sets");
    constexpr int NG = BIG ? 4 : 2;        // softmax warpgroups per item
    constexpr int LAG = BIG ? 0 : 1;       // items between a score MMA and its PV MMA in the issue order
    // TMEM columns of a set's score tile and of its output accumulator
    auto col_s = [&](int set) { return (uint32_t)(SEP ? set * 224 : set * 256); };
    auto col_o = [&](int set) { return (uint32_t)(SEP ? 448 + set * 32 : set * 256); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total = p.B * p.H * p.nqt;
    const int n_mine = (total > (int)blockIdx.x) ? (total - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int S = p.S;
    const int nk16 = (S + 15) >> 4;          // PV contraction steps = 16-column units of a score row
    const int npad = nk16 << 4;              // N of the score MMA(s)

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmK)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmV)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar(s, 0), 1); mbar_init(bar(s, 1), 1); mbar_init(bar(s, 2), 1);
            mbar_init(bar(s, 3), NG * 128); mbar_init(bar(s, 4), 1); mbar_init(bar(s, 5), NG * 128);
            mbar_init(bar(s, 6), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // P tiles start as zeros: key columns the softmax never writes must not hold NaN bit patterns
    for (int i = threadIdx.x; i < 2 * AT_P_BYTES / 16; i += AT_FWD_THREADS) {
        const int set = i / (AT_P_BYTES / 16), off = i % (AT_P_BYTES / 16);
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sP(set) + off * 16), "r"(0) : "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    pdl_launch_dependents();
    pdl_wait();
    const DropArgs dr = DROP ? drop_resolve(p.drop) : p.drop;

    auto decode = [&](int i, int& b, int& h, int& qt) {
        const int w = blockIdx.x + i * gridDim.x;
        qt = w % p.nqt;
        const int t = w / p.nqt;
        h = t % p.H;
        b = t / p.H;
    };
    // buffer set and per-set sequence number of item i
    auto set_of = [&](int i) { return BIG ? 0 : (i & 1); };
    auto seq_of = [&](int i) { return BIG ? i : (i >> 1); };

    // diagnostics: event ev of item i -> trace[i * 16 + ev] (CTA 0, first 8 items, one lane per role)
    auto T = [&](int i, int ev) {
        if (p.trace != nullptr && blockIdx.x == 0 && i < 8) p.trace[i * 16 + ev] = clock64();
    };

    if (warp == 0) {
        // Q / K loader: the buffers of a set are free as soon as the score MMA of the set's previous item has completed
        // (s_full), i.e. one whole softmax + PV + epilogue before they are needed again -> the load latency is hidden
        if (lane == 0) {
            for (int i = 0; i < n_mine; ++i) {
                const int set = set_of(i), k = seq_of(i);
                int b, h, qt;
                decode(i, b, h, qt);
                if (k > 0) mbar_wait(bar(set, 2), (k - 1) & 1);
                T(i, 0);
                mbar_expect_tx(bar(set, 0), AT_Q_BYTES + (BIG ? 2 : 1) * AT_KV_BYTES);
                tma_load_3d(sQ(set), &tmQ, bar(set, 0), h * AT_DH, qt * AT_QT, b);
                tma_load_3d(sK(set), &tmK, bar(set, 0), h * AT_DH, 0, b);
                if (BIG) tma_load_3d(sK(1), &tmK, bar(set, 0), h * AT_DH, AT_KBOX, b);
            }
        }
    } else if (warp == 3) {
        // V loader: V is read by the PV MMA only, after the softmax, so its load (issued when the previous PV of the set
        // is done) runs under the score MMA and the softmax of its own item
        if (lane == 0) {
            for (int i = 0; i < n_mine; ++i) {
                const int set = set_of(i), k = seq_of(i);
                int b, h, qt;
                decode(i, b, h, qt);
                mbar_wait(bar(set, 1), (k & 1) ^ 1);
                T(i, 1);
                mbar_expect_tx(bar(set, 6), (BIG ? 2 : 1) * AT_KV_BYTES);
                tma_load_3d(sV(set), &tmV, bar(set, 6), h * AT_DH, 0, b);
                if (BIG) tma_load_3d(sV(1), &tmV, bar(set, 6), h * AT_DH, AT_KBOX, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_s0 = make_idesc(128, BIG ? 256 : npad, false, false);
            const uint32_t idesc_s1 = make_idesc(128, BIG ? npad - 256 : 16, false, false);
            const uint32_t idesc_o = make_idesc(128, AT_DH, false, true);
            auto issue_score = [&](int i) {
                const int set = set_of(i);
#pragma unroll
                for (int kk = 0; kk < AT_DH / 16; ++kk) {
                    const uint64_t ad = make_desc(sQ(set) + kk * 32, 16, 512, LAYOUT_SW64);
                    const uint64_t bd = make_desc(sK(set) + kk * 32, 16, 512, LAYOUT_SW64);
                    umma_bf16(tmem_base + col_s(set), ad, bd, idesc_s0, kk > 0 ? 1u : 0u);
                }
                if (BIG) {
#pragma unroll
                    for (int kk = 0; kk < AT_DH / 16; ++kk) {
                        const uint64_t ad = make_desc(sQ(0) + kk * 32, 16, 512, LAYOUT_SW64);
                        const uint64_t bd = make_desc(sK(1) + kk * 32, 16, 512, LAYOUT_SW64);
                        umma_bf16(tmem_base + 256, ad, bd, idesc_s1, kk > 0 ? 1u : 0u);
                    }
                }
                umma_commit(bar(set, 2));
            };
            auto issue_pv = [&](int j) {
                const int set = set_of(j);
                for (int t = 0; t < nk16; ++t) {
                    const int ps = BIG ? (t >> 4) : set, tt = t & 15;  // 16 steps (256 keys) per buffer set
                    const uint64_t ad = make_desc(sP(ps) + (tt >> 2) * 16384 + (tt & 3) * 32, 16, 1024, LAYOUT_SW128);
                    const uint64_t bd = make_desc(sV(ps) + tt * 1024, 512, 512, LAYOUT_SW64);
                    umma_bf16(tmem_base + col_o(set), ad, bd, idesc_o, t > 0 ? 1u : 0u);
                }
                umma_commit(bar(set, 4));
                umma_commit(bar(set, 1));
            };
            if (SEP) {
                for (int i = 0; i < 2 && i < n_mine; ++i) {
                    mbar_wait(bar(i, 0), 0);
                    T(i, 2);
                    T(i, 3);
                    tc_fence_after();
                    issue_score(i);
                }
                for (int j = 0; j < n_mine; ++j) {
                    const int set = j & 1, k = j >> 1;
                    mbar_wait(bar(set, 3), k & 1);   // P(j) written: the set's S columns are free
                    T(j, 4);
                    if (j + 2 < n_mine) {
                        mbar_wait(bar(set, 0), (k + 1) & 1);
                        T(j + 2, 2);
                        T(j + 2, 3);
                        tc_fence_after();
                        issue_score(j + 2);
                    }
                    mbar_wait(bar(set, 6), k & 1);
                    mbar_wait(bar(set, 5), (k & 1) ^ 1);  // O columns of the set read out by the (deferred) epilogue of item j - 2
                    T(j, 5);
                    tc_fence_after();
                    issue_pv(j);
                    T(j, 6);
                }
            } else {
                for (int it = 0; it < n_mine + LAG; ++it) {
                    if (it < n_mine) {
                        const int i = it, set = set_of(i), k = seq_of(i);
                        mbar_wait(bar(set, 0), k & 1);
                        T(i, 2);
                        // S and O share TMEM columns: wait until the epilogue of the set's previous item has read O out
                        mbar_wait(bar(set, 5), (k & 1) ^ 1);
                        T(i, 3);
                        tc_fence_after();
                        issue_score(i);
                    }
                    if (it >= LAG) {
                        const int j = it - LAG, set = set_of(j), k = seq_of(j);
                        mbar_wait(bar(set, 3), k & 1);
                        T(j, 4);
                        mbar_wait(bar(set, 6), k & 1);
                        T(j, 5);
                        tc_fence_after();
                        issue_pv(j);
                        T(j, 6);
                    }
                }
            }
        }
    } else if (warp >= 4) {
        const int w = warp - 4;
        const int wg = w >> 2;               // softmax warpgroup
        const int q4 = warp & 3;             // TMEM lane quadrant of this warp
        const int row = q4 * 32 + lane;
        const int set = BIG ? 0 : (wg >> 1);
        const int part = BIG ? wg : (wg & 1);          // which share of the row's columns
        const int xbar = BIG ? 1 : 1 + set;            // named barrier of the item's softmax warpgroups
        const float sc = p.scale * 1.4426950408889634f;
        // this thread's columns: 16-column units [u0, u1) of the nk16 units of a row
        const int u0 = (part * nk16) / NG, u1 = ((part + 1) * nk16) / NG;
        const int c0 = u0 * 16, nfull = (u1 - u0) >> 1;
        const bool tail = ((u1 - u0) & 1) != 0;
        const bool tr = (w == 0 && lane == 0);
        const uint32_t xmax = xch, xsum = xch + 4 * AT_QT * 4;
        // ---- epilogue of an item: this thread's share of the 32 output columns, 1 / rowsum, bf16, lse ----
        auto epilogue = [&](int i, int b, int h, int qt, int k, float sum, float mx) {
            mbar_wait(bar(set, 4), k & 1);
            if (tr) T(i, 12);
            tc_fence_after();
            constexpr int OC = AT_DH / NG;  // 16 or 8 columns
            uint32_t r[OC];
            const uint32_t t_o = tmem_base + col_o(set) + part * OC + ((uint32_t)(q4 * 32) << 16);
            if constexpr (OC == 16) tmem_ld16(t_o, r); else tmem_ld8(t_o, r);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(bar(set, 5));
            const int q = qt * AT_QT + row;
            if (q < S) {
                const float inv = sum > 0.f ? 1.f / sum : 0.f;
                uint32_t ob[OC / 2];
#pragma unroll
                for (int e = 0; e < OC; e += 2) {
                    __nv_bfloat162 bb = __floats2bfloat162_rn(__uint_as_float(r[e]) * inv, __uint_as_float(r[e + 1]) * inv);
                    ob[e >> 1] = *reinterpret_cast<uint32_t*>(&bb);
                }
                uint4* dst = reinterpret_cast<uint4*>(p.o + ((int64_t)b * S + q) * p.ldo + h * AT_DH + part * OC);
#pragma unroll
                for (int j = 0; j < OC / 8; ++j) dst[j] = make_uint4(ob[4 * j], ob[4 * j + 1], ob[4 * j + 2], ob[4 * j + 3]);
                if (part == 0) p.lse[((int64_t)b * p.H + h) * S + q] = sum > 0.f ? mx * p.scale + logf(sum) : -INFINITY;
            }
            if (tr) T(i, 13);
        };
        bool have_prev = false;   // SEP: the previous item of this set, whose epilogue is still to run
        int pv_b = 0, pv_h = 0, pv_qt = 0, pv_k = 0, pv_i = 0;
        float pv_sum = 0.f, pv_mx = 0.f;
        for (int i = BIG ? 0 : set; i < n_mine; i += (BIG ? 1 : 2)) {
            const int k = seq_of(i);
            int b, h, qt;
            decode(i, b, h, qt);
            if (tr) T(i, 7);
            const bool live = qt * AT_QT + q4 * 32 < S;   // warp-uniform: any valid query row in this warp's 32
            // key mask -> one bit per key of this thread's chunks (1 = masked), identical in every lane
            uint32_t mw[5];
#pragma unroll
            for (int c = 0; c < 5; ++c) {
                const int key = c0 + c * 32 + lane;
                bool m = key >= S;
                if (!m && p.key_mask) m = p.key_mask[(int64_t)b * S + key] != 0;
                mw[c] = __ballot_sync(0xffffffffu, m);
            }
            // dropout with precomputed keep bits (DropArgs::bits): the words covering this thread's columns of its row, loaded
            // here so that their latency hides behind the wait for the score MMA
            uint32_t kwr[6] = {0u, 0u, 0u, 0u, 0u, 0u};
            if (DROP && dr.bits != nullptr) {
                const int qrow = qt * AT_QT + row;
                if (qrow < S) {
                    const uint32_t* rb = dr.bits + (((int64_t)b * p.H + h) * S + qrow) * dr.wpr + (c0 >> 5);
                    const int nw = dr.wpr - (c0 >> 5);
#pragma unroll
                    for (int j = 0; j < 6; ++j) kwr[j] = j < nw ? __ldg(rb + j) : 0u;
                }
            }
            if (tr) T(i, 8);
            mbar_wait(bar(set, 2), k & 1);
            if (tr) T(i, 9);
            tc_fence_after();
            const uint32_t t_row = tmem_base + col_s(set) + c0 + ((uint32_t)(q4 * 32) << 16);
            // ---- pass 1: row max over this thread's columns ----
            float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
            if (live) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c < nfull) {  // warp-uniform
                        uint32_t r[32];
                        tmem_ld32(t_row + c * 32, r);
                        tmem_ld_wait();
                        const uint32_t wm = mw[c];
                        if (wm == 0u) {
#pragma unroll
                            for (int e = 0; e < 32; e += 4) {
                                m0 = fmaxf(m0, __uint_as_float(r[e]));
                                m1 = fmaxf(m1, __uint_as_float(r[e + 1]));
                                m2 = fmaxf(m2, __uint_as_float(r[e + 2]));
                                m3 = fmaxf(m3, __uint_as_float(r[e + 3]));
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 32; ++e)
                                if (!((wm >> e) & 1u)) m0 = fmaxf(m0, __uint_as_float(r[e]));
                        }
                    }
                }
                if (tail) {
                    uint32_t r[16];
                    tmem_ld16(t_row + nfull * 32, r);
                    tmem_ld_wait();
                    const uint32_t wm = (nfull == 0 ? mw[0] : nfull == 1 ? mw[1] : nfull == 2 ? mw[2] : nfull == 3 ? mw[3] : mw[4]);
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        if (!((wm >> e) & 1u)) m1 = fmaxf(m1, __uint_as_float(r[e]));
                }
            }
            float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(xmax + (wg * AT_QT + row) * 4), "f"(mx) : "memory");
            named_bar_sync(xbar, NG * 128);
#pragma unroll
            for (int g2 = 0; g2 < NG; ++g2) {
                float o;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(o) : "r"(xmax + (((BIG ? 0 : set * 2) + g2) * AT_QT + row) * 4));
                mx = fmaxf(mx, o);
            }
            if (tr) T(i, 10);
            if (SEP && have_prev) epilogue(pv_i, pv_b, pv_h, pv_qt, pv_k, pv_sum, pv_mx);
            const float ms = (mx == -INFINITY) ? 0.f : mx * sc;
            // ---- pass 2: p = exp2(s * sc - ms), partial row sum, P as bf16 into the swizzled smem tile ----
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            // dropout: element index of (b, h, query, key 0) in the [B, H, S, S] probability tensor; the row sum (and so
            // lse and the 1/sum of the epilogue) stays that of the undropped softmax
            const uint64_t drow = DROP ? (((uint64_t)b * p.H + h) * S + (qt * AT_QT + row)) * (uint64_t)S : 0ull;
            auto exp_chunk = [&](auto wtag, const uint32_t* r, const int col, const uint32_t wm, const uint32_t kw) {
                constexpr int W = decltype(wtag)::value;  // kw: keep bits of the chunk's columns (bit e = column col + e)
                uint32_t pk[W / 2];
#pragma unroll
                for (int e = 0; e < W; e += 4) {
                    float p0 = ex2_approx(fmaf(__uint_as_float(r[e]), sc, -ms));
                    float p1 = ex2_approx(fmaf(__uint_as_float(r[e + 1]), sc, -ms));
                    float p2 = ex2_approx(fmaf(__uint_as_float(r[e + 2]), sc, -ms));
                    float p3 = ex2_approx(fmaf(__uint_as_float(r[e + 3]), sc, -ms));
                    if (wm != 0u) {  // a chunk with masked keys (warp-uniform; rare: the padded tail / text padding)
                        p0 = ((wm >> e) & 1u) ? 0.f : p0;
                        p1 = ((wm >> (e + 1)) & 1u) ? 0.f : p1;
                        p2 = ((wm >> (e + 2)) & 1u) ? 0.f : p2;
                        p3 = ((wm >> (e + 3)) & 1u) ? 0.f : p3;
                    }
                    s0 += p0; s1 += p1; s2 += p2; s3 += p3;
                    if (DROP) {
                        if (dr.bits != nullptr) {
                            p0 = ((kw >> e) & 1u) ? p0 * dr.scale : 0.f;
                            p1 = ((kw >> (e + 1)) & 1u) ? p1 * dr.scale : 0.f;
                            p2 = ((kw >> (e + 2)) & 1u) ? p2 * dr.scale : 0.f;
                            p3 = ((kw >> (e + 3)) & 1u) ? p3 * dr.scale : 0.f;
                        } else {
                            p0 *= drop_mult(dr, drow + col + e);
                            p1 *= drop_mult(dr, drow + col + e + 1);
                            p2 *= drop_mult(dr, drow + col + e + 2);
                            p3 *= drop_mult(dr, drow + col + e + 3);
                        }
                    }
                    __nv_bfloat162 b01 = __floats2bfloat162_rn(p0, p1), b23 = __floats2bfloat162_rn(p2, p3);
                    pk[e >> 1] = *reinterpret_cast<uint32_t*>(&b01);
                    pk[(e >> 1) + 1] = *reinterpret_cast<uint32_t*>(&b23);
                }
#pragma unroll
                for (int j = 0; j < W / 8; ++j) {  // 16-byte pieces = 8 keys; 64-key blocks of 16 KB, 4 per buffer set
                    const int cc = col + j * 8, blk = cc >> 6;
                    const uint32_t a = sP(BIG ? (blk >> 2) : set) + (blk & 3) * 16384 + row * 128 + ((((cc & 63) >> 3) ^ (row & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pk[4 * j]), "r"(pk[4 * j + 1]),
                                 "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3]) : "memory");
                }
            };
            if (live) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c < nfull) {
                        uint32_t r[32];
                        tmem_ld32(t_row + c * 32, r);
                        tmem_ld_wait();
                        // chunk c starts (c0 & 31) bits into word c of kwr (c0 is a multiple of 16)
                        exp_chunk(std::integral_constant<int, 32>(), r, c0 + c * 32, mw[c], __funnelshift_r(kwr[c], kwr[c + 1], c0 & 31));
                    }
                }
                if (tail) {
                    uint32_t r[16];
                    tmem_ld16(t_row + nfull * 32, r);
                    tmem_ld_wait();
                    const uint32_t wm = (nfull == 0 ? mw[0] : nfull == 1 ? mw[1] : nfull == 2 ? mw[2] : nfull == 3 ? mw[3] : mw[4]);
                    const uint32_t ka = (nfull == 0 ? kwr[0] : nfull == 1 ? kwr[1] : nfull == 2 ? kwr[2] : nfull == 3 ? kwr[3] : kwr[4]);
                    const uint32_t kb = (nfull == 0 ? kwr[1] : nfull == 1 ? kwr[2] : nfull == 2 ? kwr[3] : nfull == 3 ? kwr[4] : kwr[5]);
                    exp_chunk(std::integral_constant<int, 16>(), r, c0 + nfull * 32, wm, __funnelshift_r(ka, kb, c0 & 31));
                }
            }
            float sum = (s0 + s1) + (s2 + s3);
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(xsum + (wg * AT_QT + row) * 4), "f"(sum) : "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_fence_before();
            if (tr) T(i, 11);
            mbar_arrive(bar(set, 3));
            named_bar_sync(xbar, NG * 128);
            sum = 0.f;
#pragma unroll
            for (int g2 = 0; g2 < NG; ++g2) {  // same order in every thread of the row: identical sums
                float o;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(o) : "r"(xsum + (((BIG ? 0 : set * 2) + g2) * AT_QT + row) * 4));
                sum += o;
            }
            if (SEP) { pv_b = b; pv_h = h; pv_qt = qt; pv_k = k; pv_i = i; pv_sum = sum; pv_mx = mx; have_prev = true; }
            else epilogue(i, b, h, qt, k, sum, mx);
        }
        if (SEP && have_prev) epilogue(pv_i, pv_b, pv_h, pv_qt, pv_k, pv_sum, pv_mx);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}


// ==================================================================================================
// Backward.  Work item = (frame, head): all S <= 512 queries and keys.  TMA brings Q, K, V, dO head slices
// [S x 32] (64B swizzle, rows >= S zero-filled).  For every (128-key tile kt, 128-query tile qt) block, kt outer:
//   MMA     S  = Q_qt K_kt^T, dP = dO_qt V_kt^T                 2 x (128 x 128 x 32) -> 2 x 128 TMEM columns
//   threads p  = exp2(s*scale*log2e - lse*log2e), ds = p (dp - delta) scale, delta_i = dO_i . O_i;
//           256 threads: a (row, 64-column half) each; P and dS as bf16 into 128B-swizzled smem tiles [128 x 128]
//   MMA     dV_kt += P^T dO_qt, dK_kt += dS^T Q_qt   (the P / dS tiles read as MN-major A operands: no transpose)
//           dQ_qt += dS K_kt                           (the dS tile read as K-major A operand)
// dK, dV (per key tile) and dQ (per item, all query tiles) accumulate in TMEM and are written as bf16 straight to global.
// TMEM columns: S 0..127 | dP 128..255 | dK 256 + 32 buf | dV 320 + 32 buf | dQ 384 + 32 slot (all 512 columns).
// Pipeline: the threads copy a block's S / dP into registers and release the columns at once, so the score MMAs of block
// n+1 are issued before the gradient MMAs of block n and run under the threads' exp / pack work; the threads wait for the
// gradient MMAs of block n-1 only just before they overwrite the P / dS tiles with block n.
// S <= 256: two load-buffer sets (the next item's operands arrive under the current item).  BIG (S <= 512): one set.
// ==================================================================================================
constexpr int AB_LOAD_BYTES = 4 * AT_KV_BYTES;             // Q, K, V, dO of <= 256 rows: 64 KB per item
constexpr int AB_TILE_BYTES = 128 * 128 * 2;               // P or dS tile: 32 KB
constexpr int AB_PRO_BYTES = 2 * (2 * 512 * 4 + 64);       // per item buffer: lse2[512], delta[512] floats + 16 key-mask words
constexpr int AB_SMEM = 2 * AB_LOAD_BYTES + 2 * AB_TILE_BYTES + 1024 + 256 + AB_PRO_BYTES;
constexpr int AB_THREADS = 384;
// S | dP | dK x 2 | dV x 2 | dQ: 4 tiles (BIG: one item's query tiles; else 2 tiles x 2 items).  dK / dV alternate between two
// buffers per key tile and dQ (S <= 256) between two per item, so the first gradient MMAs of a key tile / an item never wait
// for the read-out of the previous one.
constexpr uint32_t AB_COL_S = 0, AB_COL_DP = 128, AB_COL_DK = 256, AB_COL_DV = 320, AB_COL_DQ = 384;

struct AttnBwdParams {
    const __nv_bfloat16* o;    // forward output [B*S, ldo]
    const __nv_bfloat16* d_o;  // [B*S, lddo]
    int64_t ldo, lddo;
    __nv_bfloat16 *dq, *dk, *dv;
    int64_t lddq, lddk, lddv;
    const uint8_t* key_mask;
    const float* lse;
    int B, H, S;
    float scale;
    DropArgs drop;             // dropout on the probabilities (DROP instantiation only)
    long long* trace;          // diagnostics (TRACE instantiation only): SM clock at the phase boundaries of CTA 0's first blocks
};

__device__ __forceinline__ void store_row_bf16x32(__nv_bfloat16* dst, const uint32_t (&r)[32]) {
    uint32_t ob[16];
#pragma unroll
    for (int e = 0; e < 32; e += 2) {
        __nv_bfloat162 bb = __floats2bfloat162_rn(__uint_as_float(r[e]), __uint_as_float(r[e + 1]));
        ob[e >> 1] = *reinterpret_cast<uint32_t*>(&bb);
    }
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int j = 0; j < 4; ++j) d4[j] = make_uint4(ob[4 * j], ob[4 * j + 1], ob[4 * j + 2], ob[4 * j + 3]);
}

template <bool DROP, bool TRACE, bool BIG>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                   const AttnBwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    constexpr int NT = BIG ? 4 : 2;                              // max 128-row tiles per item
    constexpr uint32_t TEN = BIG ? 2 * AT_KV_BYTES : AT_KV_BYTES;  // bytes per operand tensor in a load buffer
    auto sQ = [&](int set) { return base + set * AB_LOAD_BYTES; };
    auto sK = [&](int set) { return base + set * AB_LOAD_BYTES + TEN; };
    auto sV = [&](int set) { return base + set * AB_LOAD_BYTES + 2 * TEN; };
    auto sDO = [&](int set) { return base + set * AB_LOAD_BYTES + 3 * TEN; };
    const uint32_t sP = base + 2 * AB_LOAD_BYTES;
    const uint32_t sDS = sP + AB_TILE_BYTES;
    const uint32_t bar_base = sDS + AB_TILE_BYTES;
    // 0,1 load_full[set]; 2,3 load_free[set]; 4 sd_full; 5 sd_free; 6 pds_full; 7 pds_free; 8 dq_full; 9 dq_free;
    // 10 dkv_full; 11 dkv_free; 12,13 pro_full[buf]; 14,15 pro_free[buf]; 16..19: second dq_full, dq_free, dkv_full, dkv_free
    auto bar = [&](int which) { return bar_base + 8u * which; };
    // accumulator buffers: dK / dV of key-tile sequence number j live in buffer j & 1; dQ of item `it` in buffer it & 1 (one
    // buffer in BIG).  Each buffer has its own full / free barrier pair, used every other time.
    constexpr int NDQ = BIG ? 1 : 2;
    auto dq_full = [&](int buf) { return bar(buf ? 16 : 8); };
    auto dq_free = [&](int buf) { return bar(buf ? 17 : 9); };
    auto dkv_full = [&](int buf) { return bar(buf ? 18 : 10); };
    auto dkv_free = [&](int buf) { return bar(buf ? 19 : 11); };
    auto col_dq = [&](int it, int qt) { return AB_COL_DQ + (uint32_t)((BIG ? 0 : (it & 1) * 2) + qt) * 32u; };
    const uint32_t tmem_slot = bar_base + 8u * 20;
    // per-item row statistics prepared by warps 2, 3 one item ahead (buffer = item & 1): lse * log2e and delta = dO . O of
    // every query row, and the key mask as 32-key words
    const uint32_t pro_base = bar_base + 256;
    auto sLse = [&](int buf) { return pro_base + buf * (AB_PRO_BYTES / 2); };
    auto sDlt = [&](int buf) { return pro_base + buf * (AB_PRO_BYTES / 2) + 512 * 4; };
    auto sMsk = [&](int buf) { return pro_base + buf * (AB_PRO_BYTES / 2) + 2 * 512 * 4; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total = p.B * p.H;
    const int n_mine = (total > (int)blockIdx.x) ? (total - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int S = p.S;
    const int nt = (S + 127) >> 7;  // query tiles == key tiles
    auto set_of = [&](int it) { return BIG ? 0 : (it & 1); };
    auto seq_of = [&](int it) { return BIG ? it : (it >> 1); };

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmK)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmV)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmDO)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(bar(i), 1);
        mbar_init(bar(4), 1); mbar_init(bar(5), 256); mbar_init(bar(6), 256); mbar_init(bar(7), 1);
        mbar_init(bar(8), 1); mbar_init(bar(9), 256); mbar_init(bar(10), 1); mbar_init(bar(11), 256);
        mbar_init(bar(12), 64); mbar_init(bar(13), 64); mbar_init(bar(14), 256); mbar_init(bar(15), 256);
        mbar_init(bar(16), 1); mbar_init(bar(17), 256); mbar_init(bar(18), 1); mbar_init(bar(19), 256);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    pdl_launch_dependents();
    pdl_wait();
    const DropArgs dr = DROP ? drop_resolve(p.drop) : p.drop;
    // diagnostics: event ev of block blk (CTA 0, its first 16 blocks) -> trace[blk * 8 + ev]
    auto T = [&](uint32_t blk, int ev) {
        if (TRACE && p.trace != nullptr && blockIdx.x == 0 && blk < 16) p.trace[blk * 8 + ev] = clock64();
    };

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < n_mine; ++it) {
                const int set = set_of(it), k = seq_of(it);
                const int w = blockIdx.x + it * gridDim.x;
                const int h = w % p.H, b = w / p.H;
                mbar_wait(bar(2 + set), (k & 1) ^ 1);
                mbar_expect_tx(bar(set), (BIG ? 2 : 1) * AB_LOAD_BYTES);
                tma_load_3d(sQ(set), &tmQ, bar(set), h * AT_DH, 0, b);
                tma_load_3d(sK(set), &tmK, bar(set), h * AT_DH, 0, b);
                tma_load_3d(sV(set), &tmV, bar(set), h * AT_DH, 0, b);
                tma_load_3d(sDO(set), &tmDO, bar(set), h * AT_DH, 0, b);
                if (BIG) {
                    tma_load_3d(sQ(set) + AT_KV_BYTES, &tmQ, bar(set), h * AT_DH, AT_KBOX, b);
                    tma_load_3d(sK(set) + AT_KV_BYTES, &tmK, bar(set), h * AT_DH, AT_KBOX, b);
                    tma_load_3d(sV(set) + AT_KV_BYTES, &tmV, bar(set), h * AT_DH, AT_KBOX, b);
                    tma_load_3d(sDO(set) + AT_KV_BYTES, &tmDO, bar(set), h * AT_DH, AT_KBOX, b);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_s = make_idesc(128, 128, false, false);   // S, dP: A K-major, B K-major
            const uint32_t idesc_t = make_idesc(128, AT_DH, true, true);   // dV, dK: A MN-major (tile^T), B MN-major
            const uint32_t idesc_q = make_idesc(128, AT_DH, false, true);  // dQ: A K-major, B MN-major
            const int nblk = nt * nt;
            const int nsteps = n_mine * nblk;
            // step s: score MMAs of block s, then gradient MMAs of block s - 1 (block = (item, kt, qt), kt outer).  BIG has one
            // load buffer: the next item's operands can only be requested once the last gradient MMAs of the current item
            // have been issued, so at an item boundary the order is reversed.
            auto issue_scores = [&](int s) {
                const int it = s / nblk, r = s - it * nblk, kt = r / nt, qt = r - kt * nt, set = set_of(it);
                if (r == 0) mbar_wait(bar(set), seq_of(it) & 1);   // the item's operands have landed
                mbar_wait(bar(5), (s & 1) ^ 1);                    // S / dP columns copied out by the threads
                T(s, 0);
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    const uint64_t aq = make_desc(sQ(set) + qt * 8192 + kk * 32, 16, 512, LAYOUT_SW64);
                    const uint64_t bk = make_desc(sK(set) + kt * 8192 + kk * 32, 16, 512, LAYOUT_SW64);
                    umma_bf16(tmem_base + AB_COL_S, aq, bk, idesc_s, kk);
                }
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    const uint64_t ad = make_desc(sDO(set) + qt * 8192 + kk * 32, 16, 512, LAYOUT_SW64);
                    const uint64_t bv = make_desc(sV(set) + kt * 8192 + kk * 32, 16, 512, LAYOUT_SW64);
                    umma_bf16(tmem_base + AB_COL_DP, ad, bv, idesc_s, kk);
                }
                umma_commit(bar(4));
            };
            auto issue_grads = [&](int g) {
                const int it = g / nblk, r = g - it * nblk, kt = r / nt, qt = r - kt * nt, set = set_of(it);
                const int nq16 = min(8, (S - qt * 128 + 15) >> 4);  // 16-row groups of valid queries
                const int nk16 = min(8, (S - kt * 128 + 15) >> 4);
                mbar_wait(bar(6), g & 1);
                T(g, 1);  // P / dS tiles written
                const int jkv = it * nt + kt;   // key-tile sequence number of this CTA
                if (r == 0) {                   // dQ buffer read out by the item that used it last (it - NDQ)
                    const int u = it / NDQ;
                    mbar_wait(dq_free(it % NDQ), (u & 1) ^ 1);
                }
                if (qt == 0) mbar_wait(dkv_free(jkv & 1), ((jkv >> 1) & 1) ^ 1);  // dK / dV buffer read out (key tile jkv - 2)
                T(g, 2);  // accumulators free: gradient MMAs issued
                tc_fence_after();
                // descriptors differ per step only in the start-address field (bits 0..13, units of 16 bytes)
                const uint64_t ap0 = make_desc(sP, 16384, 1024, LAYOUT_SW128), as0 = make_desc(sDS, 16384, 1024, LAYOUT_SW128);
                const uint64_t bd0 = make_desc(sDO(set) + qt * 8192, 512, 512, LAYOUT_SW64);
                const uint64_t bq0 = make_desc(sQ(set) + qt * 8192, 512, 512, LAYOUT_SW64);
                const uint64_t bk0 = make_desc(sK(set) + kt * 8192, 512, 512, LAYOUT_SW64);
                const uint64_t ak0 = make_desc(sDS, 16, 1024, LAYOUT_SW128);
#pragma unroll 4
                for (int t = 0; t < nq16; ++t) {  // contraction over the queries of this tile
                    umma_bf16(tmem_base + AB_COL_DV + (jkv & 1) * 32, ap0 + (uint64_t)(t * 128), bd0 + (uint64_t)(t * 64), idesc_t, (qt > 0 || t > 0) ? 1u : 0u);
                    umma_bf16(tmem_base + AB_COL_DK + (jkv & 1) * 32, as0 + (uint64_t)(t * 128), bq0 + (uint64_t)(t * 64), idesc_t, (qt > 0 || t > 0) ? 1u : 0u);
                }
#pragma unroll 4
                for (int t = 0; t < nk16; ++t) {  // contraction over the keys of this tile
                    umma_bf16(tmem_base + col_dq(it, qt), ak0 + (uint64_t)((t >> 2) * 1024 + (t & 3) * 2), bk0 + (uint64_t)(t * 64), idesc_q,
                              (kt > 0 || t > 0) ? 1u : 0u);
                }
                umma_commit(bar(7));
                if (qt == nt - 1) umma_commit(dkv_full(jkv & 1));
                if (r == nblk - 1) { umma_commit(dq_full(it % NDQ)); umma_commit(bar(2 + set)); }
            };
            for (int s = 0; s <= nsteps; ++s) {
                const bool late = BIG && s > 0 && (s % nblk) == 0;
                if (s < nsteps && !late) issue_scores(s);
                if (s >= 1) issue_grads(s - 1);
                if (s < nsteps && late) issue_scores(s);
            }
        }
    } else if (warp == 2 || warp == 3) {
        // row statistics of the NEXT item while the threads below work on the current one
        const int tid = (warp - 2) * 32 + lane;  // 0..63
        for (int it = 0; it < n_mine; ++it) {
            const int buf = it & 1;
            const int w = blockIdx.x + it * gridDim.x;
            const int h = w % p.H, b = w / p.H;
            mbar_wait(bar(14 + buf), ((it >> 1) & 1) ^ 1);
            for (int q = tid; q < nt * 128; q += 64) {
                float delta = 0.f, l2 = -INFINITY;
                if (q < S) {
                    const uint4* po = reinterpret_cast<const uint4*>(p.o + ((int64_t)b * S + q) * p.ldo + h * AT_DH);
                    const uint4* pg = reinterpret_cast<const uint4*>(p.d_o + ((int64_t)b * S + q) * p.lddo + h * AT_DH);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint4 a = __ldg(po + j), g = __ldg(pg + j);
                        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw[e]));
                            const float2 fg = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gw[e]));
                            delta = fmaf(fa.x, fg.x, delta);
                            delta = fmaf(fa.y, fg.y, delta);
                        }
                    }
                    l2 = p.lse[((int64_t)b * p.H + h) * S + q] * 1.4426950408889634f;
                }
                // -inf: padded query row or fully masked row -> every p and ds of the row is 0
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(sLse(buf) + q * 4), "f"(l2) : "memory");
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(sDlt(buf) + q * 4), "f"(delta) : "memory");
            }
            for (int c = warp - 2; c < nt * 4; c += 2) {  // 32 keys per word
                const int key = c * 32 + lane;
                bool m = key >= S;
                if (!m && p.key_mask) m = p.key_mask[(int64_t)b * S + key] != 0;
                const uint32_t wd = __ballot_sync(0xffffffffu, m);
                if (lane == 0) asm volatile("st.shared.b32 [%0], %1;" ::"r"(sMsk(buf) + c * 4), "r"(wd) : "memory");
            }
            mbar_arrive(bar(12 + buf));
        }
    } else if (warp >= 4) {
        const int wg = (warp - 4) >> 2;   // column half of the block: keys [wg*64, wg*64+64)
        const int q4 = warp & 3;
        const int row = q4 * 32 + lane;
        const float sc = p.scale * 1.4426950408889634f;
        const uint32_t lane_off = (uint32_t)(q4 * 32) << 16;
        const bool tr = TRACE && warp == 4 && lane == 0;
        const int nblk = nt * nt;
        const int nsteps = n_mine * nblk;
        uint32_t ndkv = 0;
        // read-outs of the accumulators are deferred by one block: by then the gradient MMAs that complete them have long
        // finished, and the score MMAs of the block in between were issued before them
        auto dkv_out = [&](int it, int kt) {  // dK (wg 0) and dV (wg 1) of a key tile: thread row = key
            const int w = blockIdx.x + it * gridDim.x;
            const int h = w % p.H, b = w / p.H;
            mbar_wait(dkv_full(ndkv & 1), (ndkv >> 1) & 1);
            tc_fence_after();
            uint32_t r0[32];
            tmem_ld32(tmem_base + lane_off + (wg ? AB_COL_DV : AB_COL_DK) + (ndkv & 1) * 32, r0);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(dkv_free(ndkv & 1));
            ++ndkv;
            __nv_bfloat16* dst = wg ? p.dv : p.dk;
            const int64_t ldd = wg ? p.lddv : p.lddk;
            const int key = kt * 128 + row;
            if (key < S) store_row_bf16x32(dst + ((int64_t)b * S + key) * ldd + h * AT_DH, r0);
        };
        auto dq_out = [&](int it) {  // dQ of every query tile: wg 0 -> dims 0..15, wg 1 -> dims 16..31
            const int w = blockIdx.x + it * gridDim.x;
            const int h = w % p.H, b = w / p.H;
            mbar_wait(dq_full(it % NDQ), (it / NDQ) & 1);
            tc_fence_after();
            uint32_t r16[NT][16];
#pragma unroll
            for (int qt = 0; qt < NT; ++qt)
                if (qt < nt) tmem_ld16(tmem_base + lane_off + col_dq(it, qt) + wg * 16, r16[qt]);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(dq_free(it % NDQ));
            mbar_arrive(bar(14 + (it & 1)));  // the item's row statistics are no longer needed
#pragma unroll
            for (int qt = 0; qt < NT; ++qt) {
                const int q = qt * 128 + row;
                if (qt < nt && q < S) {
                    uint32_t ob[8];
#pragma unroll
                    for (int e = 0; e < 16; e += 2) {
                        __nv_bfloat162 bb = __floats2bfloat162_rn(__uint_as_float(r16[qt][e]), __uint_as_float(r16[qt][e + 1]));
                        ob[e >> 1] = *reinterpret_cast<uint32_t*>(&bb);
                    }
                    uint4* dst = reinterpret_cast<uint4*>(p.dq + ((int64_t)b * S + q) * p.lddq + h * AT_DH + wg * 16);
                    dst[0] = make_uint4(ob[0], ob[1], ob[2], ob[3]);
                    dst[1] = make_uint4(ob[4], ob[5], ob[6], ob[7]);
                }
            }
        };
        auto deferred = [&](int g) {  // accumulators completed by the gradient MMAs of block g
            const int it = g / nblk, r = g - it * nblk, kt = r / nt, qt = r - kt * nt;
            if (qt == nt - 1) dkv_out(it, kt);
            if (r == nblk - 1) dq_out(it);
        };
        for (int s2 = 0; s2 < nsteps; ++s2) {
            const uint32_t nb = (uint32_t)s2;
            const int it = s2 / nblk, r = s2 - it * nblk, kt = r / nt, qt = r - kt * nt;
            const int buf = it & 1;
            const int w = blockIdx.x + it * gridDim.x;
            const int h = w % p.H, b = w / p.H;
            if (r == 0) mbar_wait(bar(12 + buf), (it >> 1) & 1);
            float l2, delta;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(l2) : "r"(sLse(buf) + (qt * 128 + row) * 4));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(delta) : "r"(sDlt(buf) + (qt * 128 + row) * 4));
            uint32_t mw2[2];
            asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(mw2[0]), "=r"(mw2[1]) : "r"(sMsk(buf) + (kt * 4 + wg * 2) * 4));
            const bool dead = l2 == -INFINITY;
            // dropout (o = (P o M) V): delta = dO . O still equals sum_j P_ij (M o dP)_ij; the P tile feeding dV is
            // P o M, and dS = P o (M o dP - delta) * scale
            const uint64_t drow = DROP ? (((uint64_t)b * p.H + h) * S + (qt * 128 + row)) * (uint64_t)S : 0ull;
            // precomputed keep bits (DropArgs::bits) of this thread's two 32-key chunks of the block: loaded before the wait for
            // the score MMAs; key column kt * 128 + wg * 64 + c * 32 is word-aligned
            uint32_t kbw[2] = {0u, 0u};
            if (DROP && dr.bits != nullptr && qt * 128 + row < S) {
                const uint32_t* rb = dr.bits + (((int64_t)b * p.H + h) * S + (qt * 128 + row)) * dr.wpr + kt * 4 + wg * 2;
                kbw[0] = kt * 4 + wg * 2 < dr.wpr ? __ldg(rb) : 0u;
                kbw[1] = kt * 4 + wg * 2 + 1 < dr.wpr ? __ldg(rb + 1) : 0u;
            }
            if (tr) T(nb, 3);  // threads ready for the block
            mbar_wait(bar(4), nb & 1);
            if (tr) T(nb, 4);  // S / dP landed in TMEM
            tc_fence_after();
            uint32_t pp[2][16], pd[2][16];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t rs[32], rd[32];
                const uint32_t col = wg * 64 + c * 32;
                tmem_ld32(tmem_base + lane_off + AB_COL_S + col, rs);
                tmem_ld32(tmem_base + lane_off + AB_COL_DP + col, rd);
                tmem_ld_wait();
                if (c == 1) {  // both chunks are in registers: the score MMAs of the next block may overwrite the columns
                    tc_fence_before();
                    mbar_arrive(bar(5));
                }
                const uint32_t mwv = mw2[c];  // warp-uniform
                const uint64_t dcol = drow + kt * 128 + col;
                if (mwv == 0u) {
                    // no masked key in this chunk (the common case): no per-element predicates; a dead row
                    // (padding / fully masked) has lse_eff = +inf, so every p and ds is exactly 0
                    const float lse_eff = dead ? INFINITY : l2;
                    const float dsc = p.scale;
#pragma unroll
                    for (int e = 0; e < 32; e += 2) {
                        const float p0 = ex2_approx(fmaf(__uint_as_float(rs[e]), sc, -lse_eff));
                        const float p1 = ex2_approx(fmaf(__uint_as_float(rs[e + 1]), sc, -lse_eff));
                        const float m0 = !DROP ? 1.f : dr.bits != nullptr ? (((kbw[c] >> e) & 1u) ? dr.scale : 0.f) : drop_mult(dr, dcol + e);
                        const float m1 = !DROP ? 1.f : dr.bits != nullptr ? (((kbw[c] >> (e + 1)) & 1u) ? dr.scale : 0.f) : drop_mult(dr, dcol + e + 1);
                        const float d0 = dead ? 0.f : p0 * (__uint_as_float(rd[e]) * m0 - delta) * dsc;
                        const float d1 = dead ? 0.f : p1 * (__uint_as_float(rd[e + 1]) * m1 - delta) * dsc;
                        __nv_bfloat162 bp = __floats2bfloat162_rn(p0 * m0, p1 * m1), bd = __floats2bfloat162_rn(d0, d1);
                        pp[c][e >> 1] = *reinterpret_cast<uint32_t*>(&bp);
                        pd[c][e >> 1] = *reinterpret_cast<uint32_t*>(&bd);
                    }
                } else {
                    const uint32_t wmask = dead ? 0xffffffffu : mwv;
#pragma unroll
                    for (int e = 0; e < 32; e += 2) {
                        float p0 = 0.f, p1 = 0.f, d0 = 0.f, d1 = 0.f;
                        if (!((wmask >> e) & 1u)) {
                            const float m0 = !DROP ? 1.f : dr.bits != nullptr ? (((kbw[c] >> e) & 1u) ? dr.scale : 0.f) : drop_mult(dr, dcol + e);
                            p0 = ex2_approx(fmaf(__uint_as_float(rs[e]), sc, -l2));
                            d0 = p0 * (__uint_as_float(rd[e]) * m0 - delta) * p.scale;
                            p0 *= m0;
                        }
                        if (!((wmask >> (e + 1)) & 1u)) {
                            const float m1 = !DROP ? 1.f : dr.bits != nullptr ? (((kbw[c] >> (e + 1)) & 1u) ? dr.scale : 0.f) : drop_mult(dr, dcol + e + 1);
                            p1 = ex2_approx(fmaf(__uint_as_float(rs[e + 1]), sc, -l2));
                            d1 = p1 * (__uint_as_float(rd[e + 1]) * m1 - delta) * p.scale;
                            p1 *= m1;
                        }
                        __nv_bfloat162 bp = __floats2bfloat162_rn(p0, p1), bd = __floats2bfloat162_rn(d0, d1);
                        pp[c][e >> 1] = *reinterpret_cast<uint32_t*>(&bp);
                        pd[c][e >> 1] = *reinterpret_cast<uint32_t*>(&bd);
                    }
                }
            }
            // the P / dS tiles are free once the gradient MMAs of the previous block have completed
            mbar_wait(bar(7), (nb & 1) ^ 1);
            if (tr) T(nb, 5);
            const uint32_t off = wg * 16384 + row * 128;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int chunk = (c * 4 + j) ^ (row & 7);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sP + off + chunk * 16), "r"(pp[c][4 * j]),
                                 "r"(pp[c][4 * j + 1]), "r"(pp[c][4 * j + 2]), "r"(pp[c][4 * j + 3]) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sDS + off + chunk * 16), "r"(pd[c][4 * j]),
                                 "r"(pd[c][4 * j + 1]), "r"(pd[c][4 * j + 2]), "r"(pd[c][4 * j + 3]) : "memory");
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (tr) T(nb, 6);  // P / dS written
            mbar_arrive(bar(6));
            if (s2 >= 1) deferred(s2 - 1);
        }
        if (nsteps >= 1) deferred(nsteps - 1);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

}  // namespace tc

// --------------------------------------------------------------------------------------------------
// dispatch (called from stcat_attention_fwd)
// --------------------------------------------------------------------------------------------------
int attn_tc_fwd_supported(int dtype, const void* q2, const void* p_avg, int B, int H, int Lq, int Lk, const void* q,
                          const void* k, const void* v, const void* o, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo) {
    if (getenv("STCAT_DISABLE_TC_ATTN")) return 0;
    if (dtype != STCAT_BF16 || q2 || p_avg) return 0;
    if (Lq != Lk || Lq < 64 || Lq > tc::AT_KMAX) return 0;
    // few short sequences (the temporal layers: one video, T <= 128 tokens) are served better by the mma.sync kernel
    // (attention_small_mma.cu, L <= 128); everything longer runs here whatever the batch
    if (Lq <= 128 && B * H < 16) return 0;
    auto al = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
    if (!al(q) || !al(k) || !al(v) || !al(o)) return 0;
    if ((ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 8)) return 0;
    return 1;
}

static long long* g_attn_trace = nullptr;
void attn_tc_set_trace(long long* buf) { g_attn_trace = buf; }

int attn_tc_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo,
                const uint8_t* key_mask, float* lse, int B, int H, int S, float scale, cudaStream_t st, const DropArgs& drop) {
    using namespace tc;
    CUtensorMap tmQ, tmK, tmV;
    int rc;
    if ((rc = make_map_3d(&tmQ, q, B, S, (int64_t)H * AT_DH, ldq, AT_DH, AT_QT, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    if ((rc = make_map_3d(&tmK, k, B, S, (int64_t)H * AT_DH, ldk, AT_DH, AT_KBOX, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    if ((rc = make_map_3d(&tmV, v, B, S, (int64_t)H * AT_DH, ldv, AT_DH, AT_KBOX, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    AttnFwdParams p;
    p.o = (__nv_bfloat16*)o;
    p.ldo = ldo;
    p.key_mask = key_mask;
    p.lse = lse;
    p.B = B; p.H = H; p.S = S;
    p.nqt = (S + AT_QT - 1) / AT_QT;
    p.scale = scale;
    p.drop = drop;
    p.trace = g_attn_trace;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(attn_tc_fwd_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_FWD_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_tc_fwd_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_FWD_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_tc_fwd_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_FWD_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_tc_fwd_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_FWD_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_tc_fwd_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_FWD_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_tc_fwd_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_FWD_SMEM);
        if (e != cudaSuccess) return set_err((int)e, "attn_tc_fwd: smem attribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    const int total = B * H * p.nqt;
    const int sms = num_sms();
    const int grid = total < sms ? total : sms;
    const bool big = S > AT_KBOX;
    auto go = [&](auto kern) { launch_pdl(kern, dim3(grid), dim3(AT_FWD_THREADS), AT_FWD_SMEM, st, tmQ, tmK, tmV, p); };
    const bool sep = !big && ((S + 15) / 16) * 16 <= 224 && !getenv("STCAT_ATTN_FWD_NOSEP");
    if (big) { if (drop.thresh) go(attn_tc_fwd_kernel<true, true, false>); else go(attn_tc_fwd_kernel<false, true, false>); }
    else if (sep) { if (drop.thresh) go(attn_tc_fwd_kernel<true, false, true>); else go(attn_tc_fwd_kernel<false, false, true>); }
    else { if (drop.thresh) go(attn_tc_fwd_kernel<true, false, false>); else go(attn_tc_fwd_kernel<false, false, false>); }
    return check_launch("attn_tc_fwd_kernel");
}

int attn_tc_bwd_supported(int dtype, const void* q2, const void* dp_avg, const void* o, int B, int H, int Lq, int Lk,
                          const void* q, const void* k, const void* v, const void* d_o, const void* dq, const void* dk,
                          const void* dv, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, int64_t lddo, int64_t lddq,
                          int64_t lddk, int64_t lddv) {
    if (getenv("STCAT_DISABLE_TC_ATTN") || getenv("STCAT_DISABLE_TC_ATTN_BWD")) return 0;
    if (dtype != STCAT_BF16 || q2 || dp_avg || !o) return 0;
    if (Lq != Lk || Lq < 64 || Lq > tc::AT_KMAX) return 0;
    if (Lq <= 128 && B * H < 16) return 0;  // see attn_tc_fwd_supported
    auto al = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
    if (!al(q) || !al(k) || !al(v) || !al(o) || !al(d_o) || !al(dq) || !al(dk) || !al(dv)) return 0;
    if ((ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 8) || (lddo % 8) || (lddq % 8) || (lddk % 8) || (lddv % 8)) return 0;
    return 1;
}

int attn_tc_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* o,
                int64_t ldo, const void* d_o, int64_t lddo, const uint8_t* key_mask, const float* lse, void* dq,
                int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int B, int H, int S, float scale,
                cudaStream_t st, const DropArgs& drop) {
    using namespace tc;
    CUtensorMap tmQ, tmK, tmV, tmDO;
    int rc;
    const int64_t E = (int64_t)H * AT_DH;
    if ((rc = make_map_3d(&tmQ, q, B, S, E, ldq, AT_DH, AT_KBOX, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    if ((rc = make_map_3d(&tmK, k, B, S, E, ldk, AT_DH, AT_KBOX, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    if ((rc = make_map_3d(&tmV, v, B, S, E, ldv, AT_DH, AT_KBOX, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    if ((rc = make_map_3d(&tmDO, d_o, B, S, E, lddo, AT_DH, AT_KBOX, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    AttnBwdParams p;
    p.o = (const __nv_bfloat16*)o; p.d_o = (const __nv_bfloat16*)d_o;
    p.ldo = ldo; p.lddo = lddo;
    p.dq = (__nv_bfloat16*)dq; p.dk = (__nv_bfloat16*)dk; p.dv = (__nv_bfloat16*)dv;
    p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
    p.key_mask = key_mask;
    p.lse = lse;
    p.B = B; p.H = H; p.S = S;
    p.scale = scale;
    p.drop = drop;
    p.trace = g_attn_trace;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(attn_tc_bwd_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_tc_bwd_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_tc_bwd_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_tc_bwd_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_tc_bwd_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM);
        if (e != cudaSuccess) return set_err((int)e, "attn_tc_bwd: smem attribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    const int total = B * H;
    const int sms = num_sms();
    const int grid = total < sms ? total : sms;
    const bool big = S > AT_KBOX;
    auto go = [&](auto kern) { launch_pdl(kern, dim3(grid), dim3(AB_THREADS), AB_SMEM, st, tmQ, tmK, tmV, tmDO, p); };
    if (big) { if (drop.thresh) go(attn_tc_bwd_kernel<true, false, true>); else go(attn_tc_bwd_kernel<false, false, true>); }
    else if (drop.thresh) go(attn_tc_bwd_kernel<true, false, false>);
    else if (g_attn_trace) go(attn_tc_bwd_kernel<false, true, false>);
    else go(attn_tc_bwd_kernel<false, false, false>);
    return check_launch("attn_tc_bwd_kernel");
}

}  // namespace stcat
