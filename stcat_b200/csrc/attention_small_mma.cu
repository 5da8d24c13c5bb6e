// Short-sequence attention (Lq, Lk <= 128) for bf16 operands on the warp-level tensor-core path (mma.sync m16n8k16,
// fp32 accumulate): forward and one-kernel backward, one CTA per (batch, head), one warp per 16 query rows.
//
// Same role as attention_small.cu (the temporal self-attentions on the step's dependent chain) -- that file stays the
// exact-fp32 path; this one serves bf16 mode, where the SIMT kernels were shared-memory-latency bound (11 us forward /
// 20 us backward at L = 64 for 5 MFLOP).  tcgen05 is the wrong tool here: a 64 x 64 x 32 problem per head does not
// fill one 128-row UMMA tile, and TMEM allocation + descriptor set-up cost more than the math.
//
// Forward, per warp (16 query rows, all keys):  S = Q K^T as (Lk/8) accumulator tiles in registers, masked row max /
// exp / row sum with 4-lane shuffles, P re-used straight from the accumulator registers as the A operand of P V
// (accumulator layout of two adjacent 8-column tiles == A-fragment layout of one 16-wide k step), O = P V.
// Backward: phase 1 (warp = 16 query rows) recomputes P, forms dP = dO V^T, delta, dS, writes dQ = dS K and stores
// P^T, dS^T (bf16) in shared memory; phase 2 (warp = 16 key rows) dV = P^T dO, dK = dS^T Q.
// P and dS are rounded to bf16 for the second products, like every bf16 attention kernel (the softmax statistics, the
// head-averaged weights output and delta are fp32).
#include "common.cuh"
#include <math.h>

namespace stcat {

constexpr int MM_DH = 32;
constexpr int MM_LMAX = 128;
constexpr int MM_PR = 40;  // row pitch (bf16) of the [L][32] tiles: 80 B, fragment loads are bank-conflict free

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t lds32(const __nv_bfloat16* p) { return *reinterpret_cast<const uint32_t*>(p); }

// A fragment (16 rows x 16 k) of a row-major [rows][pitch] bf16 tile at (row0, k0)
__device__ __forceinline__ void load_a(uint32_t (&a)[4], const __nv_bfloat16* tile, int pitch, int row0, int k0, int g, int t) {
    const __nv_bfloat16* p = tile + (row0 + g) * pitch + k0 + 2 * t;
    a[0] = lds32(p);
    a[1] = lds32(p + 8 * pitch);
    a[2] = lds32(p + 8);
    a[3] = lds32(p + 8 * pitch + 8);
}
// B fragment (16 k x 8 n) where B[k][n] = tile[n0 + n][k0 + k] (tile row-major [n][pitch]: "col" operand)
__device__ __forceinline__ void load_b(uint32_t (&b)[2], const __nv_bfloat16* tile, int pitch, int n0, int k0, int g, int t) {
    const __nv_bfloat16* p = tile + (n0 + g) * pitch + k0 + 2 * t;
    b[0] = lds32(p);
    b[1] = lds32(p + 8);
}

// [LP][40] row-major tile and (optionally) its transpose [32][pitchT] from rows of a global bf16 matrix; rows >= L zero
__device__ __forceinline__ void mm_load(__nv_bfloat16* dst, __nv_bfloat16* dstT, int pitchT, const __nv_bfloat16* __restrict__ src,
                                        int64_t ld, int64_t base_row, int col0, int L, int LP) {
    for (int i = threadIdx.x; i < LP * 4; i += blockDim.x) {
        const int r = i >> 2, c = (i & 3) * 8;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (r < L) v = *reinterpret_cast<const uint4*>(src + (base_row + r) * ld + col0 + c);
        *reinterpret_cast<uint4*>(dst + r * MM_PR + c) = v;
        if (dstT) {
            const __nv_bfloat16* e = reinterpret_cast<const __nv_bfloat16*>(&v);
#pragma unroll
            for (int j = 0; j < 8; ++j) dstT[(c + j) * pitchT + r] = e[j];
        }
    }
}

template <int NT>  // NT = key tiles of 8 (Lk padded to 16 -> NT even), compile-time so the accumulators stay in registers
__global__ void __launch_bounds__(256)
attn_mma_fwd_kernel(const __nv_bfloat16* __restrict__ q, int64_t ldq, const __nv_bfloat16* __restrict__ k, int64_t ldk,
                    const __nv_bfloat16* __restrict__ v, int64_t ldv, __nv_bfloat16* __restrict__ o, int64_t ldo,
                    const uint8_t* __restrict__ key_mask, float* __restrict__ lse, float* __restrict__ p_avg, int H, int Lq,
                    int Lk, float scale, const DropArgs drop_in) {
    const DropArgs drop = drop_resolve(drop_in);
    extern __shared__ __align__(16) uint8_t mm_smem[];
    pdl_launch_dependents();
    pdl_wait();
    constexpr int LKP = NT * 8;
    constexpr int PT = LKP + 8;  // pitch of the transposed V tile
    const int LQP = (Lq + 15) & ~15;
    __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(mm_smem);
    __nv_bfloat16* Ks = Qs + LQP * MM_PR;
    __nv_bfloat16* VsT = Ks + LKP * MM_PR;  // [32][PT]
    __nv_bfloat16* Vtmp = VsT + MM_DH * PT;  // row-major scratch for the load helper
    __shared__ uint8_t msk[MM_LMAX];
    const int h = blockIdx.x, b = blockIdx.y;
    const int col = h * MM_DH;
    mm_load(Qs, nullptr, 0, q, ldq, (int64_t)b * Lq, col, Lq, LQP);
    mm_load(Ks, nullptr, 0, k, ldk, (int64_t)b * Lk, col, Lk, LKP);
    mm_load(Vtmp, VsT, PT, v, ldv, (int64_t)b * Lk, col, Lk, LKP);
    for (int j = threadIdx.x; j < LKP; j += blockDim.x) msk[j] = (j >= Lk) || (key_mask && key_mask[(int64_t)b * Lk + j]);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int row0 = warp * 16;
    if (row0 >= LQP) return;
    // ---- S = Q K^T ----
    float s[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < MM_DH; kk += 16) {
        uint32_t a[4];
        load_a(a, Qs, MM_PR, row0, kk, g, t);
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            uint32_t bb[2];
            load_b(bb, Ks, MM_PR, j * 8, kk, g, t);
            mma_bf16_16816(s[j], a, bb);
        }
    }
    // ---- masked softmax over the keys; this thread holds rows g (c0, c1) and g + 8 (c2, c3), columns 8 j + 2 t (+1) ----
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const bool m0 = msk[j * 8 + 2 * t], m1 = msk[j * 8 + 2 * t + 1];
        s[j][0] = m0 ? -INFINITY : s[j][0] * scale; s[j][1] = m1 ? -INFINITY : s[j][1] * scale;
        s[j][2] = m0 ? -INFINITY : s[j][2] * scale; s[j][3] = m1 ? -INFINITY : s[j][3] * scale;
        mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
        mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        s[j][0] = mx0 == -INFINITY ? 0.f : expf(s[j][0] - mx0); s[j][1] = mx0 == -INFINITY ? 0.f : expf(s[j][1] - mx0);
        s[j][2] = mx1 == -INFINITY ? 0.f : expf(s[j][2] - mx1); s[j][3] = mx1 == -INFINITY ? 0.f : expf(s[j][3] - mx1);
        sum0 += s[j][0] + s[j][1];
        sum1 += s[j][2] + s[j][3];
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    const float inv0 = sum0 > 0.f ? 1.f / sum0 : 0.f, inv1 = sum1 > 0.f ? 1.f / sum1 : 0.f;
    const int i0 = row0 + g, i1 = row0 + g + 8;
    const float invH = 1.f / (float)H;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        s[j][0] *= inv0; s[j][1] *= inv0; s[j][2] *= inv1; s[j][3] *= inv1;
        if (drop.thresh) {
            // train-mode dropout on the normalised probabilities, element (b, h, i, j) of [B, H, Lq, Lk] (the same counter-based
            // mask as every other attention kernel): the head-averaged weights and P V both see the dropped values, like torch
            const int c = j * 8 + 2 * t;
            const uint64_t r0 = (((uint64_t)b * H + h) * Lq + i0) * (uint64_t)Lk, r1 = (((uint64_t)b * H + h) * Lq + i1) * (uint64_t)Lk;
            if (i0 < Lq && c < Lk) s[j][0] = drop_apply(drop, r0 + c, s[j][0]);
            if (i0 < Lq && c + 1 < Lk) s[j][1] = drop_apply(drop, r0 + c + 1, s[j][1]);
            if (i1 < Lq && c < Lk) s[j][2] = drop_apply(drop, r1 + c, s[j][2]);
            if (i1 < Lq && c + 1 < Lk) s[j][3] = drop_apply(drop, r1 + c + 1, s[j][3]);
        }
        if (p_avg) {
            const int c = j * 8 + 2 * t;
            if (i0 < Lq && c < Lk) atomicAdd(p_avg + ((int64_t)b * Lq + i0) * Lk + c, s[j][0] * invH);
            if (i0 < Lq && c + 1 < Lk) atomicAdd(p_avg + ((int64_t)b * Lq + i0) * Lk + c + 1, s[j][1] * invH);
            if (i1 < Lq && c < Lk) atomicAdd(p_avg + ((int64_t)b * Lq + i1) * Lk + c, s[j][2] * invH);
            if (i1 < Lq && c + 1 < Lk) atomicAdd(p_avg + ((int64_t)b * Lq + i1) * Lk + c + 1, s[j][3] * invH);
        }
    }
    if (t == 0) {
        if (i0 < Lq) lse[((int64_t)b * H + h) * Lq + i0] = sum0 > 0.f ? mx0 + logf(sum0) : -INFINITY;
        if (i1 < Lq) lse[((int64_t)b * H + h) * Lq + i1] = sum1 > 0.f ? mx1 + logf(sum1) : -INFINITY;
    }
    // ---- O = P V: P from the accumulator registers (two adjacent key tiles = one 16-key k step) ----
    float oacc[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) { oacc[n][0] = oacc[n][1] = oacc[n][2] = oacc[n][3] = 0.f; }
#pragma unroll
    for (int kt = 0; kt < NT / 2; ++kt) {
        uint32_t a[4] = {pack2(s[2 * kt][0], s[2 * kt][1]), pack2(s[2 * kt][2], s[2 * kt][3]),
                         pack2(s[2 * kt + 1][0], s[2 * kt + 1][1]), pack2(s[2 * kt + 1][2], s[2 * kt + 1][3])};
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            uint32_t bb[2];
            load_b(bb, VsT, PT, n * 8, kt * 16, g, t);  // B[k = key][n = dim] = VsT[dim][key]
            mma_bf16_16816(oacc[n], a, bb);
        }
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        const int c = col + n * 8 + 2 * t;
        if (i0 < Lq) *reinterpret_cast<uint32_t*>(o + ((int64_t)b * Lq + i0) * ldo + c) = pack2(oacc[n][0], oacc[n][1]);
        if (i1 < Lq) *reinterpret_cast<uint32_t*>(o + ((int64_t)b * Lq + i1) * ldo + c) = pack2(oacc[n][2], oacc[n][3]);
    }
}

template <int NT>
__global__ void __launch_bounds__(256)
attn_mma_bwd_kernel(const __nv_bfloat16* __restrict__ q, int64_t ldq, const __nv_bfloat16* __restrict__ k, int64_t ldk,
                    const __nv_bfloat16* __restrict__ v, int64_t ldv, const __nv_bfloat16* __restrict__ d_o, int64_t lddo,
                    const uint8_t* __restrict__ key_mask, const float* __restrict__ lse, const float* __restrict__ dp_avg,
                    __nv_bfloat16* __restrict__ dq, int64_t lddq, __nv_bfloat16* __restrict__ dk, int64_t lddk,
                    __nv_bfloat16* __restrict__ dv, int64_t lddv, int H, int Lq, int Lk, float scale, const DropArgs drop_in) {
    const DropArgs drop = drop_resolve(drop_in);
    extern __shared__ __align__(16) uint8_t mm_smem[];
    pdl_launch_dependents();
    pdl_wait();
    constexpr int LKP = NT * 8;
    const int LQP = (Lq + 15) & ~15;
    const int PQ = LQP + 8;       // pitch of tiles transposed over the query index
    constexpr int PK = LKP + 8;   // pitch of tiles transposed over the key index
    __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(mm_smem);  // [LQP][40]
    __nv_bfloat16* Gs = Qs + LQP * MM_PR;                             // dO [LQP][40]
    __nv_bfloat16* Ks = Gs + LQP * MM_PR;                             // [LKP][40]
    __nv_bfloat16* Vs = Ks + LKP * MM_PR;                             // [LKP][40]
    __nv_bfloat16* QsT = Vs + LKP * MM_PR;                            // [32][PQ]
    __nv_bfloat16* GsT = QsT + MM_DH * PQ;                            // [32][PQ]
    __nv_bfloat16* KsT = GsT + MM_DH * PQ;                            // [32][PK]
    __nv_bfloat16* PT = KsT + MM_DH * PK;                             // P^T  [LKP][PQ]
    __nv_bfloat16* dST = PT + LKP * PQ;                               // dS^T [LKP][PQ]
    __shared__ uint8_t msk[MM_LMAX];
    const int h = blockIdx.x, b = blockIdx.y;
    const int col = h * MM_DH;
    mm_load(Qs, QsT, PQ, q, ldq, (int64_t)b * Lq, col, Lq, LQP);
    mm_load(Gs, GsT, PQ, d_o, lddo, (int64_t)b * Lq, col, Lq, LQP);
    mm_load(Ks, KsT, PK, k, ldk, (int64_t)b * Lk, col, Lk, LKP);
    mm_load(Vs, nullptr, 0, v, ldv, (int64_t)b * Lk, col, Lk, LKP);
    for (int j = threadIdx.x; j < LKP; j += blockDim.x) msk[j] = (j >= Lk) || (key_mask && key_mask[(int64_t)b * Lk + j]);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const float invH = 1.f / (float)H;
    // ================= phase 1: this warp's 16 query rows =================
    {
        const int row0 = warp * 16;
        if (row0 < LQP) {
            float s[NT][4], dp[NT][4];
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
                dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
            }
#pragma unroll
            for (int kk = 0; kk < MM_DH; kk += 16) {
                uint32_t a[4], ag[4];
                load_a(a, Qs, MM_PR, row0, kk, g, t);
                load_a(ag, Gs, MM_PR, row0, kk, g, t);
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    uint32_t bk[2], bv[2];
                    load_b(bk, Ks, MM_PR, j * 8, kk, g, t);
                    load_b(bv, Vs, MM_PR, j * 8, kk, g, t);  // dP = dO V^T: B[k = dim][n = key] = V[key][dim]
                    mma_bf16_16816(s[j], a, bk);
                    mma_bf16_16816(dp[j], ag, bv);
                }
            }
            const int i0 = row0 + g, i1 = row0 + g + 8;
            const float l0 = i0 < Lq ? lse[((int64_t)b * H + h) * Lq + i0] : -INFINITY;
            const float l1 = i1 < Lq ? lse[((int64_t)b * H + h) * Lq + i1] : -INFINITY;
            float D0 = 0.f, D1 = 0.f;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const int c = j * 8 + 2 * t;
                const bool m0 = msk[c], m1 = msk[c + 1];
                const bool d00 = m0 || l0 == -INFINITY, d01 = m1 || l0 == -INFINITY;
                const bool d10 = m0 || l1 == -INFINITY, d11 = m1 || l1 == -INFINITY;
                s[j][0] = d00 ? 0.f : expf(s[j][0] * scale - l0); s[j][1] = d01 ? 0.f : expf(s[j][1] * scale - l0);
                s[j][2] = d10 ? 0.f : expf(s[j][2] * scale - l1); s[j][3] = d11 ? 0.f : expf(s[j][3] * scale - l1);
                if (dp_avg) {
                    if (!d00) dp[j][0] += dp_avg[((int64_t)b * Lq + i0) * Lk + c] * invH;
                    if (!d01) dp[j][1] += dp_avg[((int64_t)b * Lq + i0) * Lk + c + 1] * invH;
                    if (!d10) dp[j][2] += dp_avg[((int64_t)b * Lq + i1) * Lk + c] * invH;
                    if (!d11) dp[j][3] += dp_avg[((int64_t)b * Lq + i1) * Lk + c + 1] * invH;
                }
                if (drop.thresh) {  // o (and the weights output) were formed from the DROPPED probabilities: dP passes the mask
                    const uint64_t r0 = (((uint64_t)b * H + h) * Lq + i0) * (uint64_t)Lk, r1 = (((uint64_t)b * H + h) * Lq + i1) * (uint64_t)Lk;
                    if (!d00) dp[j][0] = drop_apply(drop, r0 + c, dp[j][0]);
                    if (!d01) dp[j][1] = drop_apply(drop, r0 + c + 1, dp[j][1]);
                    if (!d10) dp[j][2] = drop_apply(drop, r1 + c, dp[j][2]);
                    if (!d11) dp[j][3] = drop_apply(drop, r1 + c + 1, dp[j][3]);
                }
                D0 += s[j][0] * dp[j][0] + s[j][1] * dp[j][1];
                D1 += s[j][2] * dp[j][2] + s[j][3] * dp[j][3];
            }
            D0 += __shfl_xor_sync(0xffffffffu, D0, 1); D0 += __shfl_xor_sync(0xffffffffu, D0, 2);
            D1 += __shfl_xor_sync(0xffffffffu, D1, 1); D1 += __shfl_xor_sync(0xffffffffu, D1, 2);
#pragma unroll
            for (int j = 0; j < NT; ++j) {  // dp <- dS; P^T, dS^T to shared memory for phase 2
                dp[j][0] = s[j][0] * (dp[j][0] - D0) * scale; dp[j][1] = s[j][1] * (dp[j][1] - D0) * scale;
                dp[j][2] = s[j][2] * (dp[j][2] - D1) * scale; dp[j][3] = s[j][3] * (dp[j][3] - D1) * scale;
                const int c = j * 8 + 2 * t;
                if (drop.thresh) {  // dV = Pm^T dO with Pm the dropped probabilities (what multiplied V in the forward)
                    const uint64_t r0 = (((uint64_t)b * H + h) * Lq + i0) * (uint64_t)Lk, r1 = (((uint64_t)b * H + h) * Lq + i1) * (uint64_t)Lk;
                    if (s[j][0] != 0.f) s[j][0] = drop_apply(drop, r0 + c, s[j][0]);
                    if (s[j][1] != 0.f) s[j][1] = drop_apply(drop, r0 + c + 1, s[j][1]);
                    if (s[j][2] != 0.f) s[j][2] = drop_apply(drop, r1 + c, s[j][2]);
                    if (s[j][3] != 0.f) s[j][3] = drop_apply(drop, r1 + c + 1, s[j][3]);
                }
                PT[c * PQ + i0] = __float2bfloat16_rn(s[j][0]); PT[(c + 1) * PQ + i0] = __float2bfloat16_rn(s[j][1]);
                PT[c * PQ + i1] = __float2bfloat16_rn(s[j][2]); PT[(c + 1) * PQ + i1] = __float2bfloat16_rn(s[j][3]);
                dST[c * PQ + i0] = __float2bfloat16_rn(dp[j][0]); dST[(c + 1) * PQ + i0] = __float2bfloat16_rn(dp[j][1]);
                dST[c * PQ + i1] = __float2bfloat16_rn(dp[j][2]); dST[(c + 1) * PQ + i1] = __float2bfloat16_rn(dp[j][3]);
            }
            // dQ = dS K (dS from the registers as A operand; B[k = key][n = dim] = KsT[dim][key])
            float acc[4][4];
#pragma unroll
            for (int n = 0; n < 4; ++n) { acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f; }
#pragma unroll
            for (int kt = 0; kt < NT / 2; ++kt) {
                uint32_t a[4] = {pack2(dp[2 * kt][0], dp[2 * kt][1]), pack2(dp[2 * kt][2], dp[2 * kt][3]),
                                 pack2(dp[2 * kt + 1][0], dp[2 * kt + 1][1]), pack2(dp[2 * kt + 1][2], dp[2 * kt + 1][3])};
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    uint32_t bb[2];
                    load_b(bb, KsT, PK, n * 8, kt * 16, g, t);
                    mma_bf16_16816(acc[n], a, bb);
                }
            }
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                const int c = col + n * 8 + 2 * t;
                if (i0 < Lq) *reinterpret_cast<uint32_t*>(dq + ((int64_t)b * Lq + i0) * lddq + c) = pack2(acc[n][0], acc[n][1]);
                if (i1 < Lq) *reinterpret_cast<uint32_t*>(dq + ((int64_t)b * Lq + i1) * lddq + c) = pack2(acc[n][2], acc[n][3]);
            }
        }
    }
    __syncthreads();
    // ================= phase 2: this warp's 16 key rows: dV = P^T dO, dK = dS^T Q =================
    {
        const int key0 = warp * 16;
        if (key0 < LKP) {
            float av[4][4], ak[4][4];
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                av[n][0] = av[n][1] = av[n][2] = av[n][3] = 0.f;
                ak[n][0] = ak[n][1] = ak[n][2] = ak[n][3] = 0.f;
            }
            for (int kq = 0; kq < LQP; kq += 16) {  // contraction over the queries
                uint32_t ap[4], as_[4];
                load_a(ap, PT, PQ, key0, kq, g, t);
                load_a(as_, dST, PQ, key0, kq, g, t);
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    uint32_t bg[2], bq[2];
                    load_b(bg, GsT, PQ, n * 8, kq, g, t);  // B[k = query][n = dim] = dO^T[dim][query]
                    load_b(bq, QsT, PQ, n * 8, kq, g, t);
                    mma_bf16_16816(av[n], ap, bg);
                    mma_bf16_16816(ak[n], as_, bq);
                }
            }
            const int j0 = key0 + g, j1 = key0 + g + 8;
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                const int c = col + n * 8 + 2 * t;
                if (j0 < Lk) {
                    *reinterpret_cast<uint32_t*>(dv + ((int64_t)b * Lk + j0) * lddv + c) = pack2(av[n][0], av[n][1]);
                    *reinterpret_cast<uint32_t*>(dk + ((int64_t)b * Lk + j0) * lddk + c) = pack2(ak[n][0], ak[n][1]);
                }
                if (j1 < Lk) {
                    *reinterpret_cast<uint32_t*>(dv + ((int64_t)b * Lk + j1) * lddv + c) = pack2(av[n][2], av[n][3]);
                    *reinterpret_cast<uint32_t*>(dk + ((int64_t)b * Lk + j1) * lddk + c) = pack2(ak[n][2], ak[n][3]);
                }
            }
        }
    }
}

static size_t mma_fwd_bytes(int Lq, int LKP) {
    const int LQP = (Lq + 15) & ~15;
    return 2 * ((size_t)LQP * MM_PR + 2 * (size_t)LKP * MM_PR + (size_t)MM_DH * (LKP + 8));
}
static size_t mma_bwd_bytes(int Lq, int LKP) {
    const int LQP = (Lq + 15) & ~15;
    return 2 * (2 * (size_t)LQP * MM_PR + 2 * (size_t)LKP * MM_PR + 2 * (size_t)MM_DH * (LQP + 8) + (size_t)MM_DH * (LKP + 8) +
                2 * (size_t)LKP * (LQP + 8));
}

int attn_mma_supported(int dtype, const void* q2, int B, int H, int Lq, int Lk, const void* const* ptrs, const int64_t* lds, int n) {
    if (getenv("STCAT_DISABLE_MMA_ATTN")) return 0;
    if (dtype != STCAT_BF16 || q2 || Lq < 2 || Lq > MM_LMAX || Lk < 1 || Lk > MM_LMAX || B > 65535 || H > 65535) return 0;
    for (int i = 0; i < n; ++i)
        if (ptrs[i] && ((((uintptr_t)ptrs[i]) & 15) || (lds[i] % 8))) return 0;  // 16-byte row segments
    return 1;
}

template <int NT>
static int mma_fwd_launch(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo,
                          const uint8_t* key_mask, float* lse, float* p_avg, int B, int H, int Lq, int Lk, float scale, cudaStream_t st,
                          const DropArgs& drop) {
    const size_t smem = mma_fwd_bytes(Lq, NT * 8);
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(attn_mma_fwd_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mma_fwd_bytes(MM_LMAX, NT * 8));
        if (e != cudaSuccess) return set_err((int)e, "attn_mma_fwd: smem attribute: %s", cudaGetErrorString(e));
        attr = true;
    }
    const int warps = ((Lq + 15) / 16);
    cudaError_t le = launch_pdl(attn_mma_fwd_kernel<NT>, dim3(H, B), dim3(warps * 32), smem, st, (const __nv_bfloat16*)q, ldq,
                                (const __nv_bfloat16*)k, ldk, (const __nv_bfloat16*)v, ldv, (__nv_bfloat16*)o, ldo, key_mask, lse, p_avg, H,
                                Lq, Lk, scale, drop);
    if (le != cudaSuccess) return set_err((int)le, "attn_mma_fwd launch: %s", cudaGetErrorString(le));
    return check_launch("attn_mma_fwd_kernel");
}

template <int NT>
static int mma_bwd_launch(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* d_o,
                          int64_t lddo, const uint8_t* key_mask, const float* lse, const float* dp_avg, void* dq, int64_t lddq,
                          void* dk, int64_t lddk, void* dv, int64_t lddv, int B, int H, int Lq, int Lk, float scale, cudaStream_t st,
                          const DropArgs& drop) {
    const size_t smem = mma_bwd_bytes(Lq, NT * 8);
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(attn_mma_bwd_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mma_bwd_bytes(MM_LMAX, NT * 8));
        if (e != cudaSuccess) return set_err((int)e, "attn_mma_bwd: smem attribute: %s", cudaGetErrorString(e));
        attr = true;
    }
    const int lmax = Lq > Lk ? Lq : Lk;
    const int warps = ((lmax + 15) / 16);
    cudaError_t le = launch_pdl(attn_mma_bwd_kernel<NT>, dim3(H, B), dim3(warps * 32), smem, st, (const __nv_bfloat16*)q, ldq,
                                (const __nv_bfloat16*)k, ldk, (const __nv_bfloat16*)v, ldv, (const __nv_bfloat16*)d_o, lddo, key_mask, lse,
                                dp_avg, (__nv_bfloat16*)dq, lddq, (__nv_bfloat16*)dk, lddk, (__nv_bfloat16*)dv, lddv, H, Lq, Lk, scale, drop);
    if (le != cudaSuccess) return set_err((int)le, "attn_mma_bwd launch: %s", cudaGetErrorString(le));
    return check_launch("attn_mma_bwd_kernel");
}

#define STCAT_MMA_DISPATCH(FN, ...)                    \
    switch ((Lk + 15) / 16) {                          \
        case 1: return FN<2>(__VA_ARGS__);             \
        case 2: return FN<4>(__VA_ARGS__);             \
        case 3: return FN<6>(__VA_ARGS__);             \
        case 4: return FN<8>(__VA_ARGS__);             \
        case 5: return FN<10>(__VA_ARGS__);            \
        case 6: return FN<12>(__VA_ARGS__);            \
        case 7: return FN<14>(__VA_ARGS__);            \
        default: return FN<16>(__VA_ARGS__);           \
    }

int attn_mma_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo,
                 const uint8_t* key_mask, float* lse, float* p_avg, int B, int H, int Lq, int Lk, float scale, cudaStream_t st,
                 const DropArgs& drop) {
    STCAT_MMA_DISPATCH(mma_fwd_launch, q, ldq, k, ldk, v, ldv, o, ldo, key_mask, lse, p_avg, B, H, Lq, Lk, scale, st, drop)
}

int attn_mma_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* d_o, int64_t lddo,
                 const uint8_t* key_mask, const float* lse, const float* dp_avg, void* dq, int64_t lddq, void* dk, int64_t lddk,
                 void* dv, int64_t lddv, int B, int H, int Lq, int Lk, float scale, cudaStream_t st, const DropArgs& drop) {
    STCAT_MMA_DISPATCH(mma_bwd_launch, q, ldq, k, ldk, v, ldv, d_o, lddo, key_mask, lse, dp_avg, dq, lddq, dk, lddk, dv, lddv, B, H,
                       Lq, Lk, scale, st, drop)
}

}  // namespace stcat
