// Short-sequence multi-head attention (Lq, Lk <= 128), forward and backward, one CTA per (batch, head).
//
// Serves the temporal self-attentions of the hot path, whose sequences are the frames of one clip: the encoder's
// temporal layers over T+1 frame-CLS tokens (modal_encoder.py:180-185) and the self-attention over the T queries in
// both decoders (query_decoder.py:341, 604).  These sit on the dependent chain of ~10^3 tiny kernels per step, so
// the design goal is latency: the whole (b, h) problem -- Q, K, V, (dO) and the L x L score matrix -- lives in
// shared memory, every phase is one pass of all 256 threads with 4 x 4 (or 4 x 2) register tiles, and the backward
// is ONE kernel (the generic SIMT path needs two kernels and two passes over the keys: 45 us -> ~5 us at L = 64).
// Exact fp32 arithmetic on fp32 or bf16 operands.  Optional head-averaged probabilities with gradient (the
// `weights` output, query_decoder.py:604-610 -> criterion.py:111-130).
#include "common.cuh"
#include <math.h>

namespace stcat {

constexpr int SM_DH = 32;
constexpr int SM_LMAX = 128;
constexpr int SM_THREADS = 1024;  // 32 warps: the phases are shared-memory-latency bound, so occupancy is what counts
constexpr int SM_LDV = SM_DH + 1;  // padded row of the [L][32] operand tiles

// [L][33] fp32 tile <- rows (base_row .. base_row + L) of a [*, ld] matrix, columns [col0, col0 + 32); rows >= L are zero
template <typename T>
__device__ __forceinline__ void sm_load_tile(float* dst, const T* __restrict__ src, int64_t ld, int64_t base_row, int col0,
                                             int L, int Lpad) {
    for (int i = threadIdx.x; i < Lpad * SM_DH; i += SM_THREADS) {
        const int r = i >> 5, d = i & 31;
        dst[r * SM_LDV + d] = r < L ? to_f32<T>(src[(base_row + r) * ld + col0 + d]) : 0.f;
    }
}

// C[i][j] = sum_e A[i][e] * B[j][e] over the 32-wide head dim for all i < LqPad, j < LkPad; 2 x 2 register tiles,
// one (or a few) per thread
template <typename F>
__device__ __forceinline__ void sm_outer_32(const float* A, const float* B, int LqPad, int LkPad, F&& store) {
    const int tq = LqPad >> 1, tk = LkPad >> 1;
    for (int t = threadIdx.x; t < tq * tk; t += SM_THREADS) {
        // rows i0, i0+1 (shared by most of a warp: broadcast), columns jl and jl + tk (consecutive lanes ->
        // consecutive rows of B, pitch 33 words: conflict-free)
        const int i0 = (t / tk) * 2, jl = t % tk;
        const float* a0 = A + i0 * SM_LDV;
        const float* b0 = B + jl * SM_LDV;
        const float* b1 = B + (jl + tk) * SM_LDV;
        float c00 = 0.f, c01 = 0.f, c10 = 0.f, c11 = 0.f;
#pragma unroll
        for (int e = 0; e < SM_DH; ++e) {
            const float x0 = a0[e], x1 = a0[SM_LDV + e], y0 = b0[e], y1 = b1[e];
            c00 = fmaf(x0, y0, c00); c01 = fmaf(x0, y1, c01);
            c10 = fmaf(x1, y0, c10); c11 = fmaf(x1, y1, c11);
        }
        store(i0, jl, c00); store(i0, jl + tk, c01);
        store(i0 + 1, jl, c10); store(i0 + 1, jl + tk, c11);
    }
}

// O[r][e] = sum_c W[r][c] * X[c][e] (TRANS = false, W row-major [R][ldw]) or sum_c W[c][r] * X[c][e] (TRANS = true);
// R x 32 outputs, 1 x 2 per thread; C = contraction length
template <bool TRANS, typename F>
__device__ __forceinline__ void sm_apply(const float* W, int ldw, const float* X, int Rpad, int C, F&& store) {
    for (int t = threadIdx.x; t < Rpad * (SM_DH / 2); t += SM_THREADS) {
        const int r = t / (SM_DH / 2), e0 = (t % (SM_DH / 2)) * 2;
        const float* wp = TRANS ? W + r : W + r * ldw;
        const int ws = TRANS ? ldw : 1;
        const float* xp = X + e0;
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
        int c = 0;
        for (; c + 1 < C; c += 2) {  // two independent accumulator pairs
            const float w0 = wp[c * ws], w1 = wp[(c + 1) * ws];
            acc0 = fmaf(w0, xp[c * SM_LDV], acc0);
            acc1 = fmaf(w0, xp[c * SM_LDV + 1], acc1);
            acc2 = fmaf(w1, xp[(c + 1) * SM_LDV], acc2);
            acc3 = fmaf(w1, xp[(c + 1) * SM_LDV + 1], acc3);
        }
        if (c < C) {
            const float w0 = wp[c * ws];
            acc0 = fmaf(w0, xp[c * SM_LDV], acc0);
            acc1 = fmaf(w0, xp[c * SM_LDV + 1], acc1);
        }
        store(r, e0, acc0 + acc2);
        store(r, e0 + 1, acc1 + acc3);
    }
}

template <typename T>
__global__ void __launch_bounds__(SM_THREADS)
attn_small_fwd_kernel(const T* __restrict__ q, int64_t ldq, const T* __restrict__ k, int64_t ldk, const T* __restrict__ v,
                      int64_t ldv, T* __restrict__ o, int64_t ldo, const uint8_t* __restrict__ key_mask,
                      float* __restrict__ lse, float* __restrict__ p_avg, int H, int Lq, int Lk, float scale) {
    extern __shared__ float smf[];
    pdl_launch_dependents();
    pdl_wait();
    const int LqP = (Lq + 3) & ~3, LkP = (Lk + 3) & ~3;
    const int lds = LkP + 1;
    float* Qs = smf;
    float* Ks = Qs + LqP * SM_LDV;
    float* Vs = Ks + LkP * SM_LDV;
    float* S = Vs + LkP * SM_LDV;  // [LqP][lds]
    const int h = blockIdx.x, b = blockIdx.y;
    const int col = h * SM_DH;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    sm_load_tile<T>(Qs, q, ldq, (int64_t)b * Lq, col, Lq, LqP);
    sm_load_tile<T>(Ks, k, ldk, (int64_t)b * Lk, col, Lk, LkP);
    sm_load_tile<T>(Vs, v, ldv, (int64_t)b * Lk, col, Lk, LkP);
    __syncthreads();
    sm_outer_32(Qs, Ks, LqP, LkP, [&](int i, int j, float a) { S[i * lds + j] = a * scale; });
    __syncthreads();
    // softmax per row: one warp per row
    const float invH = 1.f / (float)H;
    for (int i = warp; i < Lq; i += SM_THREADS / 32) {
        float* row = S + i * lds;
        float mx = -INFINITY;
        for (int j = lane; j < Lk; j += 32) {
            const bool masked = key_mask && key_mask[(int64_t)b * Lk + j];
            const float s = masked ? -INFINITY : row[j];
            row[j] = s;
            mx = fmaxf(mx, s);
        }
        mx = warp_max(mx);
        float sum = 0.f;
        for (int j = lane; j < Lk; j += 32) {
            const float p = (mx == -INFINITY) ? 0.f : expf(row[j] - mx);
            row[j] = p;
            sum += p;
        }
        sum = warp_sum(sum);
        const float inv = sum > 0.f ? 1.f / sum : 0.f;
        for (int j = lane; j < Lk; j += 32) {
            const float p = row[j] * inv;
            row[j] = p;
            if (p_avg) atomicAdd(p_avg + ((int64_t)b * Lq + i) * Lk + j, p * invH);
        }
        if (lane == 0) lse[((int64_t)b * H + h) * Lq + i] = sum > 0.f ? mx + logf(sum) : -INFINITY;
    }
    __syncthreads();
    sm_apply<false>(S, lds, Vs, LqP, Lk, [&](int i, int e, float a) {
        if (i < Lq) o[((int64_t)b * Lq + i) * ldo + col + e] = from_f32<T>(a);
    });
}

template <typename T>
__global__ void __launch_bounds__(SM_THREADS)
attn_small_bwd_kernel(const T* __restrict__ q, int64_t ldq, const T* __restrict__ k, int64_t ldk, const T* __restrict__ v,
                      int64_t ldv, const T* __restrict__ d_o, int64_t lddo, const uint8_t* __restrict__ key_mask,
                      const float* __restrict__ lse, const float* __restrict__ dp_avg, T* __restrict__ dq, int64_t lddq,
                      T* __restrict__ dk, int64_t lddk, T* __restrict__ dv, int64_t lddv, int H, int Lq, int Lk,
                      float scale) {
    extern __shared__ float smf[];
    pdl_launch_dependents();
    pdl_wait();
    const int LqP = (Lq + 3) & ~3, LkP = (Lk + 3) & ~3;
    const int lds = LkP + 1;
    float* Qs = smf;
    float* Ks = Qs + LqP * SM_LDV;
    float* Vs = Ks + LkP * SM_LDV;
    float* Gs = Vs + LkP * SM_LDV;   // dO [LqP][33]
    float* P = Gs + LqP * SM_LDV;    // [LqP][lds] probabilities
    float* dS = P + LqP * lds;       // [LqP][lds] dP, then dS
    const int h = blockIdx.x, b = blockIdx.y;
    const int col = h * SM_DH;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    sm_load_tile<T>(Qs, q, ldq, (int64_t)b * Lq, col, Lq, LqP);
    sm_load_tile<T>(Ks, k, ldk, (int64_t)b * Lk, col, Lk, LkP);
    sm_load_tile<T>(Vs, v, ldv, (int64_t)b * Lk, col, Lk, LkP);
    sm_load_tile<T>(Gs, d_o, lddo, (int64_t)b * Lq, col, Lq, LqP);
    __syncthreads();
    sm_outer_32(Qs, Ks, LqP, LkP, [&](int i, int j, float a) { P[i * lds + j] = a * scale; });
    sm_outer_32(Gs, Vs, LqP, LkP, [&](int i, int j, float a) { dS[i * lds + j] = a; });
    __syncthreads();
    const float invH = 1.f / (float)H;
    for (int i = warp; i < LqP; i += SM_THREADS / 32) {
        float* prow = P + i * lds;
        float* drow = dS + i * lds;
        const float l = i < Lq ? lse[((int64_t)b * H + h) * Lq + i] : -INFINITY;
        float D = 0.f;
        for (int j = lane; j < LkP; j += 32) {
            const bool dead = i >= Lq || j >= Lk || l == -INFINITY || (key_mask && key_mask[(int64_t)b * Lk + j]);
            const float p = dead ? 0.f : expf(prow[j] - l);
            float dp = dead ? 0.f : drow[j];
            if (!dead && dp_avg) dp += dp_avg[((int64_t)b * Lq + i) * Lk + j] * invH;
            prow[j] = p;
            drow[j] = dp;
            D += p * dp;
        }
        D = warp_sum(D);
        for (int j = lane; j < LkP; j += 32) drow[j] = prow[j] * (drow[j] - D) * scale;
    }
    __syncthreads();
    // dV = P^T dO, dK = dS^T Q, dQ = dS K
    sm_apply<true>(P, lds, Gs, LkP, Lq, [&](int j, int e, float a) {
        if (j < Lk) dv[((int64_t)b * Lk + j) * lddv + col + e] = from_f32<T>(a);
    });
    sm_apply<true>(dS, lds, Qs, LkP, Lq, [&](int j, int e, float a) {
        if (j < Lk) dk[((int64_t)b * Lk + j) * lddk + col + e] = from_f32<T>(a);
    });
    sm_apply<false>(dS, lds, Ks, LqP, Lk, [&](int i, int e, float a) {
        if (i < Lq) dq[((int64_t)b * Lq + i) * lddq + col + e] = from_f32<T>(a);
    });
}

static size_t sm_fwd_bytes(int Lq, int Lk) {
    const int LqP = (Lq + 3) & ~3, LkP = (Lk + 3) & ~3;
    return sizeof(float) * ((size_t)(LqP + 2 * LkP) * SM_LDV + (size_t)LqP * (LkP + 1));
}
static size_t sm_bwd_bytes(int Lq, int Lk) {
    const int LqP = (Lq + 3) & ~3, LkP = (Lk + 3) & ~3;
    return sizeof(float) * ((size_t)(2 * LqP + 2 * LkP) * SM_LDV + 2 * (size_t)LqP * (LkP + 1));
}

int attn_small_supported(const void* q2, int B, int H, int Lq, int Lk) {
    if (getenv("STCAT_DISABLE_SMALL_ATTN")) return 0;
    return !q2 && Lq >= 2 && Lq <= SM_LMAX && Lk >= 1 && Lk <= SM_LMAX && B <= 65535 && H <= 65535;
}

template <typename T>
static int sm_launch_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                         int64_t ldo, const uint8_t* key_mask, float* lse, float* p_avg, int B, int H, int Lq, int Lk,
                         float scale, cudaStream_t st) {
    const size_t smem = sm_fwd_bytes(Lq, Lk);
    static size_t set_to = 0;
    if (smem > 48 * 1024 && smem > set_to) {
        cudaError_t e = cudaFuncSetAttribute(attn_small_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_fwd_bytes(SM_LMAX, SM_LMAX));
        if (e != cudaSuccess) return set_err((int)e, "attn_small_fwd: smem attribute: %s", cudaGetErrorString(e));
        set_to = sm_fwd_bytes(SM_LMAX, SM_LMAX);
    }
    launch_pdl(attn_small_fwd_kernel<T>, dim3(H, B), dim3(SM_THREADS), smem, st, (const T*)q, ldq, (const T*)k, ldk, (const T*)v, ldv, (T*)o,
               ldo, key_mask, lse, p_avg, H, Lq, Lk, scale);
    return check_launch("attn_small_fwd_kernel");
}

template <typename T>
static int sm_launch_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* d_o,
                         int64_t lddo, const uint8_t* key_mask, const float* lse, const float* dp_avg, void* dq,
                         int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int B, int H, int Lq, int Lk,
                         float scale, cudaStream_t st) {
    const size_t smem = sm_bwd_bytes(Lq, Lk);
    static size_t set_to = 0;
    if (smem > 48 * 1024 && smem > set_to) {
        cudaError_t e = cudaFuncSetAttribute(attn_small_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_bwd_bytes(SM_LMAX, SM_LMAX));
        if (e != cudaSuccess) return set_err((int)e, "attn_small_bwd: smem attribute: %s", cudaGetErrorString(e));
        set_to = sm_bwd_bytes(SM_LMAX, SM_LMAX);
    }
    launch_pdl(attn_small_bwd_kernel<T>, dim3(H, B), dim3(SM_THREADS), smem, st, (const T*)q, ldq, (const T*)k, ldk, (const T*)v, ldv,
               (const T*)d_o, lddo, key_mask, lse, dp_avg, (T*)dq, lddq, (T*)dk, lddk, (T*)dv, lddv, H, Lq, Lk, scale);
    return check_launch("attn_small_bwd_kernel");
}

int attn_small_fwd(int dtype, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                   int64_t ldo, const uint8_t* key_mask, float* lse, float* p_avg, int B, int H, int Lq, int Lk, float scale,
                   cudaStream_t st) {
    if (dtype == STCAT_F32)
        return sm_launch_fwd<float>(q, ldq, k, ldk, v, ldv, o, ldo, key_mask, lse, p_avg, B, H, Lq, Lk, scale, st);
    return sm_launch_fwd<__nv_bfloat16>(q, ldq, k, ldk, v, ldv, o, ldo, key_mask, lse, p_avg, B, H, Lq, Lk, scale, st);
}

int attn_small_bwd(int dtype, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                   const void* d_o, int64_t lddo, const uint8_t* key_mask, const float* lse, const float* dp_avg, void* dq,
                   int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int B, int H, int Lq, int Lk, float scale,
                   cudaStream_t st) {
    if (dtype == STCAT_F32)
        return sm_launch_bwd<float>(q, ldq, k, ldk, v, ldv, d_o, lddo, key_mask, lse, dp_avg, dq, lddq, dk, lddk, dv, lddv, B, H,
                                    Lq, Lk, scale, st);
    return sm_launch_bwd<__nv_bfloat16>(q, ldq, k, ldk, v, ldv, d_o, lddo, key_mask, lse, dp_avg, dq, lddq, dk, lddk, dv, lddv,
                                        B, H, Lq, Lk, scale, st);
}

}  // namespace stcat
