// Exact-fp32 SIMT multi-head attention core (forward + backward), batch-major, head dim 32 per part.
//
// Correctness path and on-device checker for the tcgen05 attention kernel.  Generic in Lq / Lk (online
// softmax over 64-key chunks), optional second score part (per-head concat of content and positional
// halves, reference query_decoder.py:368-384), optional key-padding mask, optional head-averaged
// probability output with its gradient (reference query_decoder.py:604-610 -> criterion.py:111-130).
//
// Work split: one warp owns R query rows (fwd, bwd_dq) or R key rows (bwd_dkv); inside a chunk a lane
// owns one key (resp. query) for the dot products and one feature dim for the accumulations.
#include "common.cuh"
#include <math.h>

namespace stcat {

constexpr int DH = 32;       // head dim per part, and value head dim
constexpr int CH = 64;       // chunk of keys (fwd/dq) or queries (dkv) staged in shared memory
constexpr int WARPS = 4;
constexpr int R = 4;         // rows per warp
constexpr int TILE = WARPS * R;

template <typename T>
__device__ __forceinline__ void stage_chunk(float (*dst)[DH + 1], const T* __restrict__ src, int64_t ld, int row0,
                                            int nrows_total, int64_t base_row, int col0) {
    // 64 rows x 32 dims, 128 threads -> 16 elements each; a warp reads one 32-wide row segment
    for (int i = threadIdx.x; i < CH * DH; i += WARPS * 32) {
        int r = i >> 5, d = i & 31;
        int gr = row0 + r;
        float v = 0.f;
        if (gr < nrows_total) v = to_f32<T>(src[(base_row + gr) * ld + col0 + d]);
        dst[r][d] = v;
    }
}

template <typename T, bool TWO>
__global__ void __launch_bounds__(WARPS * 32)
attn_fwd_kernel(const T* __restrict__ q1, const T* __restrict__ q2, int64_t ldq, const T* __restrict__ k1,
                const T* __restrict__ k2, int64_t ldk, const T* __restrict__ v, int64_t ldv, T* __restrict__ o,
                int64_t ldo, const uint8_t* __restrict__ key_mask, float* __restrict__ lse, float* __restrict__ p_avg,
                int H, int Lq, int Lk, float scale, const DropArgs drop_in) {
    const DropArgs drop = drop_resolve(drop_in);
    __shared__ float Ks1[CH][DH + 1];
    __shared__ float Ks2[TWO ? CH : 1][DH + 1];
    __shared__ float Vs[CH][DH + 1];
    __shared__ float qs[WARPS][R][2 * DH];
    __shared__ uint8_t msk[CH];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.z, h = blockIdx.y;
    const int i0 = blockIdx.x * TILE + warp * R;
    const int col = h * DH;
    const int64_t qbase = (int64_t)b * Lq, kbase = (int64_t)b * Lk;

    float m[R], l[R], acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        m[r] = -INFINITY; l[r] = 0.f; acc[r] = 0.f;
        int i = i0 + r;
        float a = 0.f, c = 0.f;
        if (i < Lq) {
            a = to_f32<T>(q1[(qbase + i) * ldq + col + lane]);
            if (TWO) c = to_f32<T>(q2[(qbase + i) * ldq + col + lane]);
        }
        qs[warp][r][lane] = a;
        qs[warp][r][DH + lane] = c;
    }
    const int npass = p_avg ? 2 : 1;
    float lse_r[R];
    for (int pass = 0; pass < npass; ++pass) {
        for (int c0 = 0; c0 < Lk; c0 += CH) {
            __syncthreads();
            stage_chunk<T>(Ks1, k1, ldk, c0, Lk, kbase, col);
            if (TWO) stage_chunk<T>(Ks2, k2, ldk, c0, Lk, kbase, col);
            if (pass == 0) stage_chunk<T>(Vs, v, ldv, c0, Lk, kbase, col);
            if (threadIdx.x < CH) {
                int j = c0 + threadIdx.x;
                msk[threadIdx.x] = (j >= Lk) || (key_mask && key_mask[(int64_t)b * Lk + j]);
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = i0 + r;
                if (i >= Lq) continue;  // warp-uniform
                float s[2];
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    const int key = lane + 32 * kk;
                    float d = 0.f;
#pragma unroll
                    for (int e = 0; e < DH; ++e) d = fmaf(qs[warp][r][e], Ks1[key][e], d);
                    if (TWO) {
#pragma unroll
                        for (int e = 0; e < DH; ++e) d = fmaf(qs[warp][r][DH + e], Ks2[key][e], d);
                    }
                    s[kk] = msk[key] ? -INFINITY : d * scale;
                }
                if (pass == 0) {
                    float mx = warp_max(fmaxf(s[0], s[1]));
                    float mnew = fmaxf(m[r], mx);
                    float p0 = 0.f, p1 = 0.f, corr = 0.f;
                    if (mnew != -INFINITY) {
                        p0 = expf(s[0] - mnew);
                        p1 = expf(s[1] - mnew);
                        corr = expf(m[r] - mnew);
                    }
                    l[r] = l[r] * corr + warp_sum(p0 + p1);
                    if (drop.thresh) {  // dropout on the (normalised) probabilities: the row sum above stays undropped
                        const uint64_t rowi = (((uint64_t)b * H + h) * Lq + i) * (uint64_t)Lk + c0;
                        p0 = drop_apply(drop, rowi + lane, p0);
                        p1 = drop_apply(drop, rowi + lane + 32, p1);
                    }
                    float a = acc[r] * corr;
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) a = fmaf(__shfl_sync(0xffffffffu, p0, jj), Vs[jj][lane], a);
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) a = fmaf(__shfl_sync(0xffffffffu, p1, jj), Vs[32 + jj][lane], a);
                    acc[r] = a;
                    m[r] = mnew;
                } else {
                    const float invH = 1.f / (float)H;
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const int j = c0 + lane + 32 * kk;
                        if (j < Lk) {
                            float p = (s[kk] == -INFINITY) ? 0.f : expf(s[kk] - lse_r[r]);
                            if (drop.thresh) p = drop_apply(drop, (((uint64_t)b * H + h) * Lq + i) * (uint64_t)Lk + j, p);
                            atomicAdd(p_avg + ((int64_t)b * Lq + i) * Lk + j, p * invH);
                        }
                    }
                }
            }
        }
        if (pass == 0) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = i0 + r;
                lse_r[r] = (l[r] > 0.f) ? m[r] + logf(l[r]) : -INFINITY;
                if (i < Lq) {
                    float inv = (l[r] > 0.f) ? 1.f / l[r] : 0.f;
                    o[(qbase + i) * ldo + col + lane] = from_f32<T>(acc[r] * inv);
                    if (lane == 0) lse[((int64_t)b * H + h) * Lq + i] = lse_r[r];
                }
            }
        }
    }
}

// dq for R rows per warp; also writes delta[b,h,i] = sum_j p_ij * dp_ij
template <typename T, bool TWO>
__global__ void __launch_bounds__(WARPS * 32)
attn_bwd_dq_kernel(const T* __restrict__ q1, const T* __restrict__ q2, int64_t ldq, const T* __restrict__ k1,
                   const T* __restrict__ k2, int64_t ldk, const T* __restrict__ v, int64_t ldv,
                   const T* __restrict__ d_o, int64_t lddo, const uint8_t* __restrict__ key_mask,
                   const float* __restrict__ lse, const float* __restrict__ dp_avg, float* __restrict__ delta,
                   T* __restrict__ dq1, T* __restrict__ dq2, int64_t lddq, int H, int Lq, int Lk, float scale,
                   const DropArgs drop_in) {
    const DropArgs drop = drop_resolve(drop_in);
    __shared__ float Ks1[CH][DH + 1];
    __shared__ float Ks2[TWO ? CH : 1][DH + 1];
    __shared__ float Vs[CH][DH + 1];
    __shared__ float qs[WARPS][R][2 * DH];
    __shared__ float dos[WARPS][R][DH];
    __shared__ uint8_t msk[CH];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.z, h = blockIdx.y;
    const int i0 = blockIdx.x * TILE + warp * R;
    const int col = h * DH;
    const int64_t qbase = (int64_t)b * Lq, kbase = (int64_t)b * Lk;
    const float invH = 1.f / (float)H;

    float lse_r[R], D[R], a1[R], a2[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int i = i0 + r;
        float a = 0.f, c = 0.f, g = 0.f;
        lse_r[r] = 0.f;
        if (i < Lq) {
            a = to_f32<T>(q1[(qbase + i) * ldq + col + lane]);
            if (TWO) c = to_f32<T>(q2[(qbase + i) * ldq + col + lane]);
            g = to_f32<T>(d_o[(qbase + i) * lddo + col + lane]);
            lse_r[r] = lse[((int64_t)b * H + h) * Lq + i];
        }
        qs[warp][r][lane] = a;
        qs[warp][r][DH + lane] = c;
        dos[warp][r][lane] = g;
        D[r] = 0.f; a1[r] = 0.f; a2[r] = 0.f;
    }
    for (int pass = 0; pass < 2; ++pass) {
        for (int c0 = 0; c0 < Lk; c0 += CH) {
            __syncthreads();
            stage_chunk<T>(Ks1, k1, ldk, c0, Lk, kbase, col);
            if (TWO) stage_chunk<T>(Ks2, k2, ldk, c0, Lk, kbase, col);
            stage_chunk<T>(Vs, v, ldv, c0, Lk, kbase, col);
            if (threadIdx.x < CH) {
                int j = c0 + threadIdx.x;
                msk[threadIdx.x] = (j >= Lk) || (key_mask && key_mask[(int64_t)b * Lk + j]);
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = i0 + r;
                if (i >= Lq) continue;
                float p[2], dp[2];
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    const int key = lane + 32 * kk;
                    float d = 0.f, e2 = 0.f;
#pragma unroll
                    for (int e = 0; e < DH; ++e) {
                        d = fmaf(qs[warp][r][e], Ks1[key][e], d);
                        e2 = fmaf(dos[warp][r][e], Vs[key][e], e2);
                    }
                    if (TWO) {
#pragma unroll
                        for (int e = 0; e < DH; ++e) d = fmaf(qs[warp][r][DH + e], Ks2[key][e], d);
                    }
                    const int j = c0 + key;
                    if (msk[key] || lse_r[r] == -INFINITY) { p[kk] = 0.f; dp[kk] = 0.f; }
                    else {
                        p[kk] = expf(d * scale - lse_r[r]);
                        dp[kk] = e2 + (dp_avg ? dp_avg[((int64_t)b * Lq + i) * Lk + j] * invH : 0.f);
                        // o and p_avg were formed from the DROPPED probabilities: their gradient reaches p through the mask
                        if (drop.thresh) dp[kk] = drop_apply(drop, (((uint64_t)b * H + h) * Lq + i) * (uint64_t)Lk + j, dp[kk]);
                    }
                }
                if (pass == 0) {
                    D[r] += p[0] * dp[0] + p[1] * dp[1];
                } else {
                    float ds0 = p[0] * (dp[0] - D[r]) * scale;
                    float ds1 = p[1] * (dp[1] - D[r]) * scale;
                    float x1 = a1[r], x2 = a2[r];
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) {
                        float w0 = __shfl_sync(0xffffffffu, ds0, jj), w1 = __shfl_sync(0xffffffffu, ds1, jj);
                        x1 = fmaf(w0, Ks1[jj][lane], x1);
                        x1 = fmaf(w1, Ks1[32 + jj][lane], x1);
                        if (TWO) {
                            x2 = fmaf(w0, Ks2[jj][lane], x2);
                            x2 = fmaf(w1, Ks2[32 + jj][lane], x2);
                        }
                    }
                    a1[r] = x1; a2[r] = x2;
                }
            }
        }
        if (pass == 0) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                D[r] = warp_sum(D[r]);
                int i = i0 + r;
                if (i < Lq && lane == 0) delta[((int64_t)b * H + h) * Lq + i] = D[r];
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int i = i0 + r;
        if (i < Lq) {
            dq1[(qbase + i) * lddq + col + lane] = from_f32<T>(a1[r]);
            if (TWO) dq2[(qbase + i) * lddq + col + lane] = from_f32<T>(a2[r]);
        }
    }
}

// dk, dv for R key rows per warp
template <typename T, bool TWO>
__global__ void __launch_bounds__(WARPS * 32)
attn_bwd_dkv_kernel(const T* __restrict__ q1, const T* __restrict__ q2, int64_t ldq, const T* __restrict__ k1,
                    const T* __restrict__ k2, int64_t ldk, const T* __restrict__ v, int64_t ldv,
                    const T* __restrict__ d_o, int64_t lddo, const uint8_t* __restrict__ key_mask,
                    const float* __restrict__ lse, const float* __restrict__ dp_avg, const float* __restrict__ delta,
                    T* __restrict__ dk1, T* __restrict__ dk2, int64_t lddk, T* __restrict__ dv, int64_t lddv, int H,
                    int Lq, int Lk, float scale, const DropArgs drop_in) {
    const DropArgs drop = drop_resolve(drop_in);
    __shared__ float Qs1[CH][DH + 1];
    __shared__ float Qs2[TWO ? CH : 1][DH + 1];
    __shared__ float dOs[CH][DH + 1];
    __shared__ float ks[WARPS][R][2 * DH];
    __shared__ float vs[WARPS][R][DH];
    __shared__ float lses[CH], dels[CH];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.z, h = blockIdx.y;
    const int j0 = blockIdx.x * TILE + warp * R;
    const int col = h * DH;
    const int64_t qbase = (int64_t)b * Lq, kbase = (int64_t)b * Lk;
    const float invH = 1.f / (float)H;

    float ak1[R], ak2[R], av[R];
    bool masked[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int j = j0 + r;
        float a = 0.f, c = 0.f, g = 0.f;
        masked[r] = true;
        if (j < Lk) {
            a = to_f32<T>(k1[(kbase + j) * ldk + col + lane]);
            if (TWO) c = to_f32<T>(k2[(kbase + j) * ldk + col + lane]);
            g = to_f32<T>(v[(kbase + j) * ldv + col + lane]);
            masked[r] = key_mask && key_mask[(int64_t)b * Lk + j];
        }
        ks[warp][r][lane] = a;
        ks[warp][r][DH + lane] = c;
        vs[warp][r][lane] = g;
        ak1[r] = 0.f; ak2[r] = 0.f; av[r] = 0.f;
    }
    for (int c0 = 0; c0 < Lq; c0 += CH) {
        __syncthreads();
        stage_chunk<T>(Qs1, q1, ldq, c0, Lq, qbase, col);
        if (TWO) stage_chunk<T>(Qs2, q2, ldq, c0, Lq, qbase, col);
        stage_chunk<T>(dOs, d_o, lddo, c0, Lq, qbase, col);
        if (threadIdx.x < CH) {
            int i = c0 + threadIdx.x;
            lses[threadIdx.x] = (i < Lq) ? lse[((int64_t)b * H + h) * Lq + i] : -INFINITY;
            dels[threadIdx.x] = (i < Lq) ? delta[((int64_t)b * H + h) * Lq + i] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int j = j0 + r;
            if (j >= Lk || masked[r]) continue;  // warp-uniform
            float p[2], ds[2], pm[2];  // pm = dropped probabilities (what multiplied V in the forward)
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                const int qi = lane + 32 * kk;
                const int i = c0 + qi;
                float d = 0.f, e2 = 0.f;
#pragma unroll
                for (int e = 0; e < DH; ++e) {
                    d = fmaf(Qs1[qi][e], ks[warp][r][e], d);
                    e2 = fmaf(dOs[qi][e], vs[warp][r][e], e2);
                }
                if (TWO) {
#pragma unroll
                    for (int e = 0; e < DH; ++e) d = fmaf(Qs2[qi][e], ks[warp][r][DH + e], d);
                }
                if (i < Lq && lses[qi] != -INFINITY) {
                    p[kk] = expf(d * scale - lses[qi]);
                    float dp = e2 + (dp_avg ? dp_avg[((int64_t)b * Lq + i) * Lk + j] * invH : 0.f);
                    pm[kk] = p[kk];
                    if (drop.thresh) {
                        const uint64_t idx = (((uint64_t)b * H + h) * Lq + i) * (uint64_t)Lk + j;
                        dp = drop_apply(drop, idx, dp);
                        pm[kk] = drop_apply(drop, idx, p[kk]);
                    }
                    ds[kk] = p[kk] * (dp - dels[qi]) * scale;
                } else { p[kk] = 0.f; ds[kk] = 0.f; pm[kk] = 0.f; }
            }
            float x1 = ak1[r], x2 = ak2[r], xv = av[r];
#pragma unroll
            for (int ii = 0; ii < 32; ++ii) {
                float p0 = __shfl_sync(0xffffffffu, pm[0], ii), p1 = __shfl_sync(0xffffffffu, pm[1], ii);
                float s0 = __shfl_sync(0xffffffffu, ds[0], ii), s1 = __shfl_sync(0xffffffffu, ds[1], ii);
                xv = fmaf(p0, dOs[ii][lane], xv);
                xv = fmaf(p1, dOs[32 + ii][lane], xv);
                x1 = fmaf(s0, Qs1[ii][lane], x1);
                x1 = fmaf(s1, Qs1[32 + ii][lane], x1);
                if (TWO) {
                    x2 = fmaf(s0, Qs2[ii][lane], x2);
                    x2 = fmaf(s1, Qs2[32 + ii][lane], x2);
                }
            }
            ak1[r] = x1; ak2[r] = x2; av[r] = xv;
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int j = j0 + r;
        if (j < Lk) {
            dk1[(kbase + j) * lddk + col + lane] = from_f32<T>(ak1[r]);
            if (TWO) dk2[(kbase + j) * lddk + col + lane] = from_f32<T>(ak2[r]);
            dv[(kbase + j) * lddv + col + lane] = from_f32<T>(av[r]);
        }
    }
}

template <typename T, bool TWO>
static int launch_fwd(const void* q1, const void* q2, int64_t ldq, const void* k1, const void* k2, int64_t ldk,
                      const void* v, int64_t ldv, void* o, int64_t ldo, const uint8_t* key_mask, float* lse,
                      float* p_avg, int B, int H, int Lq, int Lk, float scale, cudaStream_t st, const DropArgs drop = DropArgs()) {
    dim3 grid((Lq + TILE - 1) / TILE, H, B);
    attn_fwd_kernel<T, TWO><<<grid, WARPS * 32, 0, st>>>((const T*)q1, (const T*)q2, ldq, (const T*)k1, (const T*)k2,
                                                         ldk, (const T*)v, ldv, (T*)o, ldo, key_mask, lse, p_avg, H,
                                                         Lq, Lk, scale, drop);
    return check_launch("attn_fwd_kernel");
}

template <typename T, bool TWO>
static int launch_bwd(const void* q1, const void* q2, int64_t ldq, const void* k1, const void* k2, int64_t ldk,
                      const void* v, int64_t ldv, const void* d_o, int64_t lddo, const uint8_t* key_mask,
                      const float* lse, const float* dp_avg, float* delta, void* dq1, void* dq2, int64_t lddq,
                      void* dk1, void* dk2, int64_t lddk, void* dv, int64_t lddv, int B, int H, int Lq, int Lk,
                      float scale, cudaStream_t st, const DropArgs drop = DropArgs()) {
    dim3 gq((Lq + TILE - 1) / TILE, H, B);
    attn_bwd_dq_kernel<T, TWO><<<gq, WARPS * 32, 0, st>>>((const T*)q1, (const T*)q2, ldq, (const T*)k1, (const T*)k2,
                                                          ldk, (const T*)v, ldv, (const T*)d_o, lddo, key_mask, lse,
                                                          dp_avg, delta, (T*)dq1, (T*)dq2, lddq, H, Lq, Lk, scale, drop);
    int rc = check_launch("attn_bwd_dq_kernel");
    if (rc) return rc;
    dim3 gk((Lk + TILE - 1) / TILE, H, B);
    attn_bwd_dkv_kernel<T, TWO><<<gk, WARPS * 32, 0, st>>>((const T*)q1, (const T*)q2, ldq, (const T*)k1,
                                                           (const T*)k2, ldk, (const T*)v, ldv, (const T*)d_o, lddo,
                                                           key_mask, lse, dp_avg, delta, (T*)dk1, (T*)dk2, lddk,
                                                           (T*)dv, lddv, H, Lq, Lk, scale, drop);
    return check_launch("attn_bwd_dkv_kernel");
}

// attention_tc.cu: tcgen05 kernel for the spatial encoder's shape class
int attn_tc_fwd_supported(int dtype, const void* q2, const void* p_avg, int B, int H, int Lq, int Lk, const void* q,
                          const void* k, const void* v, const void* o, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo);
int attn_tc_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo,
                const uint8_t* key_mask, float* lse, int B, int H, int S, float scale, cudaStream_t st, const DropArgs& drop);

// attention_small.cu: whole-problem-in-shared-memory kernels for short sequences (temporal self-attention)
int attn_small_supported(const void* q2, int B, int H, int Lq, int Lk);
int attn_small_fwd(int dtype, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                   int64_t ldo, const uint8_t* key_mask, float* lse, float* p_avg, int B, int H, int Lq, int Lk, float scale,
                   cudaStream_t st);
int attn_small_bwd(int dtype, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                   const void* d_o, int64_t lddo, const uint8_t* key_mask, const float* lse, const float* dp_avg, void* dq,
                   int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int B, int H, int Lq, int Lk, float scale,
                   cudaStream_t st);

// attention_small_mma.cu: the same short sequences for bf16 operands on mma.sync tensor cores
int attn_mma_supported(int dtype, const void* q2, int B, int H, int Lq, int Lk, const void* const* ptrs, const int64_t* lds, int n);
int attn_mma_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo,
                 const uint8_t* key_mask, float* lse, float* p_avg, int B, int H, int Lq, int Lk, float scale, cudaStream_t st,
                 const DropArgs& drop);
int attn_mma_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* d_o, int64_t lddo,
                 const uint8_t* key_mask, const float* lse, const float* dp_avg, void* dq, int64_t lddq, void* dk, int64_t lddk,
                 void* dv, int64_t lddv, int B, int H, int Lq, int Lk, float scale, cudaStream_t st, const DropArgs& drop);

// attention_sq.cu: single-query (Lq = 1) kernels for the decoders' time-aligned cross attention
int attn_sq_supported(int dtype, int Lq, int Lk, const void* p_avg, const void* dp_avg, const void* const* ptrs,
                      const int64_t* lds, int n);
int attn_sq_fwd(int dtype, const void* q1, const void* q2, int64_t ldq, const void* k1, const void* k2, int64_t ldk,
                const void* v, int64_t ldv, void* o, int64_t ldo, const uint8_t* key_mask, float* lse, int B, int H, int Lk,
                float scale, cudaStream_t st, const DropArgs& drop);
int attn_sq_bwd(int dtype, const void* q1, const void* q2, int64_t ldq, const void* k1, const void* k2, int64_t ldk,
                const void* v, int64_t ldv, const void* d_o, int64_t lddo, const uint8_t* key_mask, const float* lse,
                float* delta, void* dq1, void* dq2, int64_t lddq, void* dk1, void* dk2, int64_t lddk, void* dv, int64_t lddv,
                int B, int H, int Lk, float scale, cudaStream_t st, const DropArgs& drop);
int attn_tc_bwd_supported(int dtype, const void* q2, const void* dp_avg, const void* o, int B, int H, int Lq, int Lk,
                          const void* q, const void* k, const void* v, const void* d_o, const void* dq, const void* dk,
                          const void* dv, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, int64_t lddo, int64_t lddq,
                          int64_t lddk, int64_t lddv);
int attn_tc_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* o,
                int64_t ldo, const void* d_o, int64_t lddo, const uint8_t* key_mask, const float* lse, void* dq,
                int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int B, int H, int S, float scale,
                cudaStream_t st, const DropArgs& drop);

void attn_tc_set_trace(long long* buf);
void gemm_tc_set_trace(long long* buf);

}  // namespace stcat

using namespace stcat;

// Diagnostics: SM-clock timestamps at the phase boundaries of the tcgen05 attention forward (CTA 0, first 8 work items,
// 16 event slots per item; buf = 128 int64 in device memory, NULL switches it off).  See scripts/attn_timeline.py.
extern "C" int stcat_debug_attn_trace(void* buf) {
    attn_tc_set_trace((long long*)buf);
    return 0;
}

// Same for the tcgen05 GEMM (plain forward GEMMs, 128 x 256 tiles): CTA 0, its first 8 tiles, 8 event slots per tile
// (buf = 64 int64 in device memory).  See scripts/gemm_timeline.py.
extern "C" int stcat_debug_gemm_trace(void* buf) {
    gemm_tc_set_trace((long long*)buf);
    return 0;
}

// Launch counters per kernel family (host side; [0] single-query, [1] tcgen05, [2] mma.sync, [3] shared-memory fp32,
// [4] generic SIMT): lets a test assert which kernel served a shape (tests/test_gpu_bf16_parity.py).
static long long g_attn_counts[5] = {0, 0, 0, 0, 0};
extern "C" int stcat_debug_attn_counts(long long* out5) {
    STCAT_REQUIRE(out5, STCAT_EINVAL, "debug_attn_counts: null pointer");
    for (int i = 0; i < 5; ++i) out5[i] = g_attn_counts[i];
    return 0;
}

// Kernel selection shared by the plain and the dropout entry points.  With dropout (drop.thresh != 0) the single-query,
// tcgen05, mma.sync and generic kernels apply the counter-based mask of common.cuh; the fp32 shared-memory kernel has none.
static int attention_fwd_impl(const char* who, const void* q1, const void* q2, int64_t ldq, const void* k1, const void* k2,
                              int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo, int dtype, const uint8_t* key_mask,
                              float* lse, float* p_avg, int B, int H, int Lq, int Lk, int dh, float scale, void* stream,
                              const DropArgs& drop) {
    STCAT_REQUIRE(q1 && k1 && v && o && lse, STCAT_EINVAL, "%s: null pointer", who);
    STCAT_REQUIRE((q2 == nullptr) == (k2 == nullptr), STCAT_EINVAL, "%s: q2/k2 must both be set or both NULL", who);
    STCAT_REQUIRE(dh == DH, STCAT_ESHAPE, "%s: head dim %d unsupported (must be 32 = HIDDEN/HEADS)", who, dh);
    STCAT_REQUIRE(B >= 0 && H > 0 && Lq >= 0 && Lk > 0, STCAT_EINVAL, "%s: bad sizes B=%d H=%d Lq=%d Lk=%d", who, B, H, Lq, Lk);
    STCAT_REQUIRE(B <= 65535 && H <= 65535, STCAT_ESHAPE, "%s: B/H exceed grid limits", who);
    if (B == 0 || Lq == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    STCAT_REQUIRE(dtype == STCAT_F32 || dtype == STCAT_BF16, STCAT_EINVAL, "%s: bad dtype %d", who, dtype);
    const bool nodrop = drop.thresh == 0;
    {
        const void* ptrs[6] = {q1, q2, k1, k2, v, o};
        const int64_t lds[6] = {ldq, ldq, ldk, ldk, ldv, ldo};
        if (attn_sq_supported(dtype, Lq, Lk, p_avg, nullptr, ptrs, lds, 6))
            return ++g_attn_counts[0], attn_sq_fwd(dtype, q1, q2, ldq, k1, k2, ldk, v, ldv, o, ldo, key_mask, lse, B, H, Lk, scale, st, drop);
    }
    if (attn_tc_fwd_supported(dtype, q2, p_avg, B, H, Lq, Lk, q1, k1, v, o, ldq, ldk, ldv, ldo))
        return ++g_attn_counts[1], attn_tc_fwd(q1, ldq, k1, ldk, v, ldv, o, ldo, key_mask, lse, B, H, Lq, scale, st, drop);
    {
        const void* ptrs[4] = {q1, k1, v, o};
        const int64_t lds[4] = {ldq, ldk, ldv, ldo};
        if (attn_mma_supported(dtype, q2, B, H, Lq, Lk, ptrs, lds, 4))
            return ++g_attn_counts[2], attn_mma_fwd(q1, ldq, k1, ldk, v, ldv, o, ldo, key_mask, lse, p_avg, B, H, Lq, Lk, scale, st, drop);
    }
    if (nodrop && attn_small_supported(q2, B, H, Lq, Lk))
        return ++g_attn_counts[3], attn_small_fwd(dtype, q1, ldq, k1, ldk, v, ldv, o, ldo, key_mask, lse, p_avg, B, H, Lq, Lk, scale, st);
    ++g_attn_counts[4];
    if (dtype == STCAT_F32)
        return q2 ? launch_fwd<float, true>(q1, q2, ldq, k1, k2, ldk, v, ldv, o, ldo, key_mask, lse, p_avg, B, H, Lq, Lk, scale, st, drop)
                  : launch_fwd<float, false>(q1, q2, ldq, k1, k2, ldk, v, ldv, o, ldo, key_mask, lse, p_avg, B, H, Lq, Lk, scale, st, drop);
    return q2 ? launch_fwd<__nv_bfloat16, true>(q1, q2, ldq, k1, k2, ldk, v, ldv, o, ldo, key_mask, lse, p_avg, B, H, Lq, Lk, scale, st, drop)
              : launch_fwd<__nv_bfloat16, false>(q1, q2, ldq, k1, k2, ldk, v, ldv, o, ldo, key_mask, lse, p_avg, B, H, Lq, Lk, scale, st, drop);
}

static int attention_bwd_impl(const char* who, const void* q1, const void* q2, int64_t ldq, const void* k1, const void* k2,
                              int64_t ldk, const void* v, int64_t ldv, const void* o, int64_t ldo, const void* d_o, int64_t lddo,
                              int dtype, const uint8_t* key_mask, const float* lse, const float* dp_avg, float* delta, void* dq1,
                              void* dq2, int64_t lddq, void* dk1, void* dk2, int64_t lddk, void* dv, int64_t lddv, int B, int H,
                              int Lq, int Lk, int dh, float scale, void* stream, const DropArgs& drop) {
    STCAT_REQUIRE(q1 && k1 && v && d_o && lse && delta && dq1 && dk1 && dv, STCAT_EINVAL, "%s: null pointer", who);
    STCAT_REQUIRE((q2 == nullptr) == (k2 == nullptr), STCAT_EINVAL, "%s: q2/k2 must both be set or both NULL", who);
    STCAT_REQUIRE(!q2 || (dq2 && dk2), STCAT_EINVAL, "%s: dq2/dk2 required with q2/k2", who);
    STCAT_REQUIRE(dh == DH, STCAT_ESHAPE, "%s: head dim %d unsupported (must be 32)", who, dh);
    STCAT_REQUIRE(B >= 0 && H > 0 && Lq >= 0 && Lk > 0, STCAT_EINVAL, "%s: bad sizes", who);
    STCAT_REQUIRE(B <= 65535 && H <= 65535, STCAT_ESHAPE, "%s: B/H exceed grid limits", who);
    if (B == 0 || Lq == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    STCAT_REQUIRE(dtype == STCAT_F32 || dtype == STCAT_BF16, STCAT_EINVAL, "%s: bad dtype %d", who, dtype);
    const bool nodrop = drop.thresh == 0;
    {
        const void* ptrs[11] = {q1, q2, k1, k2, v, d_o, dq1, dq2, dk1, dk2, dv};
        const int64_t lds[11] = {ldq, ldq, ldk, ldk, ldv, lddo, lddq, lddq, lddk, lddk, lddv};
        if (attn_sq_supported(dtype, Lq, Lk, nullptr, dp_avg, ptrs, lds, 11))
            return ++g_attn_counts[0], attn_sq_bwd(dtype, q1, q2, ldq, k1, k2, ldk, v, ldv, d_o, lddo, key_mask, lse, delta, dq1, dq2, lddq, dk1,
                               dk2, lddk, dv, lddv, B, H, Lk, scale, st, drop);
    }
    if (attn_tc_bwd_supported(dtype, q2, dp_avg, o, B, H, Lq, Lk, q1, k1, v, d_o, dq1, dk1, dv, ldq, ldk, ldv, ldo, lddo,
                              lddq, lddk, lddv))
        return ++g_attn_counts[1], attn_tc_bwd(q1, ldq, k1, ldk, v, ldv, o, ldo, d_o, lddo, key_mask, lse, dq1, lddq, dk1, lddk, dv, lddv, B, H,
                           Lq, scale, st, drop);
    {
        const void* ptrs[7] = {q1, k1, v, d_o, dq1, dk1, dv};
        const int64_t lds[7] = {ldq, ldk, ldv, lddo, lddq, lddk, lddv};
        if (attn_mma_supported(dtype, q2, B, H, Lq, Lk, ptrs, lds, 7))
            return ++g_attn_counts[2], attn_mma_bwd(q1, ldq, k1, ldk, v, ldv, d_o, lddo, key_mask, lse, dp_avg, dq1, lddq, dk1, lddk, dv, lddv, B, H,
                                Lq, Lk, scale, st, drop);
    }
    if (nodrop && attn_small_supported(q2, B, H, Lq, Lk))
        return ++g_attn_counts[3], attn_small_bwd(dtype, q1, ldq, k1, ldk, v, ldv, d_o, lddo, key_mask, lse, dp_avg, dq1, lddq, dk1, lddk, dv, lddv,
                              B, H, Lq, Lk, scale, st);
    ++g_attn_counts[4];
    if (dtype == STCAT_F32)
        return q2 ? launch_bwd<float, true>(q1, q2, ldq, k1, k2, ldk, v, ldv, d_o, lddo, key_mask, lse, dp_avg, delta, dq1, dq2, lddq, dk1, dk2, lddk, dv, lddv, B, H, Lq, Lk, scale, st, drop)
                  : launch_bwd<float, false>(q1, q2, ldq, k1, k2, ldk, v, ldv, d_o, lddo, key_mask, lse, dp_avg, delta, dq1, dq2, lddq, dk1, dk2, lddk, dv, lddv, B, H, Lq, Lk, scale, st, drop);
    return q2 ? launch_bwd<__nv_bfloat16, true>(q1, q2, ldq, k1, k2, ldk, v, ldv, d_o, lddo, key_mask, lse, dp_avg, delta, dq1, dq2, lddq, dk1, dk2, lddk, dv, lddv, B, H, Lq, Lk, scale, st, drop)
              : launch_bwd<__nv_bfloat16, false>(q1, q2, ldq, k1, k2, ldk, v, ldv, d_o, lddo, key_mask, lse, dp_avg, delta, dq1, dq2, lddq, dk1, dk2, lddk, dv, lddv, B, H, Lq, Lk, scale, st, drop);
}

extern "C" int stcat_attention_fwd(const void* q1, const void* q2, int64_t ldq, const void* k1, const void* k2,
                                   int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo, int dtype,
                                   const uint8_t* key_mask, float* lse, float* p_avg, int B, int H, int Lq, int Lk,
                                   int dh, float scale, void* stream) {
    return attention_fwd_impl("attention_fwd", q1, q2, ldq, k1, k2, ldk, v, ldv, o, ldo, dtype, key_mask, lse, p_avg, B, H, Lq,
                              Lk, dh, scale, stream, DropArgs());
}

extern "C" int stcat_attention_bwd(const void* q1, const void* q2, int64_t ldq, const void* k1, const void* k2,
                                   int64_t ldk, const void* v, int64_t ldv, const void* o, int64_t ldo,
                                   const void* d_o, int64_t lddo, int dtype,
                                   const uint8_t* key_mask, const float* lse, const float* dp_avg, float* delta,
                                   void* dq1, void* dq2, int64_t lddq, void* dk1, void* dk2, int64_t lddk, void* dv,
                                   int64_t lddv, int B, int H, int Lq, int Lk, int dh, float scale, void* stream) {
    return attention_bwd_impl("attention_bwd", q1, q2, ldq, k1, k2, ldk, v, ldv, o, ldo, d_o, lddo, dtype, key_mask, lse, dp_avg,
                              delta, dq1, dq2, lddq, dk1, dk2, lddk, dv, lddv, B, H, Lq, Lk, dh, scale, stream, DropArgs());
}

// Attention with dropout on the probabilities (nn.MultiheadAttention dropout=p in train mode, torch functional.py;
// reference attention.py:381) with the counter-based mask of common.cuh: element ((b*H + h)*Lq + i)*Lk + j of the site
// that drew (seed, offset).  The `weights` output (p_avg) is the head average of the DROPPED probabilities, as in the
// reference; lse is that of the undropped softmax.  `o` (the forward output, optional) lets the backward take
// delta = dO . O instead of recomputing it.
extern "C" int stcat_attention_dropout_fwd(const void* q1, const void* q2, int64_t ldq, const void* k1, const void* k2,
                                           int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo, int dtype,
                                           const uint8_t* key_mask, float* lse, float* p_avg, int B, int H, int Lq, int Lk,
                                           int dh, float scale, float drop_p, uint64_t seed, uint64_t offset,
                                           const void* keep_bits, int bits_wpr, void* stream) {
    STCAT_REQUIRE(drop_p >= 0.f && drop_p < 1.f, STCAT_EINVAL, "attention_dropout_fwd: p=%f", (double)drop_p);
    STCAT_REQUIRE(!keep_bits || bits_wpr * 32 >= Lk + 32, STCAT_EINVAL, "attention_dropout_fwd: keep_bits needs Lk / 32 + 1 words per row");
    DropArgs d = make_drop(drop_p, seed, offset);
    if (d.thresh) { d.bits = (const uint32_t*)keep_bits; d.wpr = bits_wpr; }
    return attention_fwd_impl("attention_dropout_fwd", q1, q2, ldq, k1, k2, ldk, v, ldv, o, ldo, dtype, key_mask, lse, p_avg, B,
                              H, Lq, Lk, dh, scale, stream, d);
}

extern "C" int stcat_attention_dropout_bwd(const void* q1, const void* q2, int64_t ldq, const void* k1, const void* k2,
                                           int64_t ldk, const void* v, int64_t ldv, const void* o, int64_t ldo,
                                           const void* d_o, int64_t lddo, int dtype,
                                           const uint8_t* key_mask, const float* lse, const float* dp_avg, float* delta,
                                           void* dq1, void* dq2, int64_t lddq, void* dk1, void* dk2, int64_t lddk, void* dv,
                                           int64_t lddv, int B, int H, int Lq, int Lk, int dh, float scale, float drop_p,
                                           uint64_t seed, uint64_t offset, const void* keep_bits, int bits_wpr, void* stream) {
    STCAT_REQUIRE(drop_p >= 0.f && drop_p < 1.f, STCAT_EINVAL, "attention_dropout_bwd: p=%f", (double)drop_p);
    STCAT_REQUIRE(!keep_bits || bits_wpr * 32 >= Lk + 32, STCAT_EINVAL, "attention_dropout_bwd: keep_bits needs Lk / 32 + 1 words per row");
    DropArgs d = make_drop(drop_p, seed, offset);
    if (d.thresh) { d.bits = (const uint32_t*)keep_bits; d.wpr = bits_wpr; }
    return attention_bwd_impl("attention_dropout_bwd", q1, q2, ldq, k1, k2, ldk, v, ldv, o, ldo, d_o, lddo, dtype, key_mask, lse,
                              dp_avg, delta, dq1, dq2, lddq, dk1, dk2, lddk, dv, lddv, B, H, Lq, Lk, dh, scale, stream, d);
}
