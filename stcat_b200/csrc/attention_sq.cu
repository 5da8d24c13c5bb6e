// Single-query attention (Lq = 1): the decoders' time-aligned cross attention, where the one query of frame f
// attends to the HW+L memory tokens of frame f (reference query_decoder.py:350-429 custom MultiheadAttention with
// per-head [content(32) ; positional(32)] concat, and :615-651 nn.MultiheadAttention).  B = frames, H = 8 heads.
//
// This is a GEMV-shaped, HBM-bound op: per (frame, head) the kernel streams K (one or two 32-wide parts) and V
// exactly once in the forward (algorithmic bytes = Lk * (parts + 1) * 64 B in bf16) and K, V once plus dK, dV once in
// the backward.  One 128-thread block per (frame, head): a thread owns keys tid, tid+128, ... for the dot products
// (16-byte vector loads of whole 32-element rows), probabilities live in shared memory, the PV / dq reductions run
// with lanes across the 32 feature dims so that V rows are read as 64 B coalesced segments.
#include "common.cuh"
#include <math.h>

namespace stcat {

constexpr int SQ_THREADS = 128;
constexpr int SQ_MAX_LK = 4096;  // scores kept in shared memory

template <typename T> struct Row32;
template <> struct Row32<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[32]) {
        const uint4* p4 = reinterpret_cast<const uint4*>(p);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint4 a = __ldg(p4 + j);
            const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[e]));
                v[8 * j + 2 * e] = f.x;
                v[8 * j + 2 * e + 1] = f.y;
            }
        }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[32]) {
        uint4* p4 = reinterpret_cast<uint4*>(p);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                __nv_bfloat162 b = __floats2bfloat162_rn(v[8 * j + 2 * e], v[8 * j + 2 * e + 1]);
                w[e] = *reinterpret_cast<uint32_t*>(&b);
            }
            p4[j] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
};
template <> struct Row32<float> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[32]) {
        const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 a = __ldg(p4 + j);
            v[4 * j] = a.x; v[4 * j + 1] = a.y; v[4 * j + 2] = a.z; v[4 * j + 3] = a.w;
        }
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[32]) {
        float4* p4 = reinterpret_cast<float4*>(p);
#pragma unroll
        for (int j = 0; j < 8; ++j) p4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
};

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = is_max ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int w = 1; w < SQ_THREADS / 32; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
    return r;
}

template <typename T, bool TWO>
__global__ void __launch_bounds__(SQ_THREADS)
attn_sq_fwd_kernel(const T* __restrict__ q1, const T* __restrict__ q2, int64_t ldq, const T* __restrict__ k1,
                   const T* __restrict__ k2, int64_t ldk, const T* __restrict__ v, int64_t ldv, T* __restrict__ o,
                   int64_t ldo, const uint8_t* __restrict__ key_mask, float* __restrict__ lse, int H, int Lk, float scale,
                   const DropArgs drop_in) {
    const DropArgs drop = drop_resolve(drop_in);
    extern __shared__ float sm[];  // scores / probabilities [Lk]
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float red[SQ_THREADS / 32];
    __shared__ float part[SQ_THREADS / 32][32];
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int col = h * 32;
    const int64_t kbase = (int64_t)b * Lk;
    float qa[32], qb[32];
    Row32<T>::load(q1 + (int64_t)b * ldq + col, qa);
    if (TWO) Row32<T>::load(q2 + (int64_t)b * ldq + col, qb);
    float mx = -INFINITY;
    for (int j = threadIdx.x; j < Lk; j += SQ_THREADS) {
        float s = -INFINITY;
        if (!(key_mask && key_mask[kbase + j])) {
            float kr[32];
            Row32<T>::load(k1 + (kbase + j) * ldk + col, kr);
            float d = 0.f;
#pragma unroll
            for (int e = 0; e < 32; ++e) d = fmaf(qa[e], kr[e], d);
            if (TWO) {
                Row32<T>::load(k2 + (kbase + j) * ldk + col, kr);
#pragma unroll
                for (int e = 0; e < 32; ++e) d = fmaf(qb[e], kr[e], d);
            }
            s = d * scale;
        }
        sm[j] = s;
        mx = fmaxf(mx, s);
    }
    mx = block_reduce(mx, red, true);
    float sum = 0.f;
    for (int j = threadIdx.x; j < Lk; j += SQ_THREADS) {
        const float s = sm[j];
        const float pj = (s == -INFINITY) ? 0.f : expf(s - mx);
        // dropout on the probabilities: element (b, h, 0, j) of [B, H, 1, Lk]; the row sum stays undropped
        sm[j] = drop.thresh ? drop_apply(drop, (uint64_t)blockIdx.x * Lk + j, pj) : pj;
        sum += pj;
    }
    sum = block_reduce(sum, red, false);  // (its barriers also publish sm[] to every thread)
    // o[d] = sum_j p_j v[j][d]: lane = feature dim, warp = key subset
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc = 0.f;
    for (int j = warp; j < Lk; j += SQ_THREADS / 32) acc = fmaf(sm[j], to_f32<T>(v[(kbase + j) * ldv + col + lane]), acc);
    part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < SQ_THREADS / 32; ++w) t += part[w][lane];
        const float inv = sum > 0.f ? 1.f / sum : 0.f;
        o[(int64_t)b * ldo + col + lane] = from_f32<T>(t * inv);
        if (lane == 0) lse[(int64_t)b * H + h] = sum > 0.f ? mx + logf(sum) : -INFINITY;
    }
}

template <typename T, bool TWO>
__global__ void __launch_bounds__(SQ_THREADS)
attn_sq_bwd_kernel(const T* __restrict__ q1, const T* __restrict__ q2, int64_t ldq, const T* __restrict__ k1,
                   const T* __restrict__ k2, int64_t ldk, const T* __restrict__ v, int64_t ldv, const T* __restrict__ d_o,
                   int64_t lddo, const uint8_t* __restrict__ key_mask, const float* __restrict__ lse,
                   float* __restrict__ delta_out, T* __restrict__ dq1, T* __restrict__ dq2, int64_t lddq,
                   T* __restrict__ dk1, T* __restrict__ dk2, int64_t lddk, T* __restrict__ dv, int64_t lddv, int H, int Lk,
                   float scale, const DropArgs drop_in) {
    const DropArgs drop = drop_resolve(drop_in);
    extern __shared__ float sm[];  // p [Lk], dp [Lk]
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float red[SQ_THREADS / 32];
    __shared__ float part[SQ_THREADS / 32][64];
    float* sp = sm;
    float* sdp = sm + Lk;
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int col = h * 32;
    const int64_t kbase = (int64_t)b * Lk;
    const float l = lse[(int64_t)b * H + h];
    float qa[32], qb[32], g[32];
    Row32<T>::load(q1 + (int64_t)b * ldq + col, qa);
    if (TWO) Row32<T>::load(q2 + (int64_t)b * ldq + col, qb);
    Row32<T>::load(d_o + (int64_t)b * lddo + col, g);
    float dl = 0.f;
    for (int j = threadIdx.x; j < Lk; j += SQ_THREADS) {
        float pj = 0.f, dpj = 0.f;
        if (!(key_mask && key_mask[kbase + j]) && l != -INFINITY) {
            float kr[32];
            Row32<T>::load(k1 + (kbase + j) * ldk + col, kr);
            float d = 0.f;
#pragma unroll
            for (int e = 0; e < 32; ++e) d = fmaf(qa[e], kr[e], d);
            if (TWO) {
                Row32<T>::load(k2 + (kbase + j) * ldk + col, kr);
#pragma unroll
                for (int e = 0; e < 32; ++e) d = fmaf(qb[e], kr[e], d);
            }
            pj = expf(d * scale - l);
            Row32<T>::load(v + (kbase + j) * ldv + col, kr);
#pragma unroll
            for (int e = 0; e < 32; ++e) dpj = fmaf(g[e], kr[e], dpj);
            // o was formed from the DROPPED probabilities: its gradient reaches p through the mask
            if (drop.thresh) dpj = drop_apply(drop, (uint64_t)blockIdx.x * Lk + j, dpj);
        }
        sp[j] = pj;
        sdp[j] = dpj;
        dl = fmaf(pj, dpj, dl);
    }
    dl = block_reduce(dl, red, false);
    if (threadIdx.x == 0 && delta_out) delta_out[(int64_t)b * H + h] = dl;
    // dk_j = ds_j q, dv_j = p_j dO: one thread per key, whole 32-element rows
    for (int j = threadIdx.x; j < Lk; j += SQ_THREADS) {
        const float pj = sp[j];
        const float ds = pj * (sdp[j] - dl) * scale;
        sdp[j] = ds;  // reused below for dq
        float r[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) r[e] = ds * qa[e];
        Row32<T>::store(dk1 + (kbase + j) * lddk + col, r);
        if (TWO) {
#pragma unroll
            for (int e = 0; e < 32; ++e) r[e] = ds * qb[e];
            Row32<T>::store(dk2 + (kbase + j) * lddk + col, r);
        }
        const float pm = drop.thresh ? drop_apply(drop, (uint64_t)blockIdx.x * Lk + j, pj) : pj;  // what multiplied V
#pragma unroll
        for (int e = 0; e < 32; ++e) r[e] = pm * g[e];
        Row32<T>::store(dv + (kbase + j) * lddv + col, r);
    }
    __syncthreads();
    // dq = sum_j ds_j k_j: lane = feature dim, warp = key subset
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float a1 = 0.f, a2 = 0.f;
    for (int j = warp; j < Lk; j += SQ_THREADS / 32) {
        const float ds = sdp[j];
        a1 = fmaf(ds, to_f32<T>(k1[(kbase + j) * ldk + col + lane]), a1);
        if (TWO) a2 = fmaf(ds, to_f32<T>(k2[(kbase + j) * ldk + col + lane]), a2);
    }
    part[warp][lane] = a1;
    part[warp][32 + lane] = a2;
    __syncthreads();
    if (warp == 0) {
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int w = 0; w < SQ_THREADS / 32; ++w) { t1 += part[w][lane]; t2 += part[w][32 + lane]; }
        dq1[(int64_t)b * lddq + col + lane] = from_f32<T>(t1);
        if (TWO) dq2[(int64_t)b * lddq + col + lane] = from_f32<T>(t2);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// bf16 kernels (the benchmarked path).  Same block per (frame, head), but FOUR threads per key, each owning 8 of the head's
// 32 feature dims (one 16-byte load per row part): a warp-wide load covers 8 keys x 64 B, the key loop is unrolled so
// that every thread has 8 independent 16-byte loads in flight (the first version issued one dependent row at a time and ran
// at ~1/5 of the HBM rate: 17.9 us for 20.8 MB), and the P V / dq reductions run over the 8 key slots of a warp with three
// shuffle steps instead of 2-byte loads with lanes across the feature dims.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4 a, float (&v)[8]) {
    const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[e]));
        v[2 * e] = f.x;
        v[2 * e + 1] = f.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        __nv_bfloat162 b = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
        w[e] = *reinterpret_cast<uint32_t*>(&b);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ float dot8(const float (&q)[8], const uint4 a) {
    float k[8];
    unpack8(a, k);
    float d = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) d = fmaf(q[e], k[e], d);
    return d;
}
__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

template <bool TWO>
__global__ void __launch_bounds__(SQ_THREADS)
attn_sq_fwd_bf16_kernel(const __nv_bfloat16* __restrict__ q1, const __nv_bfloat16* __restrict__ q2, int64_t ldq,
                        const __nv_bfloat16* __restrict__ k1, const __nv_bfloat16* __restrict__ k2, int64_t ldk,
                        const __nv_bfloat16* __restrict__ v, int64_t ldv, __nv_bfloat16* __restrict__ o, int64_t ldo,
                        const uint8_t* __restrict__ key_mask, float* __restrict__ lse, int H, int Lk, float scale,
                        const DropArgs drop_in) {
    const DropArgs drop = drop_resolve(drop_in);
    extern __shared__ float sm[];  // scores / probabilities [Lk]
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float red[SQ_THREADS / 32];
    __shared__ float part[SQ_THREADS / 32][32];
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int tid = threadIdx.x, ks = tid >> 2, pt = tid & 3;  // key slot (32 per block), 8-dim part of the head
    const int col = h * 32 + pt * 8;
    const int64_t kbase = (int64_t)b * Lk;
    float qa[8], qb[8];
    unpack8(ldg16(q1 + (int64_t)b * ldq + col), qa);
    if (TWO) unpack8(ldg16(q2 + (int64_t)b * ldq + col), qb);
    float mx = -INFINITY;
    constexpr int UN = 8;  // keys per thread and round: all loads of a round are issued before the first use (clamped row
                           // index instead of a branch, so that nothing serialises on the mask byte or on the previous key)
    for (int j0 = 0; j0 < Lk; j0 += 32 * UN) {
        uint4 ka[UN], kb[UN];
        uint8_t mk[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int j = min(j0 + 32 * u + ks, Lk - 1);
            ka[u] = ldg16(k1 + (kbase + j) * ldk + col);
            if (TWO) kb[u] = ldg16(k2 + (kbase + j) * ldk + col);
            mk[u] = key_mask ? key_mask[kbase + j] : (uint8_t)0;
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int j = j0 + 32 * u + ks;
            float d = dot8(qa, ka[u]);
            if (TWO) d += dot8(qb, kb[u]);
            d += __shfl_xor_sync(0xffffffffu, d, 1);
            d += __shfl_xor_sync(0xffffffffu, d, 2);
            const float sc = (j < Lk && !mk[u]) ? d * scale : -INFINITY;
            if (pt == 0 && j < Lk) sm[j] = sc;
            mx = fmaxf(mx, sc);
        }
    }
    mx = block_reduce(mx, red, true);  // (its barriers also publish sm[] to every thread)
    float sum = 0.f;
    for (int j = tid; j < Lk; j += SQ_THREADS) {
        const float sc = sm[j];
        const float pj = (sc == -INFINITY) ? 0.f : expf(sc - mx);
        // dropout on the probabilities: element (b, h, 0, j) of [B, H, 1, Lk]; the row sum stays undropped
        sm[j] = drop.thresh ? drop_apply(drop, (uint64_t)blockIdx.x * Lk + j, pj) : pj;
        sum += pj;
    }
    sum = block_reduce(sum, red, false);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int j0 = 0; j0 < Lk; j0 += 32 * UN) {
        uint4 va[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) va[u] = ldg16(v + (kbase + min(j0 + 32 * u + ks, Lk - 1)) * ldv + col);
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int j = j0 + 32 * u + ks;
            const float pj = j < Lk ? sm[j] : 0.f;
            float vr[8];
            unpack8(va[u], vr);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = fmaf(pj, vr[e], acc[e]);
        }
    }
    // sum over the 8 key slots of the warp (lanes with equal pt), then over the 4 warps
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 4);
        acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8);
        acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
    }
    if (lane < 4) {
#pragma unroll
        for (int e = 0; e < 8; ++e) part[warp][lane * 8 + e] = acc[e];
    }
    __syncthreads();
    if (warp == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < SQ_THREADS / 32; ++w) t += part[w][lane];
        const float inv = sum > 0.f ? 1.f / sum : 0.f;
        o[(int64_t)b * ldo + h * 32 + lane] = __float2bfloat16_rn(t * inv);
        if (lane == 0) lse[(int64_t)b * H + h] = sum > 0.f ? mx + logf(sum) : -INFINITY;
    }
}

template <bool TWO>
__global__ void __launch_bounds__(SQ_THREADS)
attn_sq_bwd_bf16_kernel(const __nv_bfloat16* __restrict__ q1, const __nv_bfloat16* __restrict__ q2, int64_t ldq,
                        const __nv_bfloat16* __restrict__ k1, const __nv_bfloat16* __restrict__ k2, int64_t ldk,
                        const __nv_bfloat16* __restrict__ v, int64_t ldv, const __nv_bfloat16* __restrict__ d_o, int64_t lddo,
                        const uint8_t* __restrict__ key_mask, const float* __restrict__ lse, float* __restrict__ delta_out,
                        __nv_bfloat16* __restrict__ dq1, __nv_bfloat16* __restrict__ dq2, int64_t lddq,
                        __nv_bfloat16* __restrict__ dk1, __nv_bfloat16* __restrict__ dk2, int64_t lddk,
                        __nv_bfloat16* __restrict__ dv, int64_t lddv, int H, int Lk, float scale, const DropArgs drop_in) {
    const DropArgs drop = drop_resolve(drop_in);
    extern __shared__ float sm[];  // p [Lk], dp [Lk]
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float red[SQ_THREADS / 32];
    __shared__ float part[SQ_THREADS / 32][64];
    float* sp = sm;
    float* sdp = sm + Lk;
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int tid = threadIdx.x, ks = tid >> 2, pt = tid & 3;
    const int col = h * 32 + pt * 8;
    const int64_t kbase = (int64_t)b * Lk;
    const float l = lse[(int64_t)b * H + h];
    float qa[8], qb[8], g[8];
    unpack8(ldg16(q1 + (int64_t)b * ldq + col), qa);
    if (TWO) unpack8(ldg16(q2 + (int64_t)b * ldq + col), qb);
    unpack8(ldg16(d_o + (int64_t)b * lddo + col), g);
    float dl = 0.f;
    constexpr int UN = 4;  // keys per thread and round, loads issued before the first use (see the forward kernel)
    for (int j0 = 0; j0 < Lk; j0 += 32 * UN) {
        uint4 ka[UN], kb[UN], va[UN];
        uint8_t mk[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int j = min(j0 + 32 * u + ks, Lk - 1);
            ka[u] = ldg16(k1 + (kbase + j) * ldk + col);
            if (TWO) kb[u] = ldg16(k2 + (kbase + j) * ldk + col);
            va[u] = ldg16(v + (kbase + j) * ldv + col);
            mk[u] = key_mask ? key_mask[kbase + j] : (uint8_t)0;
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int j = j0 + 32 * u + ks;
            const bool valid = j < Lk && !mk[u] && l != -INFINITY;
            float d = dot8(qa, ka[u]);
            if (TWO) d += dot8(qb, kb[u]);
            float dpj = dot8(g, va[u]);
            d += __shfl_xor_sync(0xffffffffu, d, 1);
            d += __shfl_xor_sync(0xffffffffu, d, 2);
            dpj += __shfl_xor_sync(0xffffffffu, dpj, 1);
            dpj += __shfl_xor_sync(0xffffffffu, dpj, 2);
            const float pj = valid ? expf(d * scale - l) : 0.f;
            dpj = valid ? dpj : 0.f;
            // o was formed from the DROPPED probabilities: its gradient reaches p through the mask
            if (valid && drop.thresh) dpj = drop_apply(drop, (uint64_t)blockIdx.x * Lk + j, dpj);
            if (pt == 0 && j < Lk) {
                sp[j] = pj;
                sdp[j] = dpj;
                dl = fmaf(pj, dpj, dl);
            }
        }
    }
    dl = block_reduce(dl, red, false);  // (its barriers also publish sp / sdp)
    if (tid == 0 && delta_out) delta_out[(int64_t)b * H + h] = dl;
    // dk_j = ds_j q, dv_j = p_j dO (16-byte stores, 8 keys x 64 B per warp instruction); dq = sum_j ds_j k_j
    float a1[8], a2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) a1[e] = a2[e] = 0.f;
    for (int j0 = 0; j0 < Lk; j0 += 32 * UN) {
        uint4 ka[UN], kb[UN];  // K rows again (L1 / L2 hits: this block read them a moment ago)
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int j = min(j0 + 32 * u + ks, Lk - 1);
            ka[u] = ldg16(k1 + (kbase + j) * ldk + col);
            if (TWO) kb[u] = ldg16(k2 + (kbase + j) * ldk + col);
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int j = j0 + 32 * u + ks;
            if (j < Lk) {
                const float pj = sp[j];
                const float ds = pj * (sdp[j] - dl) * scale;
                float r[8], kr[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) r[e] = ds * qa[e];
                *reinterpret_cast<uint4*>(dk1 + (kbase + j) * lddk + col) = pack8(r);
                if (TWO) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) r[e] = ds * qb[e];
                    *reinterpret_cast<uint4*>(dk2 + (kbase + j) * lddk + col) = pack8(r);
                }
                const float pm = drop.thresh ? drop_apply(drop, (uint64_t)blockIdx.x * Lk + j, pj) : pj;  // what multiplied V
#pragma unroll
                for (int e = 0; e < 8; ++e) r[e] = pm * g[e];
                *reinterpret_cast<uint4*>(dv + (kbase + j) * lddv + col) = pack8(r);
                unpack8(ka[u], kr);
#pragma unroll
                for (int e = 0; e < 8; ++e) a1[e] = fmaf(ds, kr[e], a1[e]);
                if (TWO) {
                    unpack8(kb[u], kr);
#pragma unroll
                    for (int e = 0; e < 8; ++e) a2[e] = fmaf(ds, kr[e], a2[e]);
                }
            }
        }
    }
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        a1[e] += __shfl_xor_sync(0xffffffffu, a1[e], 4);
        a1[e] += __shfl_xor_sync(0xffffffffu, a1[e], 8);
        a1[e] += __shfl_xor_sync(0xffffffffu, a1[e], 16);
        if (TWO) {
            a2[e] += __shfl_xor_sync(0xffffffffu, a2[e], 4);
            a2[e] += __shfl_xor_sync(0xffffffffu, a2[e], 8);
            a2[e] += __shfl_xor_sync(0xffffffffu, a2[e], 16);
        }
    }
    if (lane < 4) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            part[warp][lane * 8 + e] = a1[e];
            part[warp][32 + lane * 8 + e] = a2[e];
        }
    }
    __syncthreads();
    if (warp == 0) {
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int w = 0; w < SQ_THREADS / 32; ++w) { t1 += part[w][lane]; t2 += part[w][32 + lane]; }
        dq1[(int64_t)b * lddq + h * 32 + lane] = __float2bfloat16_rn(t1);
        if (TWO) dq2[(int64_t)b * lddq + h * 32 + lane] = __float2bfloat16_rn(t2);
    }
}

static bool vec_ok(const void* p, int64_t ld, int elem_bytes) {
    return p == nullptr || ((((uintptr_t)p) & 15) == 0 && (ld * elem_bytes) % 16 == 0);
}

int attn_sq_supported(int dtype, int Lq, int Lk, const void* p_avg, const void* dp_avg, const void* const* ptrs,
                      const int64_t* lds, int n) {
    if (Lq != 1 || p_avg || dp_avg || Lk > SQ_MAX_LK) return 0;
    const int es = dtype == STCAT_BF16 ? 2 : 4;
    for (int i = 0; i < n; ++i)
        if (!vec_ok(ptrs[i], lds[i], es)) return 0;
    return 1;
}

template <typename T, bool TWO>
static int launch_sq_fwd(const void* q1, const void* q2, int64_t ldq, const void* k1, const void* k2, int64_t ldk,
                         const void* v, int64_t ldv, void* o, int64_t ldo, const uint8_t* key_mask, float* lse, int B, int H,
                         int Lk, float scale, cudaStream_t st, const DropArgs& drop) {
    if constexpr (sizeof(T) == 2) {
        typedef __nv_bfloat16 B16;
        launch_pdl(attn_sq_fwd_bf16_kernel<TWO>, dim3(B * H), dim3(SQ_THREADS), Lk * sizeof(float), st,
            (const B16*)q1, (const B16*)q2, ldq, (const B16*)k1, (const B16*)k2, ldk, (const B16*)v, ldv, (B16*)o, ldo, key_mask, lse, H, Lk, scale, drop);
        return check_launch("attn_sq_fwd_bf16_kernel");
    } else {
        launch_pdl(attn_sq_fwd_kernel<T, TWO>, dim3(B * H), dim3(SQ_THREADS), Lk * sizeof(float), st,
            (const T*)q1, (const T*)q2, ldq, (const T*)k1, (const T*)k2, ldk, (const T*)v, ldv, (T*)o, ldo, key_mask, lse, H, Lk, scale, drop);
        return check_launch("attn_sq_fwd_kernel");
    }
}

template <typename T, bool TWO>
static int launch_sq_bwd(const void* q1, const void* q2, int64_t ldq, const void* k1, const void* k2, int64_t ldk,
                         const void* v, int64_t ldv, const void* d_o, int64_t lddo, const uint8_t* key_mask, const float* lse,
                         float* delta, void* dq1, void* dq2, int64_t lddq, void* dk1, void* dk2, int64_t lddk, void* dv,
                         int64_t lddv, int B, int H, int Lk, float scale, cudaStream_t st, const DropArgs& drop) {
    if constexpr (sizeof(T) == 2) {
        typedef __nv_bfloat16 B16;
        launch_pdl(attn_sq_bwd_bf16_kernel<TWO>, dim3(B * H), dim3(SQ_THREADS), 2 * Lk * sizeof(float), st,
            (const B16*)q1, (const B16*)q2, ldq, (const B16*)k1, (const B16*)k2, ldk, (const B16*)v, ldv, (const B16*)d_o, lddo, key_mask, lse,
            delta, (B16*)dq1, (B16*)dq2, lddq, (B16*)dk1, (B16*)dk2, lddk, (B16*)dv, lddv, H, Lk, scale, drop);
        return check_launch("attn_sq_bwd_bf16_kernel");
    } else {
        launch_pdl(attn_sq_bwd_kernel<T, TWO>, dim3(B * H), dim3(SQ_THREADS), 2 * Lk * sizeof(float), st,
            (const T*)q1, (const T*)q2, ldq, (const T*)k1, (const T*)k2, ldk, (const T*)v, ldv, (const T*)d_o, lddo, key_mask, lse,
            delta, (T*)dq1, (T*)dq2, lddq, (T*)dk1, (T*)dk2, lddk, (T*)dv, lddv, H, Lk, scale, drop);
        return check_launch("attn_sq_bwd_kernel");
    }
}

int attn_sq_fwd(int dtype, const void* q1, const void* q2, int64_t ldq, const void* k1, const void* k2, int64_t ldk,
                const void* v, int64_t ldv, void* o, int64_t ldo, const uint8_t* key_mask, float* lse, int B, int H, int Lk,
                float scale, cudaStream_t st, const DropArgs& drop) {
    if (dtype == STCAT_BF16)
        return q2 ? launch_sq_fwd<__nv_bfloat16, true>(q1, q2, ldq, k1, k2, ldk, v, ldv, o, ldo, key_mask, lse, B, H, Lk, scale, st, drop)
                  : launch_sq_fwd<__nv_bfloat16, false>(q1, q2, ldq, k1, k2, ldk, v, ldv, o, ldo, key_mask, lse, B, H, Lk, scale, st, drop);
    return q2 ? launch_sq_fwd<float, true>(q1, q2, ldq, k1, k2, ldk, v, ldv, o, ldo, key_mask, lse, B, H, Lk, scale, st, drop)
              : launch_sq_fwd<float, false>(q1, q2, ldq, k1, k2, ldk, v, ldv, o, ldo, key_mask, lse, B, H, Lk, scale, st, drop);
}

int attn_sq_bwd(int dtype, const void* q1, const void* q2, int64_t ldq, const void* k1, const void* k2, int64_t ldk,
                const void* v, int64_t ldv, const void* d_o, int64_t lddo, const uint8_t* key_mask, const float* lse,
                float* delta, void* dq1, void* dq2, int64_t lddq, void* dk1, void* dk2, int64_t lddk, void* dv, int64_t lddv,
                int B, int H, int Lk, float scale, cudaStream_t st, const DropArgs& drop) {
    if (dtype == STCAT_BF16)
        return q2 ? launch_sq_bwd<__nv_bfloat16, true>(q1, q2, ldq, k1, k2, ldk, v, ldv, d_o, lddo, key_mask, lse, delta, dq1, dq2, lddq, dk1, dk2, lddk, dv, lddv, B, H, Lk, scale, st, drop)
                  : launch_sq_bwd<__nv_bfloat16, false>(q1, q2, ldq, k1, k2, ldk, v, ldv, d_o, lddo, key_mask, lse, delta, dq1, dq2, lddq, dk1, dk2, lddk, dv, lddv, B, H, Lk, scale, st, drop);
    return q2 ? launch_sq_bwd<float, true>(q1, q2, ldq, k1, k2, ldk, v, ldv, d_o, lddo, key_mask, lse, delta, dq1, dq2, lddq, dk1, dk2, lddk, dv, lddv, B, H, Lk, scale, st, drop)
              : launch_sq_bwd<float, false>(q1, q2, ldq, k1, k2, ldk, v, ldv, d_o, lddo, key_mask, lse, delta, dq1, dq2, lddq, dk1, dk2, lddk, dv, lddv, B, H, Lk, scale, st, drop);
}

}  // namespace stcat
