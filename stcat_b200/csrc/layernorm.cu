// Residual-add + LayerNorm forward / backward for d = 256 rows (HBM-bound; one warp per row, two
// 128-bit loads per lane per operand, shuffle reductions, no shared memory in the forward).
//
// Algorithmic bytes per row (fp32): fwd reads x, res (2*1 KB) and writes y (1 KB) [+0.5 KB bf16 copy];
// bwd reads dy, x, res (3 KB) and writes dz (1 KB) [+0.5 KB bf16 copy]; the column sums it keeps anyway for
// dgamma / dbeta also give the bias gradient of the Linear in front of the norm (db = colsum(dz)) for free.
#include "common.cuh"

namespace stcat {

constexpr int LN_D = 256;
constexpr int LN_ROWS_PER_BLOCK = 8;  // 8 warps

__device__ __forceinline__ void load_row(const float* __restrict__ p, int lane, float v[8]) {
    float4 a = *reinterpret_cast<const float4*>(p + lane * 4);
    float4 b = *reinterpret_cast<const float4*>(p + 128 + lane * 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store_row(float* __restrict__ p, int lane, const float v[8]) {
    *reinterpret_cast<float4*>(p + lane * 4) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 128 + lane * 4) = make_float4(v[4], v[5], v[6], v[7]);
}

__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ res, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float* __restrict__ y, __nv_bfloat16* __restrict__ y_bf16,
                     float* __restrict__ mean, float* __restrict__ rstd, int rows, float eps, const DropArgs drop_in) {
    const DropArgs drop = drop_resolve(drop_in);
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float g[8], bt[8];
    load_row(gamma, lane, g);
    load_row(beta, lane, bt);
    for (int row = blockIdx.x * LN_ROWS_PER_BLOCK + warp; row < rows; row += gridDim.x * LN_ROWS_PER_BLOCK) {
        float z[8];
        load_row(x + (int64_t)row * LN_D, lane, z);
        if (drop.thresh) {  // train-mode dropout on x (the block output in front of the residual), element index row * d + col
            const uint64_t e0 = (uint64_t)row * LN_D + lane * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) { z[i] = drop_apply(drop, e0 + i, z[i]); z[4 + i] = drop_apply(drop, e0 + 128 + i, z[4 + i]); }
        }
        if (res) {
            float r[8];
            load_row(res + (int64_t)row * LN_D, lane, r);
#pragma unroll
            for (int i = 0; i < 8; ++i) z[i] += r[i];
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += z[i];
        const float mu = warp_sum(s) * (1.f / LN_D);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { float c = z[i] - mu; q += c * c; }
        const float var = warp_sum(q) * (1.f / LN_D);
        const float rs = rsqrtf(var + eps);
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = (z[i] - mu) * rs * g[i] + bt[i];
        store_row(y + (int64_t)row * LN_D, lane, o);
        if (y_bf16) {
            __nv_bfloat16* yb = y_bf16 + (int64_t)row * LN_D;
            __nv_bfloat162 p0 = __floats2bfloat162_rn(o[0], o[1]), p1 = __floats2bfloat162_rn(o[2], o[3]);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(o[4], o[5]), p3 = __floats2bfloat162_rn(o[6], o[7]);
            *reinterpret_cast<uint2*>(yb + lane * 4) = make_uint2(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1));
            *reinterpret_cast<uint2*>(yb + 128 + lane * 4) = make_uint2(*reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
        }
        if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
    }
}

__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ res,
                     const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                     float* __restrict__ dz, __nv_bfloat16* __restrict__ dz_bf16, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, float* __restrict__ dbias, int rows, const DropArgs drop_in) {
    const DropArgs drop = drop_resolve(drop_in);
    __shared__ float sg[LN_ROWS_PER_BLOCK][LN_D];
    __shared__ float sb[LN_ROWS_PER_BLOCK][LN_D];
    __shared__ float sz[LN_ROWS_PER_BLOCK][LN_D];
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float g[8];
    load_row(gamma, lane, g);
    float ag[8], ab[8], az[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { ag[i] = 0.f; ab[i] = 0.f; az[i] = 0.f; }
    for (int row = blockIdx.x * LN_ROWS_PER_BLOCK + warp; row < rows; row += gridDim.x * LN_ROWS_PER_BLOCK) {
        float z[8], d[8], mk[8];
        load_row(x + (int64_t)row * LN_D, lane, z);
#pragma unroll
        for (int i = 0; i < 8; ++i) mk[i] = 1.f;
        if (drop.thresh) {  // the forward normalised drop(x) + res: the same mask gates the gradient that reaches x
            const uint64_t e0 = (uint64_t)row * LN_D + lane * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) { mk[i] = drop_mult(drop, e0 + i); mk[4 + i] = drop_mult(drop, e0 + 128 + i); }
#pragma unroll
            for (int i = 0; i < 8; ++i) z[i] *= mk[i];
        }
        if (res) {
            float r[8];
            load_row(res + (int64_t)row * LN_D, lane, r);
#pragma unroll
            for (int i = 0; i < 8; ++i) z[i] += r[i];
        }
        load_row(dy + (int64_t)row * LN_D, lane, d);
        const float mu = mean[row], rs = rstd[row];
        float s1 = 0.f, s2 = 0.f;
        float xh[8], dg[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            xh[i] = (z[i] - mu) * rs;
            dg[i] = d[i] * g[i];
            s1 += dg[i];
            s2 += dg[i] * xh[i];
            ag[i] += d[i] * xh[i];
            ab[i] += d[i];
        }
        s1 = warp_sum(s1) * (1.f / LN_D);
        s2 = warp_sum(s2) * (1.f / LN_D);
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = rs * (dg[i] - s1 - xh[i] * s2);
        store_row(dz + (int64_t)row * LN_D, lane, o);  // gradient w.r.t. (drop(x) + res): what the residual branch receives
#pragma unroll
        for (int i = 0; i < 8; ++i) { o[i] *= mk[i]; az[i] += o[i]; }  // gradient w.r.t. x: operand copy and bias column sums
        if (dz_bf16) {  // GEMM-operand copy for the dgrad / wgrad that consume dz next
            __nv_bfloat16* zb = dz_bf16 + (int64_t)row * LN_D;
            __nv_bfloat162 p0 = __floats2bfloat162_rn(o[0], o[1]), p1 = __floats2bfloat162_rn(o[2], o[3]);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(o[4], o[5]), p3 = __floats2bfloat162_rn(o[6], o[7]);
            *reinterpret_cast<uint2*>(zb + lane * 4) = make_uint2(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1));
            *reinterpret_cast<uint2*>(zb + 128 + lane * 4) = make_uint2(*reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
        }
    }
    store_row(sg[warp], lane, ag);
    store_row(sb[warp], lane, ab);
    if (dbias) store_row(sz[warp], lane, az);
    __syncthreads();
    const int c = threadIdx.x;  // 256 threads == 256 columns
    float tg = 0.f, tb = 0.f;
#pragma unroll
    for (int w = 0; w < LN_ROWS_PER_BLOCK; ++w) { tg += sg[w][c]; tb += sb[w][c]; }
    atomicAdd(dgamma + c, tg);
    atomicAdd(dbeta + c, tb);
    if (dbias) {  // column sums of dz = gradient of the bias of the Linear that produced x
        float tz = 0.f;
#pragma unroll
        for (int w = 0; w < LN_ROWS_PER_BLOCK; ++w) tz += sz[w][c];
        atomicAdd(dbias + c, tz);
    }
}

}  // namespace stcat

using namespace stcat;

static int ln_fwd(const float* x, const float* res, const float* gamma, const float* beta, float* y, void* y_bf16, float* mean,
                  float* rstd, int rows, int d, float eps, const DropArgs& drop, void* stream) {
    STCAT_REQUIRE(x && gamma && beta && y && mean && rstd, STCAT_EINVAL, "layernorm_fwd: null pointer");
    STCAT_REQUIRE(d == LN_D, STCAT_ESHAPE, "layernorm_fwd: d=%d unsupported (HIDDEN must be 256)", d);
    STCAT_REQUIRE(rows >= 0, STCAT_EINVAL, "layernorm_fwd: rows=%d", rows);
    if (rows == 0) return 0;
    int blocks = (rows + LN_ROWS_PER_BLOCK - 1) / LN_ROWS_PER_BLOCK;
    int cap = num_sms() * 8;
    if (blocks > cap) blocks = cap;
    launch_pdl(layernorm_fwd_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, x, res, gamma, beta, y, (__nv_bfloat16*)y_bf16, mean,
               rstd, rows, eps, drop);
    return check_launch("layernorm_fwd_kernel");
}

extern "C" int stcat_layernorm_fwd(const float* x, const float* res, const float* gamma, const float* beta, float* y,
                                   void* y_bf16, float* mean, float* rstd, int rows, int d, float eps, void* stream) {
    return ln_fwd(x, res, gamma, beta, y, y_bf16, mean, rstd, rows, d, eps, DropArgs(), stream);
}

extern "C" int stcat_layernorm_dropout_fwd(const float* x, const float* res, const float* gamma, const float* beta, float* y,
                                           void* y_bf16, float* mean, float* rstd, int rows, int d, float eps, float p,
                                           uint64_t seed, uint64_t offset, void* stream) {
    STCAT_REQUIRE(p >= 0.f && p < 1.f, STCAT_EINVAL, "layernorm_dropout_fwd: p=%f", (double)p);
    return ln_fwd(x, res, gamma, beta, y, y_bf16, mean, rstd, rows, d, eps, make_drop(p, seed, offset), stream);
}

static int ln_bwd(const float* dy, const float* x, const float* res, const float* gamma, const float* mean, const float* rstd,
                  float* dz, void* dz_bf16, float* dgamma, float* dbeta, float* dbias, int rows, int d, const DropArgs& drop,
                  void* stream) {
    STCAT_REQUIRE(dy && x && gamma && mean && rstd && dz && dgamma && dbeta, STCAT_EINVAL, "layernorm_bwd: null pointer");
    STCAT_REQUIRE(d == LN_D, STCAT_ESHAPE, "layernorm_bwd: d=%d unsupported (HIDDEN must be 256)", d);
    if (rows <= 0) return rows == 0 ? 0 : set_err(STCAT_EINVAL, "layernorm_bwd: rows=%d", rows);
    int blocks = (rows + LN_ROWS_PER_BLOCK - 1) / LN_ROWS_PER_BLOCK;
    int cap = num_sms() * 2;
    if (blocks > cap) blocks = cap;
    launch_pdl(layernorm_bwd_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, dy, x, res, gamma, mean, rstd, dz,
               (__nv_bfloat16*)dz_bf16, dgamma, dbeta, dbias, rows, drop);
    return check_launch("layernorm_bwd_kernel");
}

extern "C" int stcat_layernorm_bwd(const float* dy, const float* x, const float* res, const float* gamma,
                                   const float* mean, const float* rstd, float* dz, void* dz_bf16, float* dgamma,
                                   float* dbeta, float* dbias, int rows, int d, void* stream) {
    return ln_bwd(dy, x, res, gamma, mean, rstd, dz, dz_bf16, dgamma, dbeta, dbias, rows, d, DropArgs(), stream);
}

extern "C" int stcat_layernorm_dropout_bwd(const float* dy, const float* x, const float* res, const float* gamma,
                                           const float* mean, const float* rstd, float* dz, void* dz_bf16, float* dgamma,
                                           float* dbeta, float* dbias, int rows, int d, float p, uint64_t seed,
                                           uint64_t offset, void* stream) {
    STCAT_REQUIRE(p >= 0.f && p < 1.f, STCAT_EINVAL, "layernorm_dropout_bwd: p=%f", (double)p);
    return ln_bwd(dy, x, res, gamma, mean, rstd, dz, dz_bf16, dgamma, dbeta, dbias, rows, d, make_drop(p, seed, offset), stream);
}
