// Exact-fp32 SIMT GEMM family behind stcat_linear_{fwd,bwd_data,bwd_weight} for fp32 operands.
//
// This is the *correctness* path: plain FFMA, fp32 accumulate, no TF32, so the
// result matches the fp32 reference to summation-order rounding.  The production path for the same
// entry points is the tcgen05/TMA bf16 kernel in gemm_tcgen05.cu; this kernel is also its on-device
// checker.  C[m,n] = sum_k A(m,k) * B(n,k) with fully general element strides, which covers
//   fwd        A = x  (k contiguous)        B = w  (k contiguous)
//   bwd_data   A = dy (k = N contiguous)    B(k_out, n) = w[n, k_out]   (n index strided)
//   bwd_weight A(n, m) = dy[m, n]           B(k, m) = x[m, k]           (both "m-major")
#include "common.cuh"

namespace stcat {

constexpr int BM = 64, BN = 64, BK = 16;

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const TIn* __restrict__ A, int64_t sam, int64_t sak, const TIn* __restrict__ B, int64_t sbn,
                 int64_t sbk, TOut* __restrict__ C, int64_t ldc, const float* __restrict__ bias, int M, int N,
                 int K, int relu, int accumulate) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const bool a_kfast = (sak == 1);
    const bool b_kfast = (sbk == 1);
    for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int idx = tid + i * 256;
            int k, m;
            if (a_kfast) { k = idx & 15; m = idx >> 4; } else { m = idx & 63; k = idx >> 6; }
            int gm = m0 + m, gk = k0 + k;
            float v = 0.f;
            if (gm < M && gk < K) v = to_f32<TIn>(A[(int64_t)gm * sam + (int64_t)gk * sak]);
            As[k][m] = v;
            int n;
            if (b_kfast) { k = idx & 15; n = idx >> 4; } else { n = idx & 63; k = idx >> 6; }
            int gn = n0 + n;
            gk = k0 + k;
            v = 0.f;
            if (gn < N && gk < K) v = to_f32<TIn>(B[(int64_t)gn * sbn + (int64_t)gk * sbk]);
            Bs[k][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            float av[4] = {a.x, a.y, a.z, a.w};
            float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float v = acc[i][j];
            if (bias) v += bias[gn];
            TOut* p = C + (int64_t)gm * ldc + gn;
            if constexpr (sizeof(TOut) == 4) {
                // fp32 accumulation is atomic: with gradient fusion two streams may add into the same .grad
                if (accumulate && !relu) { atomicAdd(reinterpret_cast<float*>(p), v); continue; }
            }
            if (accumulate) v += to_f32<TOut>(*p);
            if (relu) v = fmaxf(v, 0.f);
            *p = from_f32<TOut>(v);
        }
    }
}

// Few-output GEMMs (the 4 / 2 / 1-wide prediction heads and anchor projection, pipeline.py:42-47, query_decoder.py:
// 446-449): one warp per output element, lanes stride the contraction index, shuffle reduction.  The 64 x 64 tile
// kernel above would run them on a single CTA with a serial K loop (15-20 us); this is launch-latency bound instead.
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
gemm_warpdot_kernel(const TIn* __restrict__ A, int64_t sam, int64_t sak, const TIn* __restrict__ B, int64_t sbn,
                    int64_t sbk, TOut* __restrict__ C, int64_t ldc, const float* __restrict__ bias, int M, int N,
                    int K, int relu, int accumulate) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= (int64_t)M * N) return;
    const int m = (int)(w / N), n = (int)(w % N);
    const TIn* a = A + (int64_t)m * sam;
    const TIn* b = B + (int64_t)n * sbn;
    float acc0 = 0.f, acc1 = 0.f;
    int k = lane;
    for (; k + 32 < K; k += 64) {
        acc0 = fmaf(to_f32<TIn>(a[(int64_t)k * sak]), to_f32<TIn>(b[(int64_t)k * sbk]), acc0);
        acc1 = fmaf(to_f32<TIn>(a[(int64_t)(k + 32) * sak]), to_f32<TIn>(b[(int64_t)(k + 32) * sbk]), acc1);
    }
    if (k < K) acc0 = fmaf(to_f32<TIn>(a[(int64_t)k * sak]), to_f32<TIn>(b[(int64_t)k * sbk]), acc0);
    float v = warp_sum(acc0 + acc1);
    if (lane != 0) return;
    if (bias) v += bias[n];
    TOut* p = C + (int64_t)m * ldc + n;
    if constexpr (sizeof(TOut) == 4) {
        if (accumulate && !relu) { atomicAdd(reinterpret_cast<float*>(p), v); return; }
    }
    if (accumulate) v += to_f32<TOut>(*p);
    if (relu) v = fmaxf(v, 0.f);
    *p = from_f32<TOut>(v);
}

// db[n] += sum_m dy[m, n]
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ dy, int64_t ld, float* __restrict__ db,
                                                     int M, int N, int rows_per_block) {
    __shared__ float red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + tx;
    const int r0 = blockIdx.y * rows_per_block;
    const int r1 = min(M, r0 + rows_per_block);
    float s = 0.f;
    if (n < N)
        for (int r = r0 + ty; r < r1; r += 8) s += to_f32<T>(dy[(int64_t)r * ld + n]);
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && n < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][tx];
        atomicAdd(db + n, t);
    }
}

template <typename TIn, typename TOut>
static int launch_gemm(const void* A, int64_t sam, int64_t sak, const void* B, int64_t sbn, int64_t sbk, void* C,
                       int64_t ldc, const float* bias, int M, int N, int K, int relu, int accumulate,
                       cudaStream_t st) {
    if ((int64_t)M * N <= 8192 && K >= 32) {  // few outputs: one warp each
        const int64_t warps = (int64_t)M * N;
        launch_pdl(gemm_warpdot_kernel<TIn, TOut>, dim3((unsigned)((warps + 7) / 8)), dim3(256), 0, st, (const TIn*)A, sam, sak, (const TIn*)B, sbn, sbk,
                   (TOut*)C, ldc, bias, M, N, K, relu, accumulate);
        return check_launch("gemm_warpdot_kernel");
    }
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
    gemm_simt_kernel<TIn, TOut><<<grid, 256, 0, st>>>((const TIn*)A, sam, sak, (const TIn*)B, sbn, sbk, (TOut*)C,
                                                      ldc, bias, M, N, K, relu, accumulate);
    return check_launch("gemm_simt_kernel");
}

int gemm_simt(const void* A, int64_t sam, int64_t sak, const void* B, int64_t sbn, int64_t sbk, int in_dtype,
              void* C, int64_t ldc, int out_dtype, const float* bias, int M, int N, int K, int relu,
              int accumulate, cudaStream_t st) {
    if (in_dtype == STCAT_F32 && out_dtype == STCAT_F32)
        return launch_gemm<float, float>(A, sam, sak, B, sbn, sbk, C, ldc, bias, M, N, K, relu, accumulate, st);
    if (in_dtype == STCAT_BF16 && out_dtype == STCAT_F32)
        return launch_gemm<__nv_bfloat16, float>(A, sam, sak, B, sbn, sbk, C, ldc, bias, M, N, K, relu, accumulate, st);
    if (in_dtype == STCAT_BF16 && out_dtype == STCAT_BF16)
        return launch_gemm<__nv_bfloat16, __nv_bfloat16>(A, sam, sak, B, sbn, sbk, C, ldc, bias, M, N, K, relu,
                                                         accumulate, st);
    if (in_dtype == STCAT_F32 && out_dtype == STCAT_BF16)
        return launch_gemm<float, __nv_bfloat16>(A, sam, sak, B, sbn, sbk, C, ldc, bias, M, N, K, relu, accumulate, st);
    return set_err(STCAT_EINVAL, "gemm_simt: bad dtype %d/%d", in_dtype, out_dtype);
}

// bf16 column sums at HBM speed: a warp reads whole rows (16 B = 8 columns per lane, 256 columns per pass), a block
// of 8 warps walks 8 row streams, partial sums meet in shared memory, one atomicAdd per column per block.
__global__ void __launch_bounds__(256) colsum_bf16_wide_kernel(const __nv_bfloat16* __restrict__ dy, int64_t ld,
                                                               float* __restrict__ db, int M, int N, int rows_per_block) {
    __shared__ float red[8][256 + 8];
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 256 + lane * 8;  // this lane's 8 columns
    const int r0 = blockIdx.y * rows_per_block;
    const int r1 = min(M, r0 + rows_per_block);
    float acc[8] = {};
    if (c0 < N) {  // N % 8 == 0 is checked by the host
        for (int r = r0 + warp; r < r1; r += 8) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(dy + (int64_t)r * ld + c0));
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                acc[2 * e] += __uint_as_float(w[e] << 16);
                acc[2 * e + 1] += __uint_as_float(w[e] & 0xffff0000u);
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) red[warp][lane * 8 + e] = acc[e];
    __syncthreads();
    const int c = threadIdx.x;  // 256 threads == 256 columns of this slab
    const int gc = blockIdx.x * 256 + c;
    if (gc < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][c];
        atomicAdd(db + gc, t);
    }
}

// several (dy, db) pairs in one launch: blockIdx.y = pair, blockIdx.x = 32-column slab, all rows in one block
struct ColsumGroup { const void* dy[12]; int64_t ld[12]; float* db[12]; int M[12]; int N[12]; };
template <typename T>
__global__ void __launch_bounds__(256) colsum_group_kernel(const ColsumGroup g) {
    __shared__ float red[8][33];
    pdl_launch_dependents();
    pdl_wait();
    const int j = blockIdx.y;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + tx;
    const int M = g.M[j], N = g.N[j];
    if (blockIdx.x * 32 >= N) return;
    const T* dy = (const T*)g.dy[j];
    float s = 0.f;
    if (n < N)
        for (int r = ty; r < M; r += 8) s += to_f32<T>(dy[(int64_t)r * g.ld[j] + n]);
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && n < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][tx];
        atomicAdd(g.db[j] + n, t);
    }
}

int colsum(const void* dy, int64_t ld, int dtype, float* db, int M, int N, int accumulate, cudaStream_t st);

int colsum_group(const void* const* dy, const int64_t* ld, float* const* db, const int* M, const int* N, int njobs, int dtype,
                 cudaStream_t st) {
    int maxM = 0, maxN = 0;
    for (int j = 0; j < njobs; ++j) { maxM = M[j] > maxM ? M[j] : maxM; maxN = N[j] > maxN ? N[j] : maxN; }
    if (maxM > 2048) {  // long column sums: the row-parallel kernel, pair by pair
        for (int j = 0; j < njobs; ++j) {
            int rc = colsum(dy[j], ld[j], dtype, db[j], M[j], N[j], 1, st);
            if (rc) return rc;
        }
        return 0;
    }
    ColsumGroup g;
    for (int j = 0; j < njobs; ++j) { g.dy[j] = dy[j]; g.ld[j] = ld[j]; g.db[j] = db[j]; g.M[j] = M[j]; g.N[j] = N[j]; }
    dim3 grid((maxN + 31) / 32, njobs);
    if (dtype == STCAT_F32) launch_pdl(colsum_group_kernel<float>, grid, dim3(256), 0, st, g);
    else launch_pdl(colsum_group_kernel<__nv_bfloat16>, grid, dim3(256), 0, st, g);
    return check_launch("colsum_group_kernel");
}

int colsum(const void* dy, int64_t ld, int dtype, float* db, int M, int N, int accumulate, cudaStream_t st) {
    if (!accumulate) {
        cudaError_t e = cudaMemsetAsync(db, 0, sizeof(float) * N, st);
        if (e != cudaSuccess) return set_err((int)e, "colsum memset: %s", cudaGetErrorString(e));
    }
    if (dtype == STCAT_BF16 && N % 8 == 0 && ld % 8 == 0 && ((uintptr_t)dy & 15) == 0 && M >= 512) {
        // enough blocks to fill the machine, at least 32 rows each
        int rpb = (int)(((int64_t)M * ((N + 255) / 256) + num_sms() * 4 - 1) / (num_sms() * 4));
        rpb = rpb < 32 ? 32 : rpb;
        dim3 g((N + 255) / 256, (M + rpb - 1) / rpb);
        launch_pdl(colsum_bf16_wide_kernel, g, dim3(256), 0, st, (const __nv_bfloat16*)dy, ld, db, M, N, rpb);
        return check_launch("colsum_bf16_wide_kernel");
    }
    int rows_per_block = 512;
    dim3 grid((N + 31) / 32, (M + rows_per_block - 1) / rows_per_block);
    if (dtype == STCAT_F32)
        colsum_kernel<float><<<grid, 256, 0, st>>>((const float*)dy, ld, db, M, N, rows_per_block);
    else
        colsum_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)dy, ld, db, M, N, rows_per_block);
    return check_launch("colsum_kernel");
}

}  // namespace stcat
