// Microbenchmark (diagnostics): issue cost of the softmax inner loop of the attention kernels -- exp2 on the XU pipe, bf16
// packing by F2FP or integer ops, a degree-3 polynomial exp2 on the FMA pipe -- at 1, 2, 4 warps per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 scripts/xu_microbench.cu -o scripts/bin/xu_microbench
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ipack(float lo, float hi) {
    return __byte_perm(__float_as_uint(lo) + 0x8000u, __float_as_uint(hi) + 0x8000u, 0x7632);
}
__device__ __forceinline__ uint32_t fpack(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
// 2^x for x <= 0 on the FMA / ALU pipes: x = n + f, n = floor(x) (magic-number rounding), 2^f by a degree-3 minimax polynomial
// on [0, 1), exponent added with an integer add
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -126.f);
    const float t = x + 12582912.f;                 // 1.5 * 2^23: rounds x to the nearest integer in the low mantissa bits
    const float n = t - 12582912.f;
    const float f = x - n;                          // in [-0.5, 0.5]
    float p = fmaf(f, 0.0558263f, 0.2402265f);
    p = fmaf(p, f, 0.6931472f);
    p = fmaf(p, f, 1.0f);
    return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
}

template <int V>
__global__ void bench(float* out, long long* clk, int iters, float sc, float ms) {
    extern __shared__ uint32_t sm[];
    float r[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) r[e] = -(float)((threadIdx.x * 7 + e * 13) % 97) * 0.1f;
    float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
            float p0, p1, p2, p3;
            if (V == 5) {  // row-max pass
                s0 = fmaxf(s0, r[e]); s1 = fmaxf(s1, r[e + 1]); s2 = fmaxf(s2, r[e + 2]); s3 = fmaxf(s3, r[e + 3]);
                continue;
            }
            const float x0 = fmaf(r[e], sc, -ms), x1 = fmaf(r[e + 1], sc, -ms), x2 = fmaf(r[e + 2], sc, -ms), x3 = fmaf(r[e + 3], sc, -ms);
            if (V == 6) { p0 = ex2_poly(x0); p1 = ex2_poly(x1); p2 = ex2_poly(x2); p3 = ex2_poly(x3); }
            else if (V == 7 || V == 8) { p0 = ex2(x0); p1 = ex2_poly(x1); p2 = ex2(x2); p3 = ex2_poly(x3); }
            else { p0 = ex2(x0); p1 = ex2(x1); p2 = ex2(x2); p3 = ex2(x3); }
            s0 += p0; s1 += p1; s2 += p2; s3 += p3;
            if (V == 1 || V == 3 || V == 7 || V == 6) { pk[e >> 1] = fpack(p0, p1); pk[(e >> 1) + 1] = fpack(p2, p3); }
            else if (V == 2 || V == 4 || V == 8) { pk[e >> 1] = ipack(p0, p1); pk[(e >> 1) + 1] = ipack(p2, p3); }
            else { pk[e >> 1] = __float_as_uint(p0) ^ __float_as_uint(p1); pk[(e >> 1) + 1] = __float_as_uint(p2) ^ __float_as_uint(p3); }
        }
        if (V == 3 || V == 4) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"((uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x * 4 + (j ^ (threadIdx.x & 3))) * 16),
                             "r"(pk[4 * j]), "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3]) : "memory");
        } else if (V != 5) {
#pragma unroll
            for (int j = 0; j < 16; ++j) acc ^= pk[j];
        }
#pragma unroll
        for (int e = 0; e < 32; ++e) r[e] += 1e-3f * (float)(it & 1);  // keeps the loop from being hoisted
    }
    const long long t1 = clock64();
    if (threadIdx.x % 32 == 0) clk[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3 + __uint_as_float(acc & 0x3fffffff);
}

template <int V>
void run(const char* name, float* out, long long* clk) {
    const int iters = 256;
    for (int warps : {4, 8, 16}) {
        cudaFuncSetAttribute(bench<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        bench<V><<<148, warps * 32, 64 * 1024>>>(out, clk, iters, 0.25f, 1.0f);
        cudaDeviceSynchronize();
        long long h[16];
        cudaMemcpy(h, clk, sizeof(long long) * warps, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < warps; ++i) mx = h[i] > mx ? h[i] : mx;
        const double per_elem_warp = (double)mx / (iters * 32.0);           // clocks per warp-wide element
        const double per_smsp = per_elem_warp / (warps / 4.0);              // clocks per warp-element per SM sub-partition
        printf("%-44s warps/SMSP %d  clk per warp-element %6.2f  -> per SMSP %5.2f clk per 32 elements\n", name, warps / 4, per_elem_warp, per_smsp);
    }
}

int main() {
    float* out; long long* clk;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&clk, 148 * 16 * 8);
    run<0>("ex2 only (fma, ex2, add)", out, clk);
    run<1>("ex2 + F2FP pack", out, clk);
    run<2>("ex2 + integer pack", out, clk);
    run<3>("ex2 + F2FP pack + st.shared.v4", out, clk);
    run<4>("ex2 + integer pack + st.shared.v4", out, clk);
    run<5>("fmax only (row-max pass)", out, clk);
    run<6>("polynomial exp2 + F2FP pack", out, clk);
    run<7>("half ex2, half polynomial + F2FP pack", out, clk);
    run<8>("half ex2, half polynomial + integer pack", out, clk);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
