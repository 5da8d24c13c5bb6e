// Microbenchmark (diagnostics, not on the product path): cost per tcgen05.mma of the shapes / operand layouts the attention
// kernels issue (N = 32 accumulate chains with MN-major operands), one CTA, one issuing thread, operands = zeros in smem.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I stcat_b200/csrc -I include scripts/mma_microbench.cu -o scripts/bin/mma_microbench -lcuda
#include "tc_common.cuh"
#include <cstdio>
using namespace stcat::tc;

template <int N, int A_MN, int B_MN, int KIND, int NACC, int REPS>
__device__ __forceinline__ void run_variant(uint32_t tmem, uint32_t sA, uint32_t sB, uint32_t bar, uint32_t& phase, long long* out) {
    const uint32_t idesc = make_idesc(128, N, A_MN, B_MN);
    for (int rep = 0; rep < 3; ++rep) {
        const long long t0 = clock64();
#pragma unroll
        for (int t = 0; t < REPS; ++t) {
            constexpr int dummy = 0; (void)dummy;
            const int tt = t & 15;
            uint64_t ad, bd;
            if (KIND == 0) {
                ad = make_desc(sA + (tt >> 2) * 16384 + (tt & 3) * 32, 16, 1024, LAYOUT_SW128);
                bd = make_desc(sB + tt * 1024, 512, 512, LAYOUT_SW64);
            } else if (KIND == 1) {
                ad = make_desc(sA + (tt & 7) * 2048, 16384, 1024, LAYOUT_SW128);
                bd = make_desc(sB + tt * 1024, 512, 512, LAYOUT_SW64);
            } else if (KIND == 2) {
                ad = make_desc(sA + (tt & 1) * 32, 16, 512, LAYOUT_SW64);
                bd = make_desc(sB + (tt & 1) * 32, 16, 512, LAYOUT_SW64);
            } else {
                ad = make_desc(sA + (tt & 3) * 32, 16, 1024, LAYOUT_SW128);
                bd = make_desc(sB + (tt & 3) * 32, 16, 1024, LAYOUT_SW128);
            }
            umma_bf16(tmem + (t % NACC) * ((N + 31) & ~31), ad, bd, idesc, t >= NACC ? 1u : 0u);
        }
        const long long t1 = clock64();
        umma_commit(bar);
        mbar_wait(bar, phase & 1);
        ++phase;
        const long long t2 = clock64();
        if (rep == 2) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
}

#define VARIANTS(X) \
    X(32, 0, 1, 0, 1, 16) X(32, 0, 1, 0, 1, 32) X(32, 0, 1, 0, 2, 32) X(32, 0, 1, 0, 4, 32) \
    X(32, 1, 1, 1, 1, 32) X(32, 1, 1, 1, 2, 32) X(32, 1, 1, 1, 4, 32) \
    X(64, 0, 1, 0, 1, 32) X(128, 0, 1, 0, 1, 32) X(256, 0, 1, 0, 1, 32) \
    X(224, 0, 0, 2, 1, 32) X(224, 0, 0, 2, 2, 32) X(128, 0, 0, 2, 1, 32) X(128, 0, 0, 2, 2, 32) \
    X(32, 0, 0, 3, 1, 32) X(32, 0, 0, 3, 4, 32) X(64, 0, 0, 3, 1, 32) X(128, 0, 0, 3, 1, 32) X(256, 0, 0, 3, 1, 32) X(256, 0, 0, 3, 2, 32) \
    X(16, 0, 1, 0, 1, 32) X(16, 0, 1, 0, 4, 32) X(96, 0, 1, 0, 1, 32)

__global__ void __launch_bounds__(128, 1) bench(long long* out) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const uint32_t bar = base + 200 * 1024, slot = bar + 16;
    for (int i = threadIdx.x; i < 200 * 1024 / 16; i += 128)
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + i * 16), "r"(0) : "memory");
    if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
    const uint32_t sA = base, sB = base + 128 * 1024;  // A region 128 KB, B region 64 KB
    if (threadIdx.x == 0) {
        uint32_t phase = 0;
        int v = 0;
#define X(N, AM, BM, KIND, NACC, REPS) run_variant<N, AM, BM, KIND, NACC, REPS>(tmem, sA, sB, bar, phase, out + 2 * (v++));
        VARIANTS(X)
#undef X
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    }
}

struct Variant { int N; int a_mn, b_mn; int kind; int nacc; int reps; };
int main() {
    Variant h[] = {
#define X(N, AM, BM, KIND, NACC, REPS) {N, AM, BM, KIND, NACC, REPS},
        VARIANTS(X)
#undef X
    };
    const int nv = sizeof(h) / sizeof(h[0]);
    long long* o;
    cudaMalloc(&o, nv * 16);
    const int smem = 202 * 1024;
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    bench<<<1, 128, smem>>>(o);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long r[2 * 64];
    cudaMemcpy(r, o, nv * 16, cudaMemcpyDeviceToHost);
    const char* kn[] = {"A K-major SW128 / B MN-major SW64 (PV, dQ)", "A MN-major SW128 / B MN-major SW64 (dV, dK)", "A,B K-major SW64 (scores)", "A,B K-major SW128 (GEMM)"};
    printf("%-50s %4s %5s %5s %10s %10s %10s\n", "operands", "N", "nacc", "reps", "issue clk", "total clk", "clk/MMA");
    for (int v = 0; v < nv; ++v)
        printf("%-50s %4d %5d %5d %10lld %10lld %10.1f\n", kn[h[v].kind], h[v].N, h[v].nacc, h[v].reps, r[2 * v], r[2 * v + 1], (double)r[2 * v + 1] / h[v].reps);
    return 0;
}
