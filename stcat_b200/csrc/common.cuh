// Shared helpers for libstcat_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>

#include "../../include/stcat_b200.h"

namespace stcat {

// thread-local error message returned by stcat_last_error()
char* err_buf();
int set_err(int code, const char* fmt, ...);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_err((int)e, "%s: %s", what, cudaGetErrorString(e));
    return 0;
}

#define STCAT_REQUIRE(cond, code, ...)                    \
    do {                                                  \
        if (!(cond)) return stcat::set_err(code, __VA_ARGS__); \
    } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// element load/store with dtype dispatch (T = float or __nv_bfloat16)
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------------
// One step is a dependent chain of ~10^3 short kernels; at a kernel boundary the GPU otherwise drains the grid, then
// launches, then the next grid runs its prologue (barrier init, TMEM allocation, descriptor prefetch) before it touches
// any data.  With the launch attribute below the next grid is scheduled as soon as every CTA of the previous one has
// passed pdl_launch_dependents(), runs its prologue under the previous kernel's tail, and blocks in pdl_wait() until
// the previous grid has completed and its writes are visible.  pdl_wait() must precede every global-memory access.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
    static const bool on = getenv("STCAT_NO_PDL") == nullptr;
    return on;
}

// launch `kernel` (which calls pdl_wait() before its first global access) with programmatic stream serialization
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- train-mode dropout (nn.Dropout / F.dropout sites of the reference, modal_encoder.py:237-240, query_decoder.py:344,
// 431-436, 612, 653-658, attention.py:381) ---------------------------------------------------------------------------
// Counter-based: element `idx` of a site whose forward drew (seed, offset) is kept iff drop_bits24(seed, offset + idx)
// >= thresh = p * 2^24.  Stateless, so the backward regenerates the mask of the forward from the same (seed, offset) -- no
// mask tensors in HBM.  tests/emu_backend.py restates it bit for bit.
// The hash: z = idx + seed * golden (64-bit counter), folded to 32 bits (lo ^ hi * 0x9E3779B1) and finished with a 32-bit
// multiply-xorshift mixer (hash-prospector "lowbias32" family: 16 / 0x21f0aaad / 15 / 0x735a2d97 / 15): ~10 integer
// instructions per element.  Round 1 used a full splitmix64 (two 64-bit multiplies, ~25 instructions), which made the
// dropout instantiations of the tcgen05 attention kernels ALU-bound on the hash (forward 21 -> 55 us at T = 64 / S = 213).
__host__ __device__ __forceinline__ uint32_t drop_bits24(uint64_t seed, uint64_t idx) {
    const uint64_t z = idx + seed * 0x9E3779B97F4A7C15ull;
    uint32_t x = (uint32_t)z ^ ((uint32_t)(z >> 32) * 0x9E3779B1u);
    x ^= x >> 16;
    x *= 0x21f0aaadu;
    x ^= x >> 15;
    x *= 0x735a2d97u;
    x ^= x >> 15;
    return x >> 8;
}
struct DropArgs {            // thresh == 0: no dropout
    uint32_t thresh = 0;
    float scale = 1.f;       // 1 / keep probability
    uint64_t seed = 0, offset = 0;
    const uint64_t* step = nullptr;  // optional device-resident step counter (stcat_set_dropout_step): see drop_resolve
    // optional precomputed keep bits of an attention site (stcat_dropout_bits): row r of the [B*H*Lq, Lk] probability matrix owns
    // `wpr` 32-bit words, bit j of word w = keep(r * Lk + 32 w + j).  The tcgen05 attention kernels read them instead of hashing
    // (the hash is ~10 integer instructions per probability inside MUFU-bound loops); every other kernel ignores them.
    const uint32_t* bits = nullptr;
    int wpr = 0;
};
// Host state: the device counter every dropout site mixes into its seed.  With it, a CUDA-graph replay of a captured
// training step draws fresh masks (the captured step increments the counter once, after its backward pass); without it
// (NULL, the default) masks are a pure function of the (seed, offset) arguments.
const uint64_t* dropout_step_ptr();
void set_dropout_step_ptr(const uint64_t* p);
inline DropArgs make_drop(float p, uint64_t seed, uint64_t offset) {
    DropArgs d;
    if (p > 0.f) {
        double t = (double)p * 16777216.0;
        d.thresh = t >= 16777215.0 ? 16777215u : (uint32_t)t;
        d.scale = (float)(16777216.0 / (16777216.0 - (double)d.thresh));
        d.seed = seed;
        d.offset = offset;
        d.step = dropout_step_ptr();
    }
    return d;
}
// optional epilogue extras of the tensor-core GEMM (gemm_tcgen05.cu), used by stcat_linear_bwd_data / stcat_linear_dropout_fwd
struct GemmEpilogue {
    const void* relu_mask = nullptr;   // bf16 [M, N]: C is zeroed where relu_mask <= 0
    int64_t ld_mask = 0;
    float* colsum = nullptr;           // [N] fp32: accumulated with the column sums of the stored C
    float alpha = 1.f;                 // C is scaled by alpha before the mask (the 1 / keep factor of a dropout behind the ReLU)
    DropArgs drop;                     // train-mode dropout on C after bias / ReLU, element index row * N + col (contiguous C)
};

#ifdef __CUDACC__
// First statement of every kernel that draws masks: fold the current value of the step counter into the seed.
__device__ __forceinline__ DropArgs drop_resolve(DropArgs d) {
    if (d.thresh != 0 && d.step != nullptr) d.seed += __ldg(d.step) * 0xD1B54A32D192ED03ull;
    return d;
}
#endif
__device__ __forceinline__ float drop_mult(const DropArgs& d, uint64_t idx) {  // the mask as a multiplier: 1/keep or 0
    return drop_bits24(d.seed, d.offset + idx) >= d.thresh ? d.scale : 0.f;
}
__device__ __forceinline__ float drop_apply(const DropArgs& d, uint64_t idx, float v) {
    return drop_bits24(d.seed, d.offset + idx) >= d.thresh ? v * d.scale : 0.f;
}

// SMs the persistent kernels (tcgen05 GEMM, tcgen05 attention) size their grids for: all of them, or the cap set by
// stcat_set_sm_cap (capi.cu) while a collective that keeps CTAs resident runs next to them -- a persistent CTA that cannot
// become resident because an NCCL CTA holds its SM would keep its statically assigned tiles waiting until the collective ends.
extern int g_sm_cap;
inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return (g_sm_cap > 0 && g_sm_cap < n) ? g_sm_cap : n;
}

}  // namespace stcat
