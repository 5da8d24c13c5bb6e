// Optimizer-side step for a contiguous fp32 range of the flat parameter buffer:
//   train_net.py:139-140   torch.nn.utils.clip_grad_norm_(model.parameters(), MAX_GRAD_NORM)
//   engine/optimizer.py:44-46  torch.optim.AdamW (decoupled weight decay, bias correction, no amsgrad)
//   engine/optimizer.py:5-22   w_ema = w_ema * decay + (1 - decay) * w
// plus the refresh of the bf16 GEMM-operand shadow of the weights, all in ONE pass over HBM: per element it reads
// p, g, m, v (+ ema) and writes p, m, v (+ ema, + bf16 shadow) = 38 B, against 3 passes of ~0.8 GB each done tensor by
// tensor from Python in the reference.  Pure HBM-bound element-wise work: 128-bit accesses, grid = 8 x SMs.
#include "common.cuh"
#include <math.h>

namespace stcat {

static int grid_cap(int64_t n, int per_thread) {
    int64_t g = (n + 256LL * per_thread - 1) / (256LL * per_thread);
    const int cap = num_sms() * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// accum += sum x^2 (one atomicAdd per block); the caller zeroes accum once per step and may add the contributions of
// parameters that live outside the flat buffer before the step kernel reads it.
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ accum) {
    __shared__ float red[8];
    float s = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) ? n / 4 : 0;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 a = __ldg(x4 + i);
        s += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    }
    for (int64_t i = n4 * 4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) s += x[i] * x[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
        t = warp_sum(t);
        if (threadIdx.x == 0) atomicAdd(accum, t);
    }
}

struct AdamArgs {
    float lr, beta1, beta2, eps, weight_decay;
    float bias_c1, bias_c2_sqrt;  // 1 - beta1^step, sqrt(1 - beta2^step)
    float max_norm;               // <= 0: no clipping
    float ema_decay;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamArgs& a, float clip) {
    g *= clip;
    p *= 1.f - a.lr * a.weight_decay;                     // decoupled weight decay
    m = m + (g - m) * (1.f - a.beta1);                    // exp_avg.lerp_(grad, 1 - beta1)
    v = v * a.beta2 + (1.f - a.beta2) * g * g;            // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    const float denom = sqrtf(v) / a.bias_c2_sqrt + a.eps;
    p -= (a.lr / a.bias_c1) * (m / denom);
}

__global__ void __launch_bounds__(256)
adamw_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                  float* __restrict__ ema, __nv_bfloat16* __restrict__ shadow, int64_t n, const AdamArgs a,
                  const float* __restrict__ total_sumsq) {
    float clip = 1.f;
    if (a.max_norm > 0.f && total_sumsq != nullptr) {  // clip_grad_norm_: coef = max_norm / (norm + 1e-6), clamped to 1
        const float c = a.max_norm / (sqrtf(__ldg(total_sumsq)) + 1e-6f);
        clip = c < 1.f ? c : 1.f;
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(ema)) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(shadow) & 7) == 0;
    const int64_t n4 = vec ? n / 4 : 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
        adam_one(pp.x, gg.x, mm.x, vv.x, a, clip);
        adam_one(pp.y, gg.y, mm.y, vv.y, a, clip);
        adam_one(pp.z, gg.z, mm.z, vv.z, a, clip);
        adam_one(pp.w, gg.w, mm.w, vv.w, a, clip);
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
        if (ema != nullptr) {
            float4 ee = reinterpret_cast<float4*>(ema)[i];
            ee.x = ee.x * a.ema_decay + (1.f - a.ema_decay) * pp.x;
            ee.y = ee.y * a.ema_decay + (1.f - a.ema_decay) * pp.y;
            ee.z = ee.z * a.ema_decay + (1.f - a.ema_decay) * pp.z;
            ee.w = ee.w * a.ema_decay + (1.f - a.ema_decay) * pp.w;
            reinterpret_cast<float4*>(ema)[i] = ee;
        }
        if (shadow != nullptr) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(pp.x, pp.y), hi = __floats2bfloat162_rn(pp.z, pp.w);
            uint2 w;
            w.x = *reinterpret_cast<uint32_t*>(&lo);
            w.y = *reinterpret_cast<uint32_t*>(&hi);
            reinterpret_cast<uint2*>(shadow)[i] = w;
        }
    }
    for (int64_t i = n4 * 4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        float pp = p[i], mm = m[i], vv = v[i];
        adam_one(pp, g[i], mm, vv, a, clip);
        p[i] = pp; m[i] = mm; v[i] = vv;
        if (ema != nullptr) ema[i] = ema[i] * a.ema_decay + (1.f - a.ema_decay) * pp;
        if (shadow != nullptr) shadow[i] = __float2bfloat16_rn(pp);
    }
}

}  // namespace stcat

using namespace stcat;

extern "C" int stcat_sumsq(const float* x, int64_t n, float* accum, void* stream) {
    STCAT_REQUIRE(x && accum && n >= 0, STCAT_EINVAL, "sumsq: bad arguments");
    if (n == 0) return 0;
    sumsq_kernel<<<grid_cap(n, 16), 256, 0, (cudaStream_t)stream>>>(x, n, accum);
    return check_launch("sumsq_kernel");
}

extern "C" int stcat_adamw_step(float* p, const float* g, float* m, float* v, float* ema, void* shadow_bf16, int64_t n, float lr,
                                float beta1, float beta2, float eps, float weight_decay, int64_t step, const float* total_sumsq,
                                float max_norm, float ema_decay, void* stream) {
    STCAT_REQUIRE(p && g && m && v && n >= 0, STCAT_EINVAL, "adamw_step: null pointer");
    STCAT_REQUIRE(step >= 1 && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f, STCAT_EINVAL,
                  "adamw_step: step=%lld beta1=%f beta2=%f", (long long)step, (double)beta1, (double)beta2);
    STCAT_REQUIRE(max_norm <= 0.f || total_sumsq, STCAT_EINVAL, "adamw_step: clipping needs the gradient sum of squares");
    if (n == 0) return 0;
    AdamArgs a;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
    a.bias_c1 = (float)(1.0 - pow((double)beta1, (double)step));
    a.bias_c2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    a.max_norm = max_norm;
    a.ema_decay = ema_decay;
    adamw_step_kernel<<<grid_cap(n, 8), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, ema, (__nv_bfloat16*)shadow_bf16, n, a, total_sumsq);
    return check_launch("adamw_step_kernel");
}
