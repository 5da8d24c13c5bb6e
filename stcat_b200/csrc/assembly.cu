// Layout glue either side of the encoder and between encoder and decoder, one launch each way (HBM-bound; 128-bit accesses
// on the token-major side, 32 x 64 shared-memory tiles for the [n, d, HW] <-> [n, HW, d] transposes).
//
//   token_assembly      (modal_encoder.py:40-72): X[f] = [frame_cls ; vis[f]^T ; text[:, v(f)]], POS[f] = [local_pos ; vpos[f]^T ; 0]
//                       plus the two GEMM-operand copies the first spatial layer needs (bf16(X + POS), bf16(X)).
//   token_assembly_bwd  : d vis = dX[:, 1:1+HW]^T, d text = sum over the frames of a video, d frame_cls = sum over all frames.
//   mem_operands        (query_decoder.py:83-96, 355-366, 633-639): the decoder's views of the encoder stream X [n, S, d]:
//                       bf16(X[:, 1:]), bf16(POS[:, 1:]), bf16(X[:, 1:] + POS[:, 1:]) as [n (S-1), d] and the fp32 CLS rows.
//   mem_operands_bwd    : dX = [g_cls ; g_mem + g_mempos].
//   template_*          (query_decoder.py:441-475): the FiLM template generator as two small kernels each way.
#include "common.cuh"

namespace stcat {

namespace {

__device__ __forceinline__ uint2 pack4(float a, float b, float c, float d) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    return make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
}
__device__ __forceinline__ float4 ld_as_f32x4(const void* p, int dtype, int64_t i4) {
    if (dtype == STCAT_F32) return reinterpret_cast<const float4*>(p)[i4];
    const uint2 u = reinterpret_cast<const uint2*>(p)[i4];
    const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&u.x), hi = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
    const float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
    return make_float4(a.x, a.y, b.x, b.y);
}

constexpr int TC = 64;  // channels per transpose tile
constexpr int TP = 32;  // positions per transpose tile

struct AsmArgs {
    const float* vis;     // [n, d, HW]
    const float* vpos;    // [n, d, HW]
    const float* text;    // [L, b, d]
    const int64_t* f2v;   // [n] video of frame f, or NULL (b == 1)
    const float* cls;     // [d] frame_cls
    const float* lpos;    // [d] local_pos_embed
    float* X;             // [n, S, d]
    float* POS;           // [n, S, d]
    __nv_bfloat16* qk_op; // [n, S, d] bf16(X + POS) or NULL
    __nv_bfloat16* x_op;  // [n, S, d] bf16(X) or NULL
    int n, d, HW, L, b, S;
    int ptiles, ctiles, vis_blocks;
};

__global__ void __launch_bounds__(256) token_assembly_kernel(AsmArgs a) {
    __shared__ float tv[TC][TP + 1];
    __shared__ float tp[TC][TP + 1];
    pdl_launch_dependents();
    pdl_wait();
    const int tid = threadIdx.x;
    if ((int)blockIdx.x < a.vis_blocks) {
        int bid = blockIdx.x;
        const int pt = bid % a.ptiles; bid /= a.ptiles;
        const int ct = bid % a.ctiles;
        const int f = bid / a.ctiles;
        const int p0 = pt * TP, c0 = ct * TC;
        {
            const int tx = tid & 31, ty = tid >> 5;  // tx: position, ty: channel row
            const int p = p0 + tx;
#pragma unroll
            for (int i = 0; i < TC / 8; ++i) {
                const int c = c0 + ty + 8 * i;
                float v = 0.f, q = 0.f;
                if (p < a.HW && c < a.d) {
                    const int64_t src = ((int64_t)f * a.d + c) * a.HW + p;
                    v = a.vis[src];
                    q = a.vpos[src];
                }
                tv[ty + 8 * i][tx] = v;
                tp[ty + 8 * i][tx] = q;
            }
        }
        __syncthreads();
        {
            const int tx = tid & 63, ty = tid >> 6;  // tx: channel, ty: position row
            const int c = c0 + tx;
#pragma unroll
            for (int i = 0; i < TP / 4; ++i) {
                const int pl = ty + 4 * i, p = p0 + pl;
                if (p < a.HW && c < a.d) {
                    const int64_t dst = ((int64_t)f * a.S + 1 + p) * a.d + c;
                    const float v = tv[tx][pl], q = tp[tx][pl];
                    a.X[dst] = v;
                    a.POS[dst] = q;
                    if (a.qk_op) a.qk_op[dst] = __float2bfloat16_rn(v + q);
                    if (a.x_op) a.x_op[dst] = __float2bfloat16_rn(v);
                }
            }
        }
        return;
    }
    // CLS row and the L text rows of every frame: 4 rows of d floats per block (d % 4 == 0)
    const int d4 = a.d >> 2;
    const int rows_per_frame = 1 + a.L;
    const int64_t total = (int64_t)a.n * rows_per_frame * d4;
    const int64_t first = ((int64_t)blockIdx.x - a.vis_blocks) * 256 + tid;
    const int64_t stride = ((int64_t)gridDim.x - a.vis_blocks) * 256;
    for (int64_t i = first; i < total; i += stride) {
        const int c4 = (int)(i % d4);
        const int64_t r = i / d4;
        const int rr = (int)(r % rows_per_frame);
        const int f = (int)(r / rows_per_frame);
        float4 v, q;
        int64_t dst;
        if (rr == 0) {
            v = reinterpret_cast<const float4*>(a.cls)[c4];
            q = reinterpret_cast<const float4*>(a.lpos)[c4];
            dst = ((int64_t)f * a.S) * d4 + c4;
        } else {
            const int l = rr - 1;
            const int64_t vdx = a.f2v ? a.f2v[f] : 0;
            v = reinterpret_cast<const float4*>(a.text)[((int64_t)l * a.b + vdx) * d4 + c4];
            q = make_float4(0.f, 0.f, 0.f, 0.f);
            dst = ((int64_t)f * a.S + 1 + a.HW + l) * d4 + c4;
        }
        reinterpret_cast<float4*>(a.X)[dst] = v;
        reinterpret_cast<float4*>(a.POS)[dst] = q;
        if (a.qk_op) reinterpret_cast<uint2*>(a.qk_op)[dst] = pack4(v.x + q.x, v.y + q.y, v.z + q.z, v.w + q.w);
        if (a.x_op) reinterpret_cast<uint2*>(a.x_op)[dst] = pack4(v.x, v.y, v.z, v.w);
    }
}

struct AsmBwdArgs {
    const float* dX;          // [n, S, d]
    float* dvis;              // [n, d, HW] or NULL
    float* dtext;             // [L, b, d] or NULL
    float* dcls;              // [d] or NULL
    const int64_t* vid_start; // [b + 1] first frame of every video (frames of a video are consecutive), or NULL (b == 1)
    int n, d, HW, L, b, S;
    int ptiles, ctiles, vis_blocks;
};

__global__ void __launch_bounds__(256) token_assembly_bwd_kernel(AsmBwdArgs a) {
    __shared__ float tv[TP][TC + 1];
    pdl_launch_dependents();
    pdl_wait();
    const int tid = threadIdx.x;
    if ((int)blockIdx.x < a.vis_blocks) {
        int bid = blockIdx.x;
        const int pt = bid % a.ptiles; bid /= a.ptiles;
        const int ct = bid % a.ctiles;
        const int f = bid / a.ctiles;
        const int p0 = pt * TP, c0 = ct * TC;
        {
            const int tx = tid & 63, ty = tid >> 6;
            const int c = c0 + tx;
#pragma unroll
            for (int i = 0; i < TP / 4; ++i) {
                const int pl = ty + 4 * i, p = p0 + pl;
                tv[pl][tx] = (p < a.HW && c < a.d) ? a.dX[((int64_t)f * a.S + 1 + p) * a.d + c] : 0.f;
            }
        }
        __syncthreads();
        {
            const int tx = tid & 31, ty = tid >> 5;
            const int p = p0 + tx;
#pragma unroll
            for (int i = 0; i < TC / 8; ++i) {
                const int cl = ty + 8 * i, c = c0 + cl;
                if (p < a.HW && c < a.d) a.dvis[((int64_t)f * a.d + c) * a.HW + p] = tv[tx][cl];
            }
        }
        return;
    }
    // reductions over frames, in frame order (deterministic): one float4 column group per thread.
    // rows: [0, b L) = (video v, text token l); row b L = the frame-CLS token (all frames)
    const int d4 = a.d >> 2;
    const int64_t nrows = (a.dtext ? (int64_t)a.b * a.L : 0) + (a.dcls ? 1 : 0);
    const int64_t total = nrows * d4;
    const int64_t first = ((int64_t)blockIdx.x - a.vis_blocks) * 256 + tid;
    const int64_t stride = ((int64_t)gridDim.x - a.vis_blocks) * 256;
    for (int64_t i = first; i < total; i += stride) {
        const int c4 = (int)(i % d4);
        int64_t r = i / d4;
        const bool is_cls = !a.dtext || r == (int64_t)a.b * a.L;
        int f0 = 0, f1 = a.n, row = 0, v = 0, l = 0;
        if (!is_cls) {
            v = (int)(r / a.L);
            l = (int)(r % a.L);
            row = 1 + a.HW + l;
            if (a.vid_start) { f0 = (int)a.vid_start[v]; f1 = (int)a.vid_start[v + 1]; }
        }
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 16
        for (int f = f0; f < f1; ++f) {
            const float4 g = reinterpret_cast<const float4*>(a.dX)[((int64_t)f * a.S + row) * d4 + c4];
            s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
        }
        if (is_cls) reinterpret_cast<float4*>(a.dcls)[c4] = s;
        else reinterpret_cast<float4*>(a.dtext)[((int64_t)l * a.b + v) * d4 + c4] = s;
    }
}

// ---- decoder memory operands ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mem_operands_kernel(const float* __restrict__ X, const float* __restrict__ POS, __nv_bfloat16* __restrict__ mem_op,
                    __nv_bfloat16* __restrict__ pos_op, __nv_bfloat16* __restrict__ mempos_op, float* __restrict__ cls,
                    int n, int S, int d4) {
    pdl_launch_dependents();
    pdl_wait();
    const int M = S - 1;
    const int64_t per_frame = (int64_t)M * d4;
    const int64_t n_mem = (int64_t)n * per_frame;
    const int64_t total = n_mem + (cls ? (int64_t)n * d4 : 0);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        if (i < n_mem) {
            const int64_t f = i / per_frame, r = i - f * per_frame;
            const int64_t src = (f * S + 1) * d4 + r;
            const float4 x = reinterpret_cast<const float4*>(X)[src];
            reinterpret_cast<uint2*>(mem_op)[i] = pack4(x.x, x.y, x.z, x.w);
            if (POS) {
                const float4 p = reinterpret_cast<const float4*>(POS)[src];
                if (pos_op) reinterpret_cast<uint2*>(pos_op)[i] = pack4(p.x, p.y, p.z, p.w);
                if (mempos_op) reinterpret_cast<uint2*>(mempos_op)[i] = pack4(x.x + p.x, x.y + p.y, x.z + p.z, x.w + p.w);
            }
        } else {
            const int64_t j = i - n_mem;
            const int64_t f = j / d4, c4 = j - f * d4;
            reinterpret_cast<float4*>(cls)[j] = reinterpret_cast<const float4*>(X)[f * S * d4 + c4];
        }
    }
}

__global__ void __launch_bounds__(256)
mem_operands_bwd_kernel(const void* __restrict__ g_mem, int dt_mem, const void* __restrict__ g_mempos, int dt_mempos,
                        const float* __restrict__ g_cls, float* __restrict__ dX, int n, int S, int d4) {
    pdl_launch_dependents();
    pdl_wait();
    const int M = S - 1;
    const int64_t per_frame = (int64_t)S * d4;
    const int64_t total = (int64_t)n * per_frame;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t f = i / per_frame, r = i - f * per_frame;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < d4) {
            if (g_cls) g = reinterpret_cast<const float4*>(g_cls)[f * d4 + r];
        } else {
            const int64_t src = f * (int64_t)M * d4 + (r - d4);
            if (g_mem) g = ld_as_f32x4(g_mem, dt_mem, src);
            if (g_mempos) {
                const float4 h = ld_as_f32x4(g_mempos, dt_mempos, src);
                g.x += h.x; g.y += h.y; g.z += h.z; g.w += h.w;
            }
        }
        reinterpret_cast<float4*>(dX)[i] = g;
    }
}

// ---- frame-CLS exchange with the temporal encoder layer (modal_encoder.py:170-195) ------------------------------------------
// gather : Y = [video ; X[:, r, :]] ([1 + n, d]) with the temporal layer's two GEMM operands bf16(Y + pos), bf16(Y) in one launch
// scatter: X[:, r, :] = Y[1:] in place, and the same rows of the stream's bf16 operand copy
__global__ void __launch_bounds__(256)
cls_gather_kernel(const float* __restrict__ X, const float* __restrict__ video, const float* __restrict__ pos, float* __restrict__ Y,
                  __nv_bfloat16* __restrict__ qk_op, __nv_bfloat16* __restrict__ y_op, int n, int S, int r, int d4) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t total = (int64_t)(1 + n) * d4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = i / d4, c4 = i - j * d4;
        const float4 v = j == 0 ? reinterpret_cast<const float4*>(video)[c4]
                                : reinterpret_cast<const float4*>(X)[((j - 1) * S + r) * d4 + c4];
        reinterpret_cast<float4*>(Y)[i] = v;
        if (qk_op) {
            const float4 p = reinterpret_cast<const float4*>(pos)[i];
            reinterpret_cast<uint2*>(qk_op)[i] = pack4(v.x + p.x, v.y + p.y, v.z + p.z, v.w + p.w);
        }
        if (y_op) reinterpret_cast<uint2*>(y_op)[i] = pack4(v.x, v.y, v.z, v.w);
    }
}
__global__ void __launch_bounds__(256)
cls_scatter_kernel(const float* __restrict__ Y, float* __restrict__ X, __nv_bfloat16* __restrict__ X_op,
                   __nv_bfloat16* __restrict__ qk_next, const float* __restrict__ pos, int n, int S, int r, int d4) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t total = (int64_t)n * d4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = i / d4, c4 = i - f * d4;
        const float4 v = reinterpret_cast<const float4*>(Y)[(1 + f) * d4 + c4];
        const int64_t dst = (f * S + r) * d4 + c4;
        reinterpret_cast<float4*>(X)[dst] = v;
        if (X_op) reinterpret_cast<uint2*>(X_op)[dst] = pack4(v.x, v.y, v.z, v.w);
        if (qk_next) {  // the next spatial layer's q/k operand bf16(X + POS) was written early from the stale rows: patch them
            const float4 p = reinterpret_cast<const float4*>(pos)[dst];
            reinterpret_cast<uint2*>(qk_next)[dst] = pack4(v.x + p.x, v.y + p.y, v.z + p.z, v.w + p.w);
        }
    }
}

// ---- template generator (query_decoder.py:441-475) ---------------------------------------------------------------------------
// film:   for the b video tokens: content = Wc v + bc, gamma = tanh(Wg v + bg), beta = tanh(Wb v + bb); v rounded to bf16 like
//         every GEMM operand, bf16 weights, fp32 accumulation.  One warp per output element.
// anchor: per frame f of video v: mod = gamma[v] * cls[f] + beta[v]; pq = Wa bf16(mod) + ba ([4]); anchor = sigmoid(pq).
__global__ void __launch_bounds__(256)
template_film_kernel(const float* __restrict__ vcls, const __nv_bfloat16* __restrict__ Wc, const float* __restrict__ bc,
                     const __nv_bfloat16* __restrict__ Wg, const float* __restrict__ bg, const __nv_bfloat16* __restrict__ Wb,
                     const float* __restrict__ bb, float* __restrict__ content, float* __restrict__ gamma,
                     float* __restrict__ beta, int b, int d) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t total = (int64_t)3 * b * d;
    if (warp >= total) return;
    const int which = (int)(warp / ((int64_t)b * d));
    const int64_t r = warp - (int64_t)which * b * d;
    const int v = (int)(r / d), j = (int)(r % d);
    const __nv_bfloat16* W = which == 0 ? Wc : (which == 1 ? Wg : Wb);
    const float* bias = which == 0 ? bc : (which == 1 ? bg : bb);
    float acc = 0.f;
    for (int k = lane; k < d; k += 32)
        acc += __bfloat162float(__float2bfloat16_rn(vcls[(int64_t)v * d + k])) * __bfloat162float(W[(int64_t)j * d + k]);
    acc = warp_sum(acc);
    if (lane == 0) {
        acc += bias[j];
        if (which == 0) content[(int64_t)v * d + j] = acc;
        else if (which == 1) gamma[(int64_t)v * d + j] = tanhf(acc);
        else beta[(int64_t)v * d + j] = tanhf(acc);
    }
}

__global__ void __launch_bounds__(256)
template_anchor_kernel(const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ cls,
                       const int64_t* __restrict__ f2v, const __nv_bfloat16* __restrict__ Wa, const float* __restrict__ ba,
                       const float* __restrict__ content, float* __restrict__ temp_query, __nv_bfloat16* __restrict__ mod_op,
                       float* __restrict__ anchor, int n, int d, int q) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int f = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (f >= n) return;
    const int64_t v = f2v ? f2v[f] : 0;
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = 0.f;
    for (int k = lane; k < d; k += 32) {
        const float m = gamma[v * d + k] * cls[(int64_t)f * d + k] + beta[v * d + k];
        const __nv_bfloat16 mb = __float2bfloat16_rn(m);
        mod_op[(int64_t)f * d + k] = mb;
        if (temp_query) temp_query[(int64_t)f * d + k] = content[v * d + k];
        const float mf = __bfloat162float(mb);
#pragma unroll
        for (int o = 0; o < 8; ++o)
            if (o < q) acc[o] += mf * __bfloat162float(Wa[(int64_t)o * d + k]);
    }
#pragma unroll
    for (int o = 0; o < 8; ++o)
        if (o < q) {
            const float s = warp_sum(acc[o]);
            if (lane == 0) anchor[(int64_t)f * q + o] = 1.f / (1.f + expf(-(s + ba[o])));
        }
}

// backward, stage 1 (one warp per frame): dpq = bf16(g_anchor a (1 - a)); dmod = dpq Wa; dcls[f] = dmod gamma[v];
// stage 2 (one thread per (video, channel), frames in order): dgamma = sum dmod cls, dbeta = sum dmod, dcontent = sum g_temp;
//          then through the tanh: dpre_g = dgamma (1 - gamma^2), dpre_b = dbeta (1 - beta^2)   -> dpre [3, b, d] (content, gamma, beta)
//          and dWa [q, d] (+)= dpq^T mod_op, dba (+)= colsum(dpq)
// stage 3: dv[v, k] = sum_j dpre_c[j] Wc[j, k] + dpre_g[j] Wg[j, k] + dpre_b[j] Wb[j, k];  dW*[j, k] (+)= sum_v bf16(dpre*[v, j]) bf16(v[v, k]),
//          db* (+)= sum_v dpre*[v, j]
__global__ void __launch_bounds__(256)
template_anchor_bwd_kernel(const float* __restrict__ g_anchor, const float* __restrict__ anchor, const float* __restrict__ gamma,
                           const int64_t* __restrict__ f2v, const __nv_bfloat16* __restrict__ Wa, __nv_bfloat16* __restrict__ dpq_op,
                           float* __restrict__ dmod, float* __restrict__ dcls, int n, int d, int q) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int f = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (f >= n) return;
    const int64_t v = f2v ? f2v[f] : 0;
    float dp[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        dp[o] = 0.f;
        if (o < q) {
            const float a = anchor[(int64_t)f * q + o];
            const __nv_bfloat16 r = __float2bfloat16_rn(g_anchor[(int64_t)f * q + o] * a * (1.f - a));
            if (lane == 0) dpq_op[(int64_t)f * q + o] = r;
            dp[o] = __bfloat162float(r);
        }
    }
    for (int k = lane; k < d; k += 32) {
        float m = 0.f;
#pragma unroll
        for (int o = 0; o < 8; ++o)
            if (o < q) m += dp[o] * __bfloat162float(Wa[(int64_t)o * d + k]);
        dmod[(int64_t)f * d + k] = m;
        dcls[(int64_t)f * d + k] = m * gamma[v * d + k];
    }
}

struct TplRedArgs {
    const float* dmod;        // [n, d]
    const float* cls;         // [n, d]
    const float* g_temp;      // [n, d] or NULL
    const float* gamma;       // [b, d]
    const float* beta;        // [b, d]
    const int64_t* vid_start; // [b + 1] or NULL (b == 1)
    const __nv_bfloat16* dpq_op;  // [n, q]
    const __nv_bfloat16* mod_op;  // [n, d]
    float* dpre;              // [3, b, d]
    float* dWa;               // [q, d] accumulated
    float* dba;               // [q] accumulated
    int n, d, q, b, kchunks;
};

// blocks [0, b kchunks): (video v, 32 channels): 8 frame groups x 32 channels, frames strided by 8, partial sums combined in
// frame-group order through shared memory (deterministic).  blocks [b kchunks, b kchunks + kchunks): dWa / dba over all frames.
__global__ void __launch_bounds__(256) template_reduce_bwd_kernel(TplRedArgs a) {
    __shared__ float red[8][8][33];
    pdl_launch_dependents();
    pdl_wait();
    const int kk = threadIdx.x & 31, fg = threadIdx.x >> 5;
    const int nb1 = a.b * a.kchunks;
    if ((int)blockIdx.x < nb1) {
        const int v = blockIdx.x / a.kchunks, k = (blockIdx.x % a.kchunks) * 32 + kk;
        int f0 = 0, f1 = a.n;
        if (a.vid_start) { f0 = (int)a.vid_start[v]; f1 = (int)a.vid_start[v + 1]; }
        float sg = 0.f, sb = 0.f, sc = 0.f;
        if (k < a.d) {
#pragma unroll 8
            for (int f = f0 + fg; f < f1; f += 8) {
                const float m = a.dmod[(int64_t)f * a.d + k];
                sg += m * a.cls[(int64_t)f * a.d + k];
                sb += m;
                if (a.g_temp) sc += a.g_temp[(int64_t)f * a.d + k];
            }
        }
        red[0][fg][kk] = sg; red[1][fg][kk] = sb; red[2][fg][kk] = sc;
        __syncthreads();
        if (fg == 0 && k < a.d) {
            sg = sb = sc = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) { sg += red[0][i][kk]; sb += red[1][i][kk]; sc += red[2][i][kk]; }
            const int64_t i = (int64_t)v * a.d + k, n_vk = (int64_t)a.b * a.d;
            const float g = a.gamma[i], bt = a.beta[i];
            a.dpre[i] = sc;
            a.dpre[n_vk + i] = sg * (1.f - g * g);
            a.dpre[2 * n_vk + i] = sb * (1.f - bt * bt);
        }
        return;
    }
    const int kc = blockIdx.x - nb1;
    const int k = kc * 32 + kk;
    float acc[8], accb[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = accb[o] = 0.f;
    if (k < a.d) {
#pragma unroll 4
        for (int f = fg; f < a.n; f += 8) {
            const float m = __bfloat162float(a.mod_op[(int64_t)f * a.d + k]);
#pragma unroll
            for (int o = 0; o < 8; ++o)
                if (o < a.q) {
                    const float dp = __bfloat162float(a.dpq_op[(int64_t)f * a.q + o]);
                    acc[o] += dp * m;
                    accb[o] += dp;
                }
        }
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) red[o][fg][kk] = acc[o];
    __syncthreads();
    if (fg == 0 && k < a.d) {
        for (int o = 0; o < a.q; ++o) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) s += red[o][i][kk];
            a.dWa[(int64_t)o * a.d + k] += s;
        }
    }
    if (kc == 0) {  // bias gradient: the same dpq column sums in every lane; lane 0 of every frame group publishes its part
        __syncthreads();
        if (kk == 0)
            for (int o = 0; o < 8; ++o) red[o][fg][0] = accb[o];
        __syncthreads();
        if (threadIdx.x < a.q) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) s += red[threadIdx.x][i][0];
            a.dba[threadIdx.x] += s;
        }
    }
}

struct TplWArgs {
    const float* dpre;   // [3, b, d]
    const float* vcls;   // [b, d]
    const __nv_bfloat16* W[3];  // content, gamma, beta projections [d, d]
    float* dW[3];        // [d, d] accumulated (may be NULL)
    float* db[3];        // [d] accumulated (may be NULL)
    float* dv;           // [b, d] written
    int b, d, w_blocks, kchunks;
};

// blocks [0, w_blocks): weight / bias gradients, thread = (which, j, k), k fastest.
// blocks [w_blocks, w_blocks + b kchunks): dv[v, 8 channels]: thread = (row group of 64, channel pair); the 3 d rows of
// [Wc ; Wg ; Wb] are strided by 64 (all loads of a thread in flight at once), partial sums combined in row-group order.
__global__ void __launch_bounds__(256) template_film_bwd_kernel(TplWArgs a) {
    __shared__ float red[64][9];
    pdl_launch_dependents();
    pdl_wait();
    const int d = a.d;
    if ((int)blockIdx.x < a.w_blocks) {
        const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= (int64_t)3 * d * d) return;
        const int which = (int)(i / ((int64_t)d * d));
        const int64_t r = i - (int64_t)which * d * d;
        const int j = (int)(r / d), k = (int)(r % d);
        if (a.dW[which]) {
            float s = 0.f;
            for (int v = 0; v < a.b; ++v)
                s += __bfloat162float(__float2bfloat16_rn(a.dpre[((int64_t)which * a.b + v) * d + j])) *
                     __bfloat162float(__float2bfloat16_rn(a.vcls[(int64_t)v * d + k]));
            a.dW[which][r] += s;
        }
        if (k == 0 && a.db[which]) {
            float s = 0.f;
            for (int v = 0; v < a.b; ++v) s += __bfloat162float(__float2bfloat16_rn(a.dpre[((int64_t)which * a.b + v) * d + j]));
            a.db[which][j] += s;
        }
        return;
    }
    const int bid = blockIdx.x - a.w_blocks;
    const int v = bid / a.kchunks, k0 = (bid % a.kchunks) * 8;
    const int rg = threadIdx.x >> 2, kq = threadIdx.x & 3;
    const int k = k0 + 2 * kq;
    float s0 = 0.f, s1 = 0.f;
    if (k < d) {  // d is even (d % 4 == 0 is required by the entry point)
#pragma unroll 4
        for (int row = rg; row < 3 * d; row += 64) {
            const int which = row / d, j = row - which * d;
            const float dp = __bfloat162float(__float2bfloat16_rn(a.dpre[((int64_t)which * a.b + v) * d + j]));
            const __nv_bfloat162 w2 = *reinterpret_cast<const __nv_bfloat162*>(a.W[which] + (int64_t)j * d + k);
            const float2 wf = __bfloat1622float2(w2);
            s0 += dp * wf.x;
            s1 += dp * wf.y;
        }
    }
    red[rg][2 * kq] = s0;
    red[rg][2 * kq + 1] = s1;
    __syncthreads();
    if (threadIdx.x < 8 && k0 + (int)threadIdx.x < d) {
        float s = 0.f;
#pragma unroll 8
        for (int i = 0; i < 64; ++i) s += red[i][threadIdx.x];
        a.dv[(int64_t)v * d + k0 + threadIdx.x] = s;
    }
}

// ---- box head + anchor refinement + sine embedding of the refined anchor (query_decoder.py:205-219, net_utils.py:29-63) ----
// One block per query row: delta = W3 h + b3 (q = 4 outputs), a' = sigmoid(delta + logit_clamped(anchor)), then the 512-wide sine
// embedding of a' (the next layer's positional input; a' is detached there) -- three dependent launches of the anchor-update
// chain between two box-decoder layers as one.
__device__ __forceinline__ float logit_clamped_(float x, float eps) {
    x = fminf(fmaxf(x, 0.f), 1.f);
    return logf(fmaxf(x, eps) / fmaxf(1.f - x, eps));
}
__global__ void __launch_bounds__(128)
box_head_kernel(const __nv_bfloat16* __restrict__ h, int64_t ldh, const __nv_bfloat16* __restrict__ W, const float* __restrict__ bias,
                const float* __restrict__ anchor, float* __restrict__ out, float* __restrict__ sine, __nv_bfloat16* __restrict__ sine_op,
                int R, int K, float eps) {
    __shared__ float red[4][4];
    pdl_launch_dependents();
    pdl_wait();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = blockIdx.x;  // one block of 4 warps per query row: the 512 sine / cosine evaluations are 4 per thread
    float pre[4];  // bias + inverse_sigmoid(anchor): loaded up front, off the dependent chain behind the reduction
#pragma unroll
    for (int o = 0; o < 4; ++o) pre[o] = bias[o] + logit_clamped_(anchor[(int64_t)r * 4 + o], eps);
    const float freq = powf(10000.f, (float)(2 * (tid >> 1)) / 128.f);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = tid * 2; k < K; k += 256) {  // K is even
        const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(h + (int64_t)r * ldh + k));
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const float2 w = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(W + (int64_t)o * K + k));
            acc[o] += x.x * w.x + x.y * w.y;
        }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        const float sum = warp_sum(acc[o]);
        if (lane == 0) red[warp][o] = sum;
    }
    __syncthreads();
    float a[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        const float z = ((red[0][o] + red[1][o]) + (red[2][o] + red[3][o])) + pre[o];
        a[o] = 1.f / (1.f + expf(-z));
    }
    if (tid < 4) out[(int64_t)r * 4 + tid] = a[tid];
    if (sine) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int j = tid + 128 * i, k = tid;  // j >> 7 == i: output block i = (y, x, w, h)[i] (anchor_sine_fwd_kernel)
            const float ac = i == 0 ? a[1] : (i == 1 ? a[0] : (i == 2 ? a[2] : a[3]));
            const float ph = ac * 6.283185307179586f / freq;
            const float v = (k & 1) ? cosf(ph) : sinf(ph);
            sine[(int64_t)r * 512 + j] = v;
            if (sine_op) sine_op[(int64_t)r * 512 + j] = __float2bfloat16_rn(v);
        }
    }
}

// backward of box_head_kernel up to the hidden activation: dz = g s (1 - s) (s = the refined anchor) -> dd_op = bf16(dz) [R, 4]
// (the operand of the W3 weight gradient), danchor (optional) through the clamped logit, and the data gradient of the last
// Linear with the ReLU mask of its input folded in: dh[r, k] = (sum_o dd_op[r, o] W3[o, k]) * (h[r, k] > 0), bf16.
__global__ void __launch_bounds__(128)
box_head_bwd_kernel(const float* __restrict__ g, const float* __restrict__ out, const float* __restrict__ anchor,
                    const __nv_bfloat16* __restrict__ W, const __nv_bfloat16* __restrict__ h, int64_t ldh, __nv_bfloat16* __restrict__ dd_op,
                    __nv_bfloat16* __restrict__ dh, float* __restrict__ danchor, int R, int K, float eps) {
    pdl_launch_dependents();
    pdl_wait();
    const int tid = threadIdx.x;
    const int r = blockIdx.x;
    float dd[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        const float sgm = out[(int64_t)r * 4 + o];
        const float dz = g[(int64_t)r * 4 + o] * sgm * (1.f - sgm);
        const __nv_bfloat16 q = __float2bfloat16_rn(dz);
        dd[o] = __bfloat162float(q);
        if (tid == o) {
            dd_op[(int64_t)r * 4 + o] = q;
            if (danchor) {
                const float x = anchor[(int64_t)r * 4 + o];
                float d = 0.f;
                if (x > 0.f && x < 1.f) {
                    if (x > eps) d += 1.f / x;
                    if (1.f - x > eps) d += 1.f / (1.f - x);
                }
                danchor[(int64_t)r * 4 + o] = dz * d;
            }
        }
    }
    for (int k = tid * 2; k < K; k += 256) {
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const float2 w = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(W + (int64_t)o * K + k));
            acc.x += dd[o] * w.x;
            acc.y += dd[o] * w.y;
        }
        const float2 hv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(h + (int64_t)r * ldh + k));
        if (!(hv.x > 0.f)) acc.x = 0.f;
        if (!(hv.y > 0.f)) acc.y = 0.f;
        *reinterpret_cast<__nv_bfloat162*>(dh + (int64_t)r * K + k) = __floats2bfloat162_rn(acc.x, acc.y);
    }
}

// out_bf16[r, c] = bf16(a[r, c] * b[r, c]) for c < cols (a has leading dimension lda: the first `cols` columns of the sine
// embedding); backward: db[r, c] = g[r, c] * a[r, c]
__global__ void __launch_bounds__(256)
mul_cast_kernel(const float* __restrict__ a, int64_t lda, const float* __restrict__ b, float* __restrict__ out_f32,
                __nv_bfloat16* __restrict__ out, const float* __restrict__ c_in, __nv_bfloat16* __restrict__ c_out, int64_t rows, int cols) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t total = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols;
        const int c = (int)(i - r * cols);
        const float v = a[r * lda + c] * b[i];
        if (out_f32) out_f32[i] = v;
        out[i] = __float2bfloat16_rn(v);
        if (c_out) c_out[i] = __float2bfloat16_rn(c_in[i]);  // a second operand copy riding in the same launch (query_pos)
    }
}
__global__ void __launch_bounds__(256)
mul_cast_bwd_kernel(const void* __restrict__ g, int g_dtype, const float* __restrict__ a, int64_t lda, float* __restrict__ db,
                    int64_t rows, int cols) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t total = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols;
        const int c = (int)(i - r * cols);
        const float gv = g_dtype == STCAT_F32 ? reinterpret_cast<const float*>(g)[i] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(g)[i]);
        db[i] = gv * a[r * lda + c];
    }
}

int grid_cap(int64_t blocks) {
    const int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace

}  // namespace stcat

using namespace stcat;

extern "C" {

STCAT_API int stcat_token_assembly(const float* vis, const float* vpos, const float* text, const int64_t* f2v, const float* frame_cls,
                                   const float* local_pos, float* X, float* POS, void* qk_op, void* x_op, int n, int d, int HW, int L,
                                   int b, void* stream) {
    STCAT_REQUIRE(vis && vpos && text && frame_cls && local_pos && X && POS, STCAT_EINVAL, "token_assembly: null pointer");
    STCAT_REQUIRE(n > 0 && HW > 0 && L >= 0 && b > 0 && d > 0 && d % 4 == 0, STCAT_ESHAPE, "token_assembly: bad shape n=%d d=%d HW=%d L=%d b=%d", n, d, HW, L, b);
    STCAT_REQUIRE(b == 1 || f2v, STCAT_EINVAL, "token_assembly: f2v is required for b > 1");
    AsmArgs a;
    a.vis = vis; a.vpos = vpos; a.text = text; a.f2v = b == 1 ? nullptr : f2v; a.cls = frame_cls; a.lpos = local_pos;
    a.X = X; a.POS = POS; a.qk_op = (__nv_bfloat16*)qk_op; a.x_op = (__nv_bfloat16*)x_op;
    a.n = n; a.d = d; a.HW = HW; a.L = L; a.b = b; a.S = 1 + HW + L;
    a.ptiles = (HW + TP - 1) / TP; a.ctiles = (d + TC - 1) / TC;
    a.vis_blocks = n * a.ptiles * a.ctiles;
    const int64_t row_items = (int64_t)n * (1 + L) * (d / 4);
    const int row_blocks = grid_cap((row_items + 255) / 256);
    cudaError_t e = launch_pdl(token_assembly_kernel, dim3(a.vis_blocks + row_blocks), dim3(256), 0, (cudaStream_t)stream, a);
    if (e != cudaSuccess) return set_err((int)e, "token_assembly: %s", cudaGetErrorString(e));
    return check_launch("token_assembly");
}

STCAT_API int stcat_token_assembly_bwd(const float* dX, float* dvis, float* dtext, float* dcls, const int64_t* vid_start, int n, int d,
                                       int HW, int L, int b, void* stream) {
    STCAT_REQUIRE(dX, STCAT_EINVAL, "token_assembly_bwd: null pointer");
    STCAT_REQUIRE(n > 0 && HW > 0 && L >= 0 && b > 0 && d > 0 && d % 4 == 0, STCAT_ESHAPE, "token_assembly_bwd: bad shape");
    STCAT_REQUIRE(b == 1 || vid_start, STCAT_EINVAL, "token_assembly_bwd: vid_start is required for b > 1");
    AsmBwdArgs a;
    a.dX = dX; a.dvis = dvis; a.dtext = L > 0 ? dtext : nullptr; a.dcls = dcls; a.vid_start = b == 1 ? nullptr : vid_start;
    a.n = n; a.d = d; a.HW = HW; a.L = L; a.b = b; a.S = 1 + HW + L;
    a.ptiles = (HW + TP - 1) / TP; a.ctiles = (d + TC - 1) / TC;
    a.vis_blocks = dvis ? n * a.ptiles * a.ctiles : 0;
    const int64_t nrows = (a.dtext ? (int64_t)b * L : 0) + (dcls ? 1 : 0);
    const int row_blocks = nrows ? grid_cap((nrows * (d / 4) + 255) / 256) : 0;
    if (a.vis_blocks + row_blocks == 0) return 0;
    cudaError_t e = launch_pdl(token_assembly_bwd_kernel, dim3(a.vis_blocks + row_blocks), dim3(256), 0, (cudaStream_t)stream, a);
    if (e != cudaSuccess) return set_err((int)e, "token_assembly_bwd: %s", cudaGetErrorString(e));
    return check_launch("token_assembly_bwd");
}

STCAT_API int stcat_mem_operands(const float* X, const float* POS, void* mem_op, void* pos_op, void* mempos_op, float* cls, int n, int S,
                                 int d, void* stream) {
    STCAT_REQUIRE(X && mem_op, STCAT_EINVAL, "mem_operands: null pointer");
    STCAT_REQUIRE(POS || (!pos_op && !mempos_op), STCAT_EINVAL, "mem_operands: pos_op / mempos_op need POS");
    STCAT_REQUIRE(n > 0 && S > 1 && d > 0 && d % 4 == 0, STCAT_ESHAPE, "mem_operands: bad shape n=%d S=%d d=%d", n, S, d);
    const int64_t items = (int64_t)n * S * (d / 4);
    cudaError_t e = launch_pdl(mem_operands_kernel, dim3(grid_cap((items + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, X, POS,
                               (__nv_bfloat16*)mem_op, (__nv_bfloat16*)pos_op, (__nv_bfloat16*)mempos_op, cls, n, S, d / 4);
    if (e != cudaSuccess) return set_err((int)e, "mem_operands: %s", cudaGetErrorString(e));
    return check_launch("mem_operands");
}

STCAT_API int stcat_mem_operands_bwd(const void* g_mem, int g_mem_dtype, const void* g_mempos, int g_mempos_dtype, const float* g_cls,
                                     float* dX, int n, int S, int d, void* stream) {
    STCAT_REQUIRE(dX, STCAT_EINVAL, "mem_operands_bwd: null pointer");
    STCAT_REQUIRE(n > 0 && S > 1 && d > 0 && d % 4 == 0, STCAT_ESHAPE, "mem_operands_bwd: bad shape n=%d S=%d d=%d", n, S, d);
    STCAT_REQUIRE((g_mem_dtype == STCAT_F32 || g_mem_dtype == STCAT_BF16) && (g_mempos_dtype == STCAT_F32 || g_mempos_dtype == STCAT_BF16),
                  STCAT_EINVAL, "mem_operands_bwd: gradients must be fp32 or bf16");
    const int64_t items = (int64_t)n * S * (d / 4);
    cudaError_t e = launch_pdl(mem_operands_bwd_kernel, dim3(grid_cap((items + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, g_mem,
                               g_mem_dtype, g_mempos, g_mempos_dtype, g_cls, dX, n, S, d / 4);
    if (e != cudaSuccess) return set_err((int)e, "mem_operands_bwd: %s", cudaGetErrorString(e));
    return check_launch("mem_operands_bwd");
}

STCAT_API int stcat_template_fwd(const float* videos_cls, const float* frames_cls, const int64_t* f2v, const void* Wc, const float* bc,
                                 const void* Wg, const float* bg, const void* Wb, const float* bb, const void* Wa, const float* ba,
                                 float* content, float* gamma, float* beta, void* mod_op, float* anchor, float* temp_query, int n, int b,
                                 int d, int q, void* stream) {
    STCAT_REQUIRE(videos_cls && frames_cls && Wc && bc && Wg && bg && Wb && bb && Wa && ba && content && gamma && beta && mod_op && anchor,
                  STCAT_EINVAL, "template_fwd: null pointer");
    STCAT_REQUIRE(n > 0 && b > 0 && d > 0 && q > 0 && q <= 8, STCAT_ESHAPE, "template_fwd: bad shape n=%d b=%d d=%d q=%d", n, b, d, q);
    STCAT_REQUIRE(b == 1 || f2v, STCAT_EINVAL, "template_fwd: f2v is required for b > 1");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t warps = (int64_t)3 * b * d;
    cudaError_t e = launch_pdl(template_film_kernel, dim3((unsigned)((warps + 7) / 8)), dim3(256), 0, st, videos_cls,
                               (const __nv_bfloat16*)Wc, bc, (const __nv_bfloat16*)Wg, bg, (const __nv_bfloat16*)Wb, bb, content, gamma,
                               beta, b, d);
    if (e != cudaSuccess) return set_err((int)e, "template_fwd(film): %s", cudaGetErrorString(e));
    e = launch_pdl(template_anchor_kernel, dim3((unsigned)((n + 7) / 8)), dim3(256), 0, st, (const float*)gamma, (const float*)beta,
                   frames_cls, b == 1 ? (const int64_t*)nullptr : f2v, (const __nv_bfloat16*)Wa, ba, (const float*)content, temp_query,
                   (__nv_bfloat16*)mod_op, anchor, n, d, q);
    if (e != cudaSuccess) return set_err((int)e, "template_fwd(anchor): %s", cudaGetErrorString(e));
    return check_launch("template_fwd");
}

STCAT_API int stcat_template_bwd(const float* g_anchor, const float* g_temp, const float* anchor, const float* videos_cls,
                                 const float* frames_cls, const int64_t* f2v, const int64_t* vid_start, const float* gamma,
                                 const float* beta, const void* mod_op, const void* Wc, const void* Wg, const void* Wb, const void* Wa,
                                 void* dpq_op, float* dmod, float* dpre, float* d_frames_cls, float* d_videos_cls, float* dWc, float* dbc,
                                 float* dWg, float* dbg, float* dWb, float* dbb, float* dWa, float* dba, int n, int b, int d, int q,
                                 void* stream) {
    STCAT_REQUIRE(g_anchor && anchor && videos_cls && frames_cls && gamma && beta && mod_op && Wc && Wg && Wb && Wa && dpq_op && dmod &&
                  dpre && d_frames_cls && d_videos_cls && dWa && dba, STCAT_EINVAL, "template_bwd: null pointer");
    STCAT_REQUIRE(n > 0 && b > 0 && d > 0 && d % 4 == 0 && q > 0 && q <= 8, STCAT_ESHAPE, "template_bwd: bad shape n=%d b=%d d=%d q=%d", n, b, d, q);
    STCAT_REQUIRE(b == 1 || (f2v && vid_start), STCAT_EINVAL, "template_bwd: f2v / vid_start are required for b > 1");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t* f2v_ = b == 1 ? nullptr : f2v;
    cudaError_t e = launch_pdl(template_anchor_bwd_kernel, dim3((unsigned)((n + 7) / 8)), dim3(256), 0, st, g_anchor, anchor, gamma, f2v_,
                               (const __nv_bfloat16*)Wa, (__nv_bfloat16*)dpq_op, dmod, d_frames_cls, n, d, q);
    if (e != cudaSuccess) return set_err((int)e, "template_bwd(anchor): %s", cudaGetErrorString(e));
    TplRedArgs r;
    r.dmod = dmod; r.cls = frames_cls; r.g_temp = g_temp; r.gamma = gamma; r.beta = beta; r.vid_start = b == 1 ? nullptr : vid_start;
    r.dpq_op = (const __nv_bfloat16*)dpq_op; r.mod_op = (const __nv_bfloat16*)mod_op; r.dpre = dpre; r.dWa = dWa; r.dba = dba;
    r.n = n; r.d = d; r.q = q; r.b = b; r.kchunks = (d + 31) / 32;
    e = launch_pdl(template_reduce_bwd_kernel, dim3((unsigned)(b * r.kchunks + r.kchunks)), dim3(256), 0, st, r);
    if (e != cudaSuccess) return set_err((int)e, "template_bwd(reduce): %s", cudaGetErrorString(e));
    TplWArgs w;
    w.dpre = dpre; w.vcls = videos_cls;
    w.W[0] = (const __nv_bfloat16*)Wc; w.W[1] = (const __nv_bfloat16*)Wg; w.W[2] = (const __nv_bfloat16*)Wb;
    w.dW[0] = dWc; w.dW[1] = dWg; w.dW[2] = dWb; w.db[0] = dbc; w.db[1] = dbg; w.db[2] = dbb;
    w.dv = d_videos_cls; w.b = b; w.d = d;
    w.w_blocks = (dWc || dWg || dWb || dbc || dbg || dbb) ? (int)(((int64_t)3 * d * d + 255) / 256) : 0;
    w.kchunks = (d + 7) / 8;
    e = launch_pdl(template_film_bwd_kernel, dim3((unsigned)(w.w_blocks + b * w.kchunks)), dim3(256), 0, st, w);
    if (e != cudaSuccess) return set_err((int)e, "template_bwd(film): %s", cudaGetErrorString(e));
    return check_launch("template_bwd");
}

STCAT_API int stcat_box_head_fwd(const void* h, int64_t ldh, const void* W, const float* bias, const float* anchor, float* out, float* sine,
                                 void* sine_op, int R, int K, float eps, void* stream) {
    STCAT_REQUIRE(h && W && bias && anchor && out, STCAT_EINVAL, "box_head_fwd: null pointer");
    STCAT_REQUIRE(R >= 0 && K > 0 && K % 2 == 0 && ldh >= K && ldh % 2 == 0, STCAT_ESHAPE, "box_head_fwd: bad shape R=%d K=%d ldh=%lld", R, K, (long long)ldh);
    STCAT_REQUIRE(sine || !sine_op, STCAT_EINVAL, "box_head_fwd: sine_op needs sine");
    if (R == 0) return 0;
    cudaError_t e = launch_pdl(box_head_kernel, dim3((unsigned)R), dim3(128), 0, (cudaStream_t)stream, (const __nv_bfloat16*)h, ldh,
                               (const __nv_bfloat16*)W, bias, anchor, out, sine, (__nv_bfloat16*)sine_op, R, K, eps);
    if (e != cudaSuccess) return set_err((int)e, "box_head_fwd: %s", cudaGetErrorString(e));
    return check_launch("box_head_fwd");
}

STCAT_API int stcat_mul_cast(const float* a, int64_t lda, const float* b, float* out_f32, void* out_bf16, const float* c_in,
                             void* c_out_bf16, int64_t rows, int cols, void* stream) {
    STCAT_REQUIRE(a && b && out_bf16, STCAT_EINVAL, "mul_cast: null pointer");
    STCAT_REQUIRE(rows >= 0 && cols > 0 && lda >= cols, STCAT_ESHAPE, "mul_cast: bad shape");
    STCAT_REQUIRE(c_in || !c_out_bf16, STCAT_EINVAL, "mul_cast: c_out needs c_in");
    if (rows == 0) return 0;
    cudaError_t e = launch_pdl(mul_cast_kernel, dim3(grid_cap((rows * cols + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, a, lda, b,
                               out_f32, (__nv_bfloat16*)out_bf16, c_in, (__nv_bfloat16*)c_out_bf16, rows, cols);
    if (e != cudaSuccess) return set_err((int)e, "mul_cast: %s", cudaGetErrorString(e));
    return check_launch("mul_cast");
}

STCAT_API int stcat_mul_cast_bwd(const void* g, int g_dtype, const float* a, int64_t lda, float* db, int64_t rows, int cols, void* stream) {
    STCAT_REQUIRE(g && a && db, STCAT_EINVAL, "mul_cast_bwd: null pointer");
    STCAT_REQUIRE(rows >= 0 && cols > 0 && lda >= cols && (g_dtype == STCAT_F32 || g_dtype == STCAT_BF16), STCAT_ESHAPE, "mul_cast_bwd: bad shape / dtype");
    if (rows == 0) return 0;
    cudaError_t e = launch_pdl(mul_cast_bwd_kernel, dim3(grid_cap((rows * cols + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, g, g_dtype,
                               a, lda, db, rows, cols);
    if (e != cudaSuccess) return set_err((int)e, "mul_cast_bwd: %s", cudaGetErrorString(e));
    return check_launch("mul_cast_bwd");
}

STCAT_API int stcat_box_head_bwd(const float* g, const float* out, const float* anchor, const void* W, const void* h, int64_t ldh,
                                 void* dd_op, void* dh, float* danchor, int R, int K, float eps, void* stream) {
    STCAT_REQUIRE(g && out && anchor && W && h && dd_op && dh, STCAT_EINVAL, "box_head_bwd: null pointer");
    STCAT_REQUIRE(R >= 0 && K > 0 && K % 2 == 0 && ldh >= K && ldh % 2 == 0, STCAT_ESHAPE, "box_head_bwd: bad shape R=%d K=%d", R, K);
    if (R == 0) return 0;
    cudaError_t e = launch_pdl(box_head_bwd_kernel, dim3((unsigned)R), dim3(128), 0, (cudaStream_t)stream, g, out, anchor,
                               (const __nv_bfloat16*)W, (const __nv_bfloat16*)h, ldh, (__nv_bfloat16*)dd_op, (__nv_bfloat16*)dh, danchor, R, K, eps);
    if (e != cudaSuccess) return set_err((int)e, "box_head_bwd: %s", cudaGetErrorString(e));
    return check_launch("box_head_bwd");
}

STCAT_API int stcat_cls_gather(const float* X, const float* video, const float* pos, float* Y, void* qk_op, void* y_op, int n, int S, int r,
                               int d, void* stream) {
    STCAT_REQUIRE(X && video && Y, STCAT_EINVAL, "cls_gather: null pointer");
    STCAT_REQUIRE(pos || !qk_op, STCAT_EINVAL, "cls_gather: qk_op needs pos");
    STCAT_REQUIRE(n > 0 && S > 0 && r >= 0 && r < S && d > 0 && d % 4 == 0, STCAT_ESHAPE, "cls_gather: bad shape n=%d S=%d r=%d d=%d", n, S, r, d);
    const int64_t items = (int64_t)(1 + n) * (d / 4);
    cudaError_t e = launch_pdl(cls_gather_kernel, dim3(grid_cap((items + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, X, video, pos, Y,
                               (__nv_bfloat16*)qk_op, (__nv_bfloat16*)y_op, n, S, r, d / 4);
    if (e != cudaSuccess) return set_err((int)e, "cls_gather: %s", cudaGetErrorString(e));
    return check_launch("cls_gather");
}

STCAT_API int stcat_cls_scatter(const float* Y, float* X, void* X_op, void* qk_next, const float* pos, int n, int S, int r, int d,
                                void* stream) {
    STCAT_REQUIRE(Y && X, STCAT_EINVAL, "cls_scatter: null pointer");
    STCAT_REQUIRE(pos || !qk_next, STCAT_EINVAL, "cls_scatter: qk_next needs pos");
    STCAT_REQUIRE(n > 0 && S > 0 && r >= 0 && r < S && d > 0 && d % 4 == 0, STCAT_ESHAPE, "cls_scatter: bad shape n=%d S=%d r=%d d=%d", n, S, r, d);
    const int64_t items = (int64_t)n * (d / 4);
    cudaError_t e = launch_pdl(cls_scatter_kernel, dim3(grid_cap((items + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, Y, X,
                               (__nv_bfloat16*)X_op, (__nv_bfloat16*)qk_next, pos, n, S, r, d / 4);
    if (e != cudaSuccess) return set_err((int)e, "cls_scatter: %s", cudaGetErrorString(e));
    return check_launch("cls_scatter");
}

}  // extern "C"
