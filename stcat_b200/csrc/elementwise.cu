// HBM-bound element-wise helpers: 128-bit loads/stores, grid sized in multiples of the SM count.
#include "common.cuh"

namespace stcat {

__device__ __forceinline__ uint2 pack_bf16x4(float a, float b, float c, float d) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    return make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
}

__global__ void __launch_bounds__(256)
add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
           __nv_bfloat16* __restrict__ out_bf16, int64_t n4, int64_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 x = reinterpret_cast<const float4*>(a)[i];
        float4 y = reinterpret_cast<const float4*>(b)[i];
        float4 z = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
        if (out) reinterpret_cast<float4*>(out)[i] = z;
        if (out_bf16) reinterpret_cast<uint2*>(out_bf16)[i] = pack_bf16x4(z.x, z.y, z.z, z.w);
    }
    // tail (n not a multiple of 4)
    for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float z = a[i] + b[i];
        if (out) out[i] = z;
        if (out_bf16) out_bf16[i] = __float2bfloat16_rn(z);
    }
}

template <typename TY, typename TD>
__global__ void __launch_bounds__(256) relu_bwd_kernel(const TY* __restrict__ y, TD* __restrict__ dy, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        if (!(to_f32<TY>(y[i]) > 0.f)) dy[i] = from_f32<TD>(0.f);
}

__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t n4, int64_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v = reinterpret_cast<const float4*>(x)[i];
        reinterpret_cast<uint2*>(out)[i] = pack_bf16x4(v.x, v.y, v.z, v.w);
    }
    for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = __float2bfloat16_rn(x[i]);
}

// out[c, r] = bf16(x[r, c]); 32x32 tiles through shared memory so both sides are coalesced
__global__ void __launch_bounds__(256)
cast_bf16_transpose_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t rows, int64_t cols) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
    for (int i = ty; i < 32; i += 8) {
        int64_t r = r0 + i, c = c0 + tx;
        tile[i][tx] = (r < rows && c < cols) ? x[r * cols + c] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        int64_t c = c0 + i, r = r0 + tx;
        if (r < rows && c < cols) out[c * rows + r] = __float2bfloat16_rn(tile[tx][i]);
    }
}

static int grid_for(int64_t work_items) {
    int64_t blocks = (work_items + 255) / 256;
    int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// ---------------------------------------------------------------------------------------------
// sted scoring (post_processor.py:30-53): one block per video
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sted_score_kernel(const float* __restrict__ sted, const int32_t* __restrict__ durations, float* __restrict__ score,
                  int32_t* __restrict__ best, int t) {
    extern __shared__ float sm[];  // ls[t], le[t]
    __shared__ float red[8];
    __shared__ float redv[8];
    __shared__ int redi[8];
    float* ls = sm;
    float* le = sm + t;
    const int v = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* s = sted + (int64_t)v * t * 2;
    const int dur = durations[v];
    // log-softmax over ALL t positions (padded ones included, as in the reference), both channels
    for (int ch = 0; ch < 2; ++ch) {
        float mx = -INFINITY;
        for (int i = tid; i < t; i += 256) mx = fmaxf(mx, s[i * 2 + ch]);
        mx = warp_max(mx);
        if (lane == 0) red[warp] = mx;
        __syncthreads();
        mx = red[0];
        for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
        __syncthreads();
        float sum = 0.f;
        for (int i = tid; i < t; i += 256) sum += expf(s[i * 2 + ch] - mx);
        sum = warp_sum(sum);
        if (lane == 0) red[warp] = sum;
        __syncthreads();
        sum = 0.f;
        for (int w = 0; w < 8; ++w) sum += red[w];
        __syncthreads();
        const float lz = mx + logf(sum);
        float* dst = ch ? le : ls;
        for (int i = tid; i < t; i += 256) dst[i] = s[i * 2 + ch] - lz;
    }
    __syncthreads();
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    const int tt = t * t;
    for (int idx = tid; idx < tt; idx += 256) {
        int i = idx / t, j = idx - i * t;
        float pen = (j <= i || i >= dur || j >= dur) ? -1e32f : 0.f;
        float val = pen + (ls[i] + le[j]);  // same association as the reference: mask + (ls + le)
        if (score) score[(int64_t)v * tt + idx] = val;
        if (val > bv || (val == bv && idx < bi)) { bv = val; bi = idx; }
    }
    // block arg-max, first index on ties
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { redv[warp] = bv; redi[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 8; ++w)
            if (redv[w] > bv || (redv[w] == bv && redi[w] < bi)) { bv = redv[w]; bi = redi[w]; }
        best[v] = bi;
    }
}

// ---------------------------------------------------------------------------------------------
// map2d pooling (map2d_head.py:39-62): valid cell (i,j) = max_{i<=t<=j} x[B,t,c]; cascaded MaxPool1d
// in the reference is exactly this range maximum (max is exact, so results are bit-identical).
// One block per (B, i): thread = channel; walks j upward keeping a running max -> O(N) per row and
// coalesced reads of x[B, j, :]; writes map[B, c, i, j] (strided by N*N over c: the layout the
// reference's conv head expects).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
map2d_pool_kernel(const float* __restrict__ x, const uint8_t* __restrict__ valid, float* __restrict__ map, int N, int d) {
    // block = (map row i, batch B, 256-channel slab); thread = channel for the running max (coalesced
    // reads of x[B, j, c0:c0+256]); 32-column tiles are transposed through shared memory so that the
    // writes map[B, c, i, j0:j0+32] are 128-byte segments.
    __shared__ float tile[256][33];
    const int bidx = blockIdx.y, i = blockIdx.x, c0 = blockIdx.z * 256;
    const int c = c0 + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float run = -INFINITY;
    for (int j0 = 0; j0 < N; j0 += 32) {
#pragma unroll 4
        for (int jj = 0; jj < 32; ++jj) {
            const int j = j0 + jj;
            float out = 0.f;
            if (j < N && j >= i && c < d) {
                run = fmaxf(run, x[((int64_t)bidx * N + j) * d + c]);
                out = valid[i * N + j] ? run : 0.f;
            }
            tile[threadIdx.x][jj] = out;
        }
        __syncthreads();
        for (int cc = warp; cc < 256; cc += 8) {
            const int j = j0 + lane;
            if (c0 + cc < d && j < N) map[(((int64_t)bidx * d + c0 + cc) * N + i) * N + j] = tile[cc][lane];
        }
        __syncthreads();
    }
}


// dx[m, n] = 0 where y[m, n] <= 0 (strided 2-D form of relu_bwd; unfused tail of stcat_linear_bwd_data)
template <typename TY, typename TD>
__global__ void __launch_bounds__(256) relu_mask_2d_kernel(const TY* __restrict__ y, int64_t ldy, TD* __restrict__ dx,
                                                           int64_t lddx, int M, int N) {
    const int64_t total = (int64_t)M * N;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / N, n = i % N;
        if (!(to_f32<TY>(y[m * ldy + n]) > 0.f)) dx[m * lddx + n] = from_f32<TD>(0.f);
    }
}

int relu_mask_2d(const void* y, int64_t ldy, int y_dtype, void* dx, int64_t lddx, int dx_dtype, int M, int N, cudaStream_t st) {
    const int64_t total = (int64_t)M * N;
    int g = (int)((total + 255) / 256);
    const int cap = num_sms() * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    if (y_dtype == STCAT_F32 && dx_dtype == STCAT_F32)
        relu_mask_2d_kernel<float, float><<<g, 256, 0, st>>>((const float*)y, ldy, (float*)dx, lddx, M, N);
    else if (y_dtype == STCAT_BF16 && dx_dtype == STCAT_F32)
        relu_mask_2d_kernel<__nv_bfloat16, float><<<g, 256, 0, st>>>((const __nv_bfloat16*)y, ldy, (float*)dx, lddx, M, N);
    else if (y_dtype == STCAT_BF16 && dx_dtype == STCAT_BF16)
        relu_mask_2d_kernel<__nv_bfloat16, __nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)y, ldy, (__nv_bfloat16*)dx, lddx, M, N);
    else if (y_dtype == STCAT_F32 && dx_dtype == STCAT_BF16)
        relu_mask_2d_kernel<float, __nv_bfloat16><<<g, 256, 0, st>>>((const float*)y, ldy, (__nv_bfloat16*)dx, lddx, M, N);
    else
        return set_err(STCAT_EINVAL, "relu_mask_2d: bad dtype %d/%d", y_dtype, dx_dtype);
    return check_launch("relu_mask_2d_kernel");
}


// ---- box-decoder anchor glue (query_decoder.py:188-219, net_utils.py:29-63): one kernel each instead of ~8 ATen kernels ----
// sine embedding of the (cx, cy, w, h) anchors: out[n, 512] ordered (y, x, w, h), 128 dims per coordinate,
// e[k] = sin(2 pi c / 10000^(2 floor(k/2) / 128)) for even k, cos(...) for odd k; optional bf16 operand copy.
__global__ void __launch_bounds__(256) anchor_sine_fwd_kernel(const float* __restrict__ anchor, float* __restrict__ out,
                                                              __nv_bfloat16* __restrict__ out_bf16, int64_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t total = n * 512;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i >> 9;
        const int j = (int)(i & 511), blk = j >> 7, k = j & 127;
        const int c = blk == 0 ? 1 : (blk == 1 ? 0 : blk);  // output block 0 = y, 1 = x, 2 = w, 3 = h
        const float freq = powf(10000.f, (float)(2 * (k >> 1)) / 128.f);
        const float p = anchor[r * 4 + c] * 6.283185307179586f / freq;
        const float v = (k & 1) ? cosf(p) : sinf(p);
        out[i] = v;
        if (out_bf16) out_bf16[i] = __float2bfloat16_rn(v);
    }
}

// d anchor[r, c] = sum_k dy[r, blk(c), k] * d e_k / d c
__global__ void __launch_bounds__(128) anchor_sine_bwd_kernel(const float* __restrict__ anchor, const float* __restrict__ dy,
                                                              float* __restrict__ danchor, int64_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t rc = blockIdx.x;  // one block per (row, coordinate)
    const int64_t r = rc >> 2;
    const int c = (int)(rc & 3);
    const int blk = c == 0 ? 1 : (c == 1 ? 0 : c);
    const int k = threadIdx.x;
    const float freq = powf(10000.f, (float)(2 * (k >> 1)) / 128.f);
    const float w = 6.283185307179586f / freq;
    const float p = anchor[r * 4 + c] * w;
    const float g = dy[r * 512 + blk * 128 + k];
    float v = g * w * ((k & 1) ? -sinf(p) : cosf(p));
    v = warp_sum(v);
    __shared__ float red[4];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) danchor[r * 4 + c] = red[0] + red[1] + red[2] + red[3];
}

// out = sigmoid(delta + logit_clamped(anchor)), logit_clamped(x) = log(max(x', eps) / max(1 - x', eps)), x' = clamp(x, 0, 1)
__device__ __forceinline__ float logit_clamped(float x, float eps) {
    x = fminf(fmaxf(x, 0.f), 1.f);
    return logf(fmaxf(x, eps) / fmaxf(1.f - x, eps));
}
__global__ void __launch_bounds__(256) box_refine_fwd_kernel(const float* __restrict__ delta, const float* __restrict__ anchor,
                                                             float* __restrict__ out, int64_t n, float eps) {
    pdl_launch_dependents();
    pdl_wait();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float z = delta[i] + logit_clamped(anchor[i], eps);
        out[i] = 1.f / (1.f + expf(-z));
    }
}
__global__ void __launch_bounds__(256) box_refine_bwd_kernel(const float* __restrict__ out, const float* __restrict__ anchor,
                                                             const float* __restrict__ g, float* __restrict__ ddelta,
                                                             float* __restrict__ danchor, int64_t n, float eps) {
    pdl_launch_dependents();
    pdl_wait();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float s = out[i];
        const float dz = g[i] * s * (1.f - s);
        ddelta[i] = dz;
        if (danchor) {
            const float x = anchor[i];
            float d = 0.f;
            if (x > 0.f && x < 1.f) {  // clamp(x, 0, 1) passes the gradient strictly inside (at the ends torch gives it too; measure zero)
                if (x > eps) d += 1.f / x;
                if (1.f - x > eps) d += 1.f / (1.f - x);
            }
            danchor[i] = dz * d;
        }
    }
}


// out[i] = keep(i) ? x[i] / (1 - p) : 0  (forward and, with the same seed / offset on dy, backward of a dropout site)
template <typename T>
__global__ void __launch_bounds__(256) dropout_kernel(const T* __restrict__ x, T* __restrict__ out, int64_t n, const DropArgs d_in) {
    pdl_launch_dependents();
    pdl_wait();
    const DropArgs d = drop_resolve(d_in);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = from_f32<T>(drop_apply(d, (uint64_t)i, to_f32<T>(x[i])));
}

}  // namespace stcat

using namespace stcat;

// ---- image positional encoding (vision_model/position_encoding.py:70-94, PositionEmbeddingSine(128, normalize=True)) ----
// mask [n, H, W] (1 = padded) -> pos [n, H, W, 2F] channels-last (the layout the token assembly reads row by row):
// y_embed = cumsum_h(!mask) / (column total + 1e-6) * scale, x_embed likewise along w; channel c < F: y, else x;
// value = embed / T^(2 floor(k/2) / F), sin for even k, cos for odd k.  One block per pixel, one thread per (y|x, k).
__global__ void __launch_bounds__(256) pos_sine_kernel(const uint8_t* __restrict__ mask, float* __restrict__ out, int H, int W, int F,
                                                       float temperature, float scale) {
    pdl_launch_dependents();
    pdl_wait();
    const int pix = blockIdx.x, w = pix % W, h = (pix / W) % H, f = pix / (W * H);
    const uint8_t* m = mask + (int64_t)f * H * W;
    __shared__ float emb[2];
    if (threadIdx.x < 2) {
        float cum = 0.f, tot = 0.f;
        if (threadIdx.x == 0) {  // along h at column w
            for (int i = 0; i < H; ++i) { const float v = m[i * W + w] ? 0.f : 1.f; tot += v; if (i <= h) cum += v; }
        } else {                 // along w at row h
            for (int j = 0; j < W; ++j) { const float v = m[h * W + j] ? 0.f : 1.f; tot += v; if (j <= w) cum += v; }
        }
        emb[threadIdx.x] = cum / (tot + 1e-6f) * scale;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * F; c += blockDim.x) {
        const int k = c < F ? c : c - F;
        const float dim_t = powf(temperature, (float)(2 * (k >> 1)) / (float)F);
        const float p = emb[c < F ? 0 : 1] / dim_t;
        out[(int64_t)pix * 2 * F + c] = (k & 1) ? cosf(p) : sinf(p);
    }
}

// ---- box interpolation of the evaluation path (engine/evaluate.py:20-38 linear_interp): sampled frames carry predicted
// boxes, every frame between two sampled frames gets the linear blend, frames outside the sampled range are skipped (-1).
// frame_ids [m] ascending sampled frame indices, boxes [m, 4]; out [n_frames, 4] for frames first..first+n_frames-1.
__global__ void __launch_bounds__(256) box_interp_kernel(const int64_t* __restrict__ frame_ids, const float* __restrict__ boxes, int m,
                                                         float* __restrict__ out, int64_t first, int n_frames) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_frames) return;
    const int64_t fid = first + i;
    int lo = 0, hi = m - 1;  // largest index with frame_ids[idx] <= fid
    if (m == 0 || fid < frame_ids[0] || fid > frame_ids[m - 1]) {
        out[i * 4] = out[i * 4 + 1] = out[i * 4 + 2] = out[i * 4 + 3] = -1.f;
        return;
    }
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (frame_ids[mid] <= fid) lo = mid; else hi = mid - 1;
    }
    const int nxt = lo + 1 < m ? lo + 1 : lo;
    const float span = (float)(frame_ids[nxt] - frame_ids[lo]);
    const float t = span > 0.f ? (float)(fid - frame_ids[lo]) / span : 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) out[i * 4 + c] = boxes[lo * 4 + c] + t * (boxes[nxt * 4 + c] - boxes[lo * 4 + c]);
}

extern "C" int stcat_add(const float* a, const float* b, float* out, void* out_bf16, int64_t n, void* stream) {
    STCAT_REQUIRE(a && b && (out || out_bf16), STCAT_EINVAL, "add: null pointer");
    STCAT_REQUIRE(n >= 0, STCAT_EINVAL, "add: n<0");
    if (n == 0) return 0;
    STCAT_REQUIRE(((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0) && ((uintptr_t)out % 16 == 0) && ((uintptr_t)out_bf16 % 8 == 0),
                  STCAT_EALIGN, "add: pointers must be 16-byte aligned");
    launch_pdl(add_kernel, dim3(grid_for(n / 4 + 1)), dim3(256), 0, (cudaStream_t)stream, a, b, out, (__nv_bfloat16*)out_bf16, n / 4, n);
    return check_launch("add_kernel");
}

extern "C" int stcat_relu_bwd(const void* y, int y_dtype, void* dy, int dy_dtype, int64_t n, void* stream) {
    STCAT_REQUIRE(y && dy, STCAT_EINVAL, "relu_bwd: null pointer");
    if (n <= 0) return n == 0 ? 0 : set_err(STCAT_EINVAL, "relu_bwd: n<0");
    cudaStream_t st = (cudaStream_t)stream;
    int g = grid_for(n);
    if (y_dtype == STCAT_F32 && dy_dtype == STCAT_F32)
        relu_bwd_kernel<float, float><<<g, 256, 0, st>>>((const float*)y, (float*)dy, n);
    else if (y_dtype == STCAT_BF16 && dy_dtype == STCAT_F32)
        relu_bwd_kernel<__nv_bfloat16, float><<<g, 256, 0, st>>>((const __nv_bfloat16*)y, (float*)dy, n);
    else if (y_dtype == STCAT_BF16 && dy_dtype == STCAT_BF16)
        relu_bwd_kernel<__nv_bfloat16, __nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)y, (__nv_bfloat16*)dy, n);
    else
        return set_err(STCAT_EINVAL, "relu_bwd: bad dtype %d/%d", y_dtype, dy_dtype);
    return check_launch("relu_bwd_kernel");
}

extern "C" int stcat_cast_bf16(const float* x, void* out, int64_t rows, int64_t cols, int transpose, void* stream) {
    STCAT_REQUIRE(x && out, STCAT_EINVAL, "cast_bf16: null pointer");
    STCAT_REQUIRE(rows >= 0 && cols >= 0, STCAT_EINVAL, "cast_bf16: negative size");
    if (rows == 0 || cols == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (!transpose) {
        int64_t n = rows * cols;
        STCAT_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)out % 8 == 0), STCAT_EALIGN, "cast_bf16: alignment");
        launch_pdl(cast_bf16_kernel, dim3(grid_for(n / 4 + 1)), dim3(256), 0, st, x, (__nv_bfloat16*)out, n / 4, n);
        return check_launch("cast_bf16_kernel");
    }
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
    STCAT_REQUIRE(grid.y <= 65535, STCAT_ESHAPE, "cast_bf16: too many rows for transpose");
    cast_bf16_transpose_kernel<<<grid, 256, 0, st>>>(x, (__nv_bfloat16*)out, rows, cols);
    return check_launch("cast_bf16_transpose_kernel");
}

extern "C" int stcat_sted_score(const float* sted, const int32_t* durations, float* score, int32_t* best, int b, int t,
                                void* stream) {
    STCAT_REQUIRE(sted && durations && best, STCAT_EINVAL, "sted_score: null pointer");
    STCAT_REQUIRE(b >= 0 && t > 0 && t <= 4096, STCAT_ESHAPE, "sted_score: b=%d t=%d unsupported", b, t);
    if (b == 0) return 0;
    sted_score_kernel<<<b, 256, 2 * t * sizeof(float), (cudaStream_t)stream>>>(sted, durations, score, best, t);
    return check_launch("sted_score_kernel");
}

extern "C" int stcat_map2d_pool(const float* x, const uint8_t* valid, float* map, int B, int N, int d, void* stream) {
    STCAT_REQUIRE(x && valid && map, STCAT_EINVAL, "map2d_pool: null pointer");
    STCAT_REQUIRE(B >= 0 && N > 0 && d > 0 && B <= 65535, STCAT_ESHAPE, "map2d_pool: bad sizes");
    if (B == 0) return 0;
    dim3 grid(N, B, (d + 255) / 256);
    map2d_pool_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, valid, map, N, d);
    return check_launch("map2d_pool_kernel");
}

extern "C" int stcat_pos_sine(const uint8_t* mask, float* out, int n, int H, int W, int num_pos_feats, float temperature, float scale,
                              void* stream) {
    STCAT_REQUIRE(mask && out && n >= 0 && H > 0 && W > 0 && num_pos_feats > 0, STCAT_EINVAL, "pos_sine: bad arguments");
    if (n == 0) return 0;
    launch_pdl(pos_sine_kernel, dim3((unsigned)(n * H * W)), dim3(256), 0, (cudaStream_t)stream, mask, out, H, W, num_pos_feats, temperature, scale);
    return check_launch("pos_sine_kernel");
}
extern "C" int stcat_box_interp(const int64_t* frame_ids, const float* boxes, int m, float* out, int64_t first, int n_frames, void* stream) {
    STCAT_REQUIRE(frame_ids && boxes && out && m >= 0 && n_frames >= 0, STCAT_EINVAL, "box_interp: bad arguments");
    if (n_frames == 0) return 0;
    launch_pdl(box_interp_kernel, dim3((unsigned)((n_frames + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, frame_ids, boxes, m, out, first, n_frames);
    return check_launch("box_interp_kernel");
}
extern "C" int stcat_anchor_sine_fwd(const float* anchor, float* out, void* out_bf16, int64_t n, void* stream) {
    STCAT_REQUIRE(anchor && out && n >= 0, STCAT_EINVAL, "anchor_sine_fwd: bad arguments");
    if (n == 0) return 0;
    launch_pdl(anchor_sine_fwd_kernel, dim3(grid_for(n * 512)), dim3(256), 0, (cudaStream_t)stream, anchor, out, (__nv_bfloat16*)out_bf16, n);
    return check_launch("anchor_sine_fwd_kernel");
}
extern "C" int stcat_anchor_sine_bwd(const float* anchor, const float* dy, float* danchor, int64_t n, void* stream) {
    STCAT_REQUIRE(anchor && dy && danchor && n >= 0, STCAT_EINVAL, "anchor_sine_bwd: bad arguments");
    if (n == 0) return 0;
    launch_pdl(anchor_sine_bwd_kernel, dim3((unsigned)(n * 4)), dim3(128), 0, (cudaStream_t)stream, anchor, dy, danchor, n);
    return check_launch("anchor_sine_bwd_kernel");
}
extern "C" int stcat_box_refine_fwd(const float* delta, const float* anchor, float* out, int64_t n, float eps, void* stream) {
    STCAT_REQUIRE(delta && anchor && out && n >= 0, STCAT_EINVAL, "box_refine_fwd: bad arguments");
    if (n == 0) return 0;
    launch_pdl(box_refine_fwd_kernel, dim3(grid_for(n)), dim3(256), 0, (cudaStream_t)stream, delta, anchor, out, n, eps);
    return check_launch("box_refine_fwd_kernel");
}
extern "C" int stcat_box_refine_bwd(const float* out, const float* anchor, const float* g, float* ddelta, float* danchor, int64_t n,
                                    float eps, void* stream) {
    STCAT_REQUIRE(out && anchor && g && ddelta && n >= 0, STCAT_EINVAL, "box_refine_bwd: bad arguments");
    if (n == 0) return 0;
    launch_pdl(box_refine_bwd_kernel, dim3(grid_for(n)), dim3(256), 0, (cudaStream_t)stream, out, anchor, g, ddelta, danchor, n, eps);
    return check_launch("box_refine_bwd_kernel");
}

// keep bits of a [rows, cols] dropout site, one thread per 32-bit word (see DropArgs::bits)
__global__ void __launch_bounds__(256)
drop_bits_kernel(uint32_t* __restrict__ bits, int64_t rows, int cols, int wpr, const DropArgs drop_in) {
    const DropArgs drop = drop_resolve(drop_in);
    pdl_launch_dependents();
    pdl_wait();
    const int64_t total = rows * wpr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / wpr;
        const int w = (int)(i - r * wpr);
        uint32_t word = 0;
        const int c0 = w * 32;
        if (c0 < cols) {
            const uint64_t e0 = (uint64_t)r * (uint64_t)cols + c0;
            const int n = min(32, cols - c0);
            const uint64_t z0 = drop.offset + e0;
#pragma unroll
            for (int j = 0; j < 32; ++j)  // 32 independent hashes
                word |= ((j < n && drop_bits24(drop.seed, z0 + j) >= drop.thresh) ? 1u : 0u) << j;
        }
        bits[i] = word;
    }
}

extern "C" int stcat_dropout_bits(void* bits, int64_t rows, int cols, int wpr, float p, uint64_t seed, uint64_t offset, void* stream) {
    STCAT_REQUIRE(bits && rows >= 0 && cols > 0 && wpr * 32 >= cols && p > 0.f && p < 1.f, STCAT_EINVAL, "dropout_bits: bad arguments");
    if (rows == 0) return 0;
    const DropArgs d = make_drop(p, seed, offset);
    launch_pdl(drop_bits_kernel, dim3(grid_for(rows * wpr)), dim3(256), 0, (cudaStream_t)stream, (uint32_t*)bits, rows, cols, wpr, d);
    return check_launch("drop_bits_kernel");
}

extern "C" int stcat_dropout(const void* x, void* out, int dtype, int64_t n, float p, uint64_t seed, uint64_t offset, void* stream) {
    STCAT_REQUIRE(x && out && n >= 0 && p >= 0.f && p < 1.f, STCAT_EINVAL, "dropout: bad arguments (p=%f)", (double)p);
    if (n == 0) return 0;
    const DropArgs d = make_drop(p, seed, offset);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == STCAT_F32) launch_pdl(dropout_kernel<float>, dim3(grid_for(n)), dim3(256), 0, st, (const float*)x, (float*)out, n, d);
    else if (dtype == STCAT_BF16)
        launch_pdl(dropout_kernel<__nv_bfloat16>, dim3(grid_for(n)), dim3(256), 0, st, (const __nv_bfloat16*)x, (__nv_bfloat16*)out, n, d);
    else return set_err(STCAT_EINVAL, "dropout: bad dtype %d", dtype);
    return check_launch("dropout_kernel");
}
