// VideoSTGLoss (reference models/criterion.py:26-207) for all decoder layers in ONE launch, forward values and the
// gradients with respect to the predictions together: box L1 + GIoU on the annotated frames (:26-44), KL between the
// softmax over time of the start / end logits and the Gaussian target distributions (:64-109), guided attention on the
// head-averaged self-attention weights of the time decoder (:111-130) and the class-weighted actioness BCE (:46-62).
// The reference evaluates these with ~60 tiny tensor ops per layer plus their autograd backward (and .cpu()/.item()
// syncs); in a step that is otherwise a CUDA graph of fused kernels that was ~180 launches on the dependent chain.
// One CTA per decoder layer, 256 threads, block reductions; everything target-dependent is precomputed by the caller.
#include "common.cuh"
#include <math.h>

namespace stcat {

constexpr int LS_THREADS = 256;

struct StgLossArgs {
    const float* coord;      // [nl, n, 4]   predicted boxes (cx, cy, w, h) in (0, 1)
    const float* sted;       // [nl, b, t, 2]
    const float* act;        // [nl, b, t]  (may be null)
    const float* attn;       // [nl, b, t, t] (may be null)
    const int64_t* slice;    // [K] rows of coord with a ground-truth box
    const float* tgt_boxes;  // [K, 4] cxcywh
    const uint8_t* time_mask;  // [b, t] 1 = inside the clip
    const float* distrib;    // [b, t, 2] target start / end distributions
    const float* neg_f;      // [b, t] 1 = frame outside the annotated segment (guided attention rows)
    const float* nb_neg;     // [b]
    const float* bce_weight; // [b, t]
    const float* actioness;  // [b, t] {0, 1}
    float coef[5];           // bbox, giou, sted, guided_attn, actioness
    float num_boxes;
    int nl, n, b, t, K;
    float* losses;           // [nl, 5] unweighted loss values (same order as coef)
    float* d_coord;          // gradients of sum_l sum_k coef_k loss_{l,k}; same shapes as the inputs, fully written
    float* d_sted;
    float* d_act;
    float* d_attn;
};

__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < LS_THREADS / 32; ++i) t += red[i];
    return t;
}

__global__ void __launch_bounds__(LS_THREADS) stg_loss_kernel(const StgLossArgs a) {
    __shared__ float red[LS_THREADS / 32];
    extern __shared__ float sm[];  // [t] scratch for the softmax over time
    pdl_launch_dependents();
    pdl_wait();
    const int l = blockIdx.x, tid = threadIdx.x;
    const int b = a.b, t = a.t;
    const float eps = 1e-6f;

    // ---------------- boxes: L1 + GIoU on the annotated frames ----------------
    const float* coord = a.coord + (int64_t)l * a.n * 4;
    float* dcoord = a.d_coord + (int64_t)l * a.n * 4;
    for (int i = tid; i < a.n * 4; i += LS_THREADS) dcoord[i] = 0.f;
    __syncthreads();
    float l1 = 0.f, lg = 0.f;
    for (int k = tid; k < a.K; k += LS_THREADS) {
        const int64_t row = a.slice[k];
        const float* p = coord + row * 4;
        const float* g = a.tgt_boxes + (int64_t)k * 4;
        const float cx = p[0], cy = p[1], w = p[2], h = p[3];
        float dl[4];
        const float pv[4] = {cx, cy, w, h};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float d = pv[c] - g[c];
            l1 += fabsf(d);
            dl[c] = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * a.coef[0] / a.num_boxes;
        }
        // GIoU of (x1, y1, x2, y2) boxes (utils/box_utils.py:94-115), pair-aligned
        const float x1 = cx - 0.5f * w, y1 = cy - 0.5f * h, x2 = cx + 0.5f * w, y2 = cy + 0.5f * h;
        const float X1 = g[0] - 0.5f * g[2], Y1 = g[1] - 0.5f * g[3], X2 = g[0] + 0.5f * g[2], Y2 = g[1] + 0.5f * g[3];
        const float a1 = (x2 - x1) * (y2 - y1), a2 = (X2 - X1) * (Y2 - Y1);
        const float iw_raw = fminf(x2, X2) - fmaxf(x1, X1), ih_raw = fminf(y2, Y2) - fmaxf(y1, Y1);
        const float iw = fmaxf(iw_raw, 0.f), ih = fmaxf(ih_raw, 0.f);
        const float inter = iw * ih, uni = a1 + a2 - inter;
        const float cw_raw = fmaxf(x2, X2) - fminf(x1, X1), ch_raw = fmaxf(y2, Y2) - fminf(y1, Y1);
        const float cw = fmaxf(cw_raw, 0.f), ch = fmaxf(ch_raw, 0.f);
        const float area = cw * ch;
        const float giou = inter / uni - (area - uni) / area;
        lg += 1.f - giou;
        // d giou / d (x1, y1, x2, y2)
        const float g_inter = 1.f / uni + inter / (uni * uni) - 1.f / area;  // via inter directly and via union = a1 + a2 - inter
        const float g_a1 = -inter / (uni * uni) + 1.f / area;                // via union
        const float g_area = -uni / (area * area);
        float dx1 = 0.f, dy1 = 0.f, dx2 = 0.f, dy2 = 0.f;
        // inter = iw * ih
        if (iw_raw > 0.f) { const float s = g_inter * ih; if (x2 < X2) dx2 += s; else if (x2 == X2) dx2 += 0.5f * s; if (x1 > X1) dx1 -= s; else if (x1 == X1) dx1 -= 0.5f * s; }
        if (ih_raw > 0.f) { const float s = g_inter * iw; if (y2 < Y2) dy2 += s; else if (y2 == Y2) dy2 += 0.5f * s; if (y1 > Y1) dy1 -= s; else if (y1 == Y1) dy1 -= 0.5f * s; }
        // a1 = (x2 - x1)(y2 - y1)
        dx2 += g_a1 * (y2 - y1); dx1 -= g_a1 * (y2 - y1); dy2 += g_a1 * (x2 - x1); dy1 -= g_a1 * (x2 - x1);
        // area = cw * ch
        if (cw_raw > 0.f) { const float s = g_area * ch; if (x2 > X2) dx2 += s; else if (x2 == X2) dx2 += 0.5f * s; if (x1 < X1) dx1 -= s; else if (x1 == X1) dx1 -= 0.5f * s; }
        if (ch_raw > 0.f) { const float s = g_area * cw; if (y2 > Y2) dy2 += s; else if (y2 == Y2) dy2 += 0.5f * s; if (y1 < Y1) dy1 -= s; else if (y1 == Y1) dy1 -= 0.5f * s; }
        const float sg = -a.coef[1] / a.num_boxes;  // loss = (1 - giou) / num_boxes
        float* d = dcoord + row * 4;  // rows of the slice are distinct
        d[0] = dl[0] + sg * (dx1 + dx2);
        d[1] = dl[1] + sg * (dy1 + dy2);
        d[2] = dl[2] + sg * 0.5f * (dx2 - dx1);
        d[3] = dl[3] + sg * 0.5f * (dy2 - dy1);
    }
    l1 = block_sum(l1, red) / a.num_boxes;
    lg = block_sum(lg, red) / a.num_boxes;

    // ---------------- start / end: KL(softmax_t(logits) || target) ----------------
    const float* sted = a.sted + (int64_t)l * b * t * 2;
    float* dsted = a.d_sted + (int64_t)l * b * t * 2;
    float lsted = 0.f;
    for (int v = 0; v < b; ++v) {
        for (int c = 0; c < 2; ++c) {
            // masked softmax over the t positions of video v, channel c
            float mx = -INFINITY;
            for (int i = tid; i < t; i += LS_THREADS)
                if (a.time_mask[v * t + i]) mx = fmaxf(mx, sted[((int64_t)v * t + i) * 2 + c]);
            mx = warp_max(mx);
            __syncthreads();
            if ((tid & 31) == 0) red[tid >> 5] = mx;
            __syncthreads();
            mx = red[0];
#pragma unroll
            for (int i = 1; i < LS_THREADS / 32; ++i) mx = fmaxf(mx, red[i]);
            float se = 0.f;
            for (int i = tid; i < t; i += LS_THREADS) {
                const float e = a.time_mask[v * t + i] ? expf(sted[((int64_t)v * t + i) * 2 + c] - mx) : 0.f;
                sm[i] = e;
                se += e;
            }
            se = block_sum(se, red);
            const float inv = se > 0.f ? 1.f / se : 0.f;
            // f_i = p_i log((p_i + eps) / d_i) on unmasked positions;  dL/ds_j = p_j (g_j - sum_i p_i g_i)
            float kl = 0.f, pg = 0.f;
            for (int i = tid; i < t; i += LS_THREADS) {
                const float p = sm[i] * inv;
                float gi = 0.f;
                if (a.time_mask[v * t + i]) {
                    const float lr = logf((p + eps) / a.distrib[((int64_t)v * t + i) * 2 + c]);
                    kl += p * lr;
                    gi = lr + p / (p + eps);
                }
                pg += p * gi;
                sm[i] = p;
                // stash g_i in the gradient buffer for the second pass
                dsted[((int64_t)v * t + i) * 2 + c] = gi;
            }
            kl = block_sum(kl, red);
            pg = block_sum(pg, red);
            lsted += kl;
            const float sc = a.coef[2] / (float)(b * t);
            for (int i = tid; i < t; i += LS_THREADS) {
                const int64_t o = ((int64_t)v * t + i) * 2 + c;
                dsted[o] = sm[i] * (dsted[o] - pg) * sc;
            }
            __syncthreads();
        }
    }
    lsted /= (float)(b * t);

    // ---------------- guided attention on the time decoder's self-attention weights ----------------
    float lattn = 0.f;
    if (a.attn) {
        const float* w = a.attn + (int64_t)l * b * t * t;
        float* dw = a.d_attn + (int64_t)l * b * t * t;
        for (int64_t i = tid; i < (int64_t)b * t * t; i += LS_THREADS) {
            const int v = (int)(i / ((int64_t)t * t)), r = (int)((i / t) % t);
            const float nf = a.neg_f[v * t + r];
            const float x = 1.f - w[i] + eps;
            const float inv_nb = 1.f / a.nb_neg[v];
            lattn += nf != 0.f ? -logf(x) * nf * inv_nb : 0.f;
            dw[i] = nf != 0.f ? a.coef[3] * nf * inv_nb / (x * (float)b) : 0.f;
        }
        lattn = block_sum(lattn, red) / (float)b;
    }

    // ---------------- actioness: weighted BCE with logits inside the clip ----------------
    float lact = 0.f;
    if (a.act) {
        const float* x = a.act + (int64_t)l * b * t;
        float* dx = a.d_act + (int64_t)l * b * t;
        for (int i = tid; i < b * t; i += LS_THREADS) {
            const float xi = x[i], y = a.actioness[i], wgt = a.bce_weight[i];
            const float m = a.time_mask[i] ? 1.f : 0.f;
            const float li = fmaxf(xi, 0.f) - xi * y + log1pf(expf(-fabsf(xi)));
            lact += wgt * li * m;
            const float sg = 1.f / (1.f + expf(-xi));
            dx[i] = a.coef[4] * wgt * (sg - y) * m / (float)(b * t);
        }
        lact = block_sum(lact, red) / (float)(b * t);
    }
    if (tid == 0) {
        float* o = a.losses + l * 5;
        o[0] = l1; o[1] = lg; o[2] = lsted; o[3] = lattn; o[4] = lact;
    }
}

}  // namespace stcat

using namespace stcat;

extern "C" int stcat_stg_loss(const float* coord, const float* sted, const float* act, const float* attn, const int64_t* slice,
                              const float* tgt_boxes, const uint8_t* time_mask, const float* distrib, const float* neg_f,
                              const float* nb_neg, const float* bce_weight, const float* actioness, const float* coef5,
                              float num_boxes, int nl, int n, int b, int t, int K, float* losses, float* d_coord, float* d_sted,
                              float* d_act, float* d_attn, void* stream) {
    STCAT_REQUIRE(coord && sted && slice && tgt_boxes && time_mask && distrib && coef5 && losses && d_coord && d_sted, STCAT_EINVAL,
                  "stg_loss: null pointer");
    STCAT_REQUIRE((act == nullptr) == (d_act == nullptr) && (attn == nullptr) == (d_attn == nullptr), STCAT_EINVAL,
                  "stg_loss: act / attn and their gradient buffers go together");
    STCAT_REQUIRE(!act || (bce_weight && actioness), STCAT_EINVAL, "stg_loss: actioness targets missing");
    STCAT_REQUIRE(!attn || (neg_f && nb_neg), STCAT_EINVAL, "stg_loss: guided-attention masks missing");
    STCAT_REQUIRE(nl > 0 && n > 0 && b > 0 && t > 0 && K >= 0 && num_boxes > 0.f, STCAT_EINVAL, "stg_loss: bad sizes");
    StgLossArgs a;
    a.coord = coord; a.sted = sted; a.act = act; a.attn = attn; a.slice = slice; a.tgt_boxes = tgt_boxes; a.time_mask = time_mask;
    a.distrib = distrib; a.neg_f = neg_f; a.nb_neg = nb_neg; a.bce_weight = bce_weight; a.actioness = actioness;
    for (int i = 0; i < 5; ++i) a.coef[i] = coef5[i];  // host array
    a.num_boxes = num_boxes; a.nl = nl; a.n = n; a.b = b; a.t = t; a.K = K;
    a.losses = losses; a.d_coord = d_coord; a.d_sted = d_sted; a.d_act = d_act; a.d_attn = d_attn;
    cudaError_t le = launch_pdl(stg_loss_kernel, dim3(nl), dim3(LS_THREADS), (size_t)t * sizeof(float), (cudaStream_t)stream, a);
    if (le != cudaSuccess) return set_err((int)le, "stg_loss launch: %s", cudaGetErrorString(le));
    return check_launch("stg_loss_kernel");
}
