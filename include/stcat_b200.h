/*
 * stcat_b200.h -- C ABI of libstcat_sm100.so: the B200 (sm_100a) kernels behind the STCAT hot path.
 *
 * The reference (jy0205/STCAT) has NO native / FFI interface: its hot path is pure-Python torch.nn
 * (SURVEY.md 2.B, 8b).  The seam that this library is dropped in behind is therefore the Python one
 * (models/grounding_model/__init__.py:5-9 build_encoder / build_decoder); each entry point below
 * replaces the torch.nn / ATen calls of the cited reference lines.  The host-side mirror of the
 * reference interface (same class/function names, arguments and errors) lives in stcat_b200/*.py
 * and binds these symbols with ctypes (see INTEGRATION.md for the binding a maintainer would add).
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless said otherwise;
 *  - matrices are row-major with an explicit leading dimension (elements);
 *  - dtype codes: STCAT_F32 = 0, STCAT_BF16 = 1;
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises,
 *    nothing allocates (workspaces are caller-provided), so every call is CUDA-graph capturable;
 *  - return 0 on success, a negative STCAT_E* for argument errors, a positive cudaError_t otherwise;
 *    stcat_last_error() returns a thread-local message.  There is no CPU fallback.
 */
#ifndef STCAT_B200_H_
#define STCAT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define STCAT_API __attribute__((visibility("default")))
#else
#define STCAT_API
#endif

#define STCAT_F32 0
#define STCAT_BF16 1

#define STCAT_EINVAL (-1)   /* bad argument (null pointer, non-positive size, bad dtype) */
#define STCAT_ESHAPE (-2)   /* shape not supported by this build (message says which) */
#define STCAT_EALIGN (-3)   /* pointer / leading dimension alignment requirement violated */

/* Bumped whenever a signature or the meaning of an argument changes; stcat_b200/cabi.py refuses a library whose version
 * differs from the one it was written against (a stale locally built .so would otherwise be called with new signatures). */
#define STCAT_ABI_VERSION 11
STCAT_API int stcat_abi_version(void);
STCAT_API const char* stcat_last_error(void);
/* compute capability major*10+minor of the current device, or <0; 100 expected */
STCAT_API int stcat_device_arch(void);

/* ------------------------------------------------------------------------------------------------
 * Linear layers (nn.Linear at modal_encoder.py:214-216,239; query_decoder.py:262-278,292-294,
 * 329-335,354-369,435,446-449,657; packed MHA in-projections torch functional.py:5866-5873; MLP
 * net_utils.py:14-25; heads pipeline.py:42-47).
 *   fwd       : y[M,N]   = act(x[M,K] . w[N,K]^T + bias[N])   (+ y if accumulate)
 *   bwd_data  : dx[M,K]  = dy[M,N] . w[N,K]                    (+ dx if accumulate)
 *               relu_y [M,K] (may be NULL): forward output of the ReLU layer that produced this Linear's input;
 *               dx is zeroed where relu_y <= 0 (F.relu backward fused into the epilogue, modal_encoder.py:239).
 *               dbias [K] fp32 (may be NULL): ACCUMULATED with the column sums of the stored dx, i.e. the bias
 *               gradient of that layer below (its dy is this dx).
 *   bwd_weight: dw[N,K] (+)= dy[M,N]^T . x[M,K];  db[N] (+)= column sums of dy   (db may be NULL)
 * fp32 operands run the exact-fp32 SIMT kernel; bf16 operands run the tcgen05/TMA kernel (fp32
 * accumulate).  y/dx/dw dtype is given separately.  relu: 0/1.  bias may be NULL.
 * ---------------------------------------------------------------------------------------------- */
STCAT_API int stcat_linear_fwd(const void* x, int64_t ldx, int x_dtype, const void* w, int64_t ldw, int w_dtype,
                     const float* bias, void* y, int64_t ldy, int y_dtype, int M, int N, int K, int relu,
                     int accumulate, void* stream);
STCAT_API int stcat_linear_bwd_data(const void* dy, int64_t lddy, int dy_dtype, const void* w, int64_t ldw, int w_dtype,
                          void* dx, int64_t lddx, int dx_dtype, const void* relu_y, int64_t ldy, int y_dtype,
                          float* dbias, int M, int N, int K, int accumulate, void* stream);
/* The FFN's inner dropout (`dropout(relu(linear1(x)))`, modal_encoder.py:239; query_decoder.py:435,657) without a pass of its own:
 *   stcat_linear_dropout_fwd     : y = drop(act(x W^T + b)); the mask (element index row * N + col, y contiguous: ldy == N; p, seed,
 *                                  offset as in stcat_dropout) is drawn in the GEMM epilogue.  Shapes the tensor-core kernel does not
 *                                  take run stcat_linear_fwd + stcat_dropout.
 *   stcat_linear_bwd_data_scaled : dx = alpha * (dy W) where relu_y > 0, else 0; dbias += colsum(dx).  With relu_y = the DROPPED
 *                                  forward activation (zero where the ReLU clamped or the mask dropped) and alpha = 1 / keep this is
 *                                  the backward of dropout and ReLU in one epilogue: no mask has to be regenerated.  bf16 operands,
 *                                  K % 64 == 0 (the fused tensor-core epilogue); STCAT_ESHAPE otherwise. */
STCAT_API int stcat_linear_dropout_fwd(const void* x, int64_t ldx, int x_dtype, const void* w, int64_t ldw, int w_dtype,
                        const float* bias, void* y, int64_t ldy, int y_dtype, int M, int N, int K, int relu, float p,
                        uint64_t seed, uint64_t offset, void* stream);
STCAT_API int stcat_linear_bwd_data_scaled(const void* dy, int64_t lddy, int dy_dtype, const void* w, int64_t ldw, int w_dtype,
                        void* dx, int64_t lddx, int dx_dtype, const void* relu_y, int64_t ldy, int y_dtype, float* dbias,
                        int M, int N, int K, float alpha, void* stream);
STCAT_API int stcat_linear_bwd_weight(const void* dy, int64_t lddy, int dy_dtype, const void* x, int64_t ldx, int x_dtype,
                            float* dw, int64_t lddw, float* db, int M, int N, int K, int accumulate, void* stream);

/* Grouped launch: up to 12 independent Linear GEMMs in ONE kernel launch, each the SUM of up to 3 terms.  The
 * reference's decoder layers are chains of [t, 256] Linears whose outputs it adds (q = q_content + q_time + q_pos,
 * query_decoder.py:329-339; k likewise; cross-attention q/k :355-366): launch latency, not FLOPs, bounds them, so the
 * independent ones share a launch and the added ones share an accumulator.  `kind` selects what a term means:
 *   0 fwd        out[rows=M, cols=N]  = act( sum_t a_t[M,k_t] . b_t[N,k_t]^T + bias_t[N] )      a = x,  b = w
 *   1 bwd_data   out[rows=M, cols=K]  =      sum_t a_t[M,k_t] . b_t[k_t,K]                      a = dy, b = w  (k_t = N_t)
 *   2 bwd_weight out[rows=N, cols=K] (+)=    a[k,N]^T . b[k,K];  dbias[N] += colsum(a)          a = dy, b = x  (k = M)
 * All jobs of a call share in_dtype; out_dtype / accumulate are per job.  bf16 operands run the tcgen05 kernel (one
 * launch when all jobs qualify), fp32 operands the exact SIMT kernel job by job. */
typedef struct stcat_linear_term {
    const void* a;
    int64_t lda;
    const void* b;
    int64_t ldb;
    const float* bias; /* kind 0 only; may be NULL */
    int32_t k;         /* contraction length of this term */
    int32_t reserved;
} stcat_linear_term;
typedef struct stcat_linear_job {
    stcat_linear_term term[3];
    int32_t nterms;
    int32_t rows, cols;
    int32_t relu, accumulate, out_dtype;
    void* out;
    int64_t ldo;
    float* dbias;      /* kind 2 only; ACCUMULATED; may be NULL */
} stcat_linear_job;
/* Upper bound on the CTAs (= SMs, the kernel is persistent) of the tcgen05 GEMM launches that follow; 0 = all SMs (default).
 * Host-side launch state, read when a GEMM is launched (a captured launch keeps its grid).  Used for GEMMs that run on a side
 * stream next to a latency-bound chain of small kernels (the decoder's memory-side projections). */
STCAT_API int stcat_set_gemm_sm_limit(int n);
/* SMs that EVERY persistent kernel of the library (tcgen05 GEMM and attention) sizes its grid for; 0 = all (default).  Host-side
 * launch state like the limit above.  For the window in which a collective with resident CTAs (the NCCL all-reduce of the
 * gradient ranges, stcat_b200/dp.py GradSync -- the reference's DistributedDataParallel buckets, scripts/train_net.py:31-36)
 * runs next to the backward pass: a persistent CTA whose SM is held by the collective cannot become resident, and its
 * statically assigned tiles would wait until the collective ends. */
STCAT_API int stcat_set_sm_cap(int n);
STCAT_API int stcat_linear_group(int kind, int in_dtype, const stcat_linear_job* jobs, int njobs, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Residual + LayerNorm (modal_encoder.py:237-238,240-241; query_decoder.py:344-345,431-432,436-437,
 * 612-613,653-654,658-659 and the final norms :222,527).  d must be 256 (cfg.MODEL.STCAT.HIDDEN).
 *   fwd: z = x + res (res may be NULL);  y = (z - mean) * rstd * gamma + beta;  mean/rstd [rows] saved.
 *        y_bf16 (may be NULL) additionally receives y rounded to bf16 (operand copy for the next GEMM).
 *   bwd: dz[rows,d] = LN backward; dgamma[d], dbeta[d] are ACCUMULATED (atomicAdd) into.
 *        dz_bf16 (may be NULL) additionally receives dz rounded to bf16 (operand copy for the dgrad / wgrad GEMMs
 *        that consume it); dbias[d] (may be NULL) is ACCUMULATED with the column sums of dz, i.e. the gradient of
 *        the bias of the Linear whose output is x (out_proj / linear2 in every block of the reference).
 * ---------------------------------------------------------------------------------------------- */
STCAT_API int stcat_layernorm_fwd(const float* x, const float* res, const float* gamma, const float* beta, float* y,
                        void* y_bf16, float* mean, float* rstd, int rows, int d, float eps, void* stream);
STCAT_API int stcat_layernorm_bwd(const float* dy, const float* x, const float* res, const float* gamma, const float* mean,
                        const float* rstd, float* dz, void* dz_bf16, float* dgamma, float* dbeta, float* dbias, int rows,
                        int d, void* stream);
/* The same pair with train-mode dropout on x folded in (the `dropout1/3/4(block output)` in front of every residual + norm of the
 * reference, modal_encoder.py:237-241; query_decoder.py:344,431-437,612,653-659): y = LayerNorm(drop(x) + res), element index of
 * the mask = row * d + column, (p, seed, offset) as in stcat_dropout below.  Backward: dz = gradient w.r.t. drop(x) + res (what
 * the residual branch receives, unmasked); dz_bf16 and dbias receive mask(dz), the gradient w.r.t. x (operand copy / bias
 * gradient of the Linear that produced x).  Replaces a separate dropout pass over x forward and over dz backward. */
STCAT_API int stcat_layernorm_dropout_fwd(const float* x, const float* res, const float* gamma, const float* beta, float* y,
                        void* y_bf16, float* mean, float* rstd, int rows, int d, float eps, float p, uint64_t seed,
                        uint64_t offset, void* stream);
STCAT_API int stcat_layernorm_dropout_bwd(const float* dy, const float* x, const float* res, const float* gamma, const float* mean,
                        const float* rstd, float* dz, void* dz_bf16, float* dgamma, float* dbeta, float* dbias, int rows,
                        int d, float p, uint64_t seed, uint64_t offset, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-head attention core, batch-major (torch functional.py:6630-6665 bmm/softmax/bmm; reference
 * attention.py:328-391).  For b in [0,B), h in [0,H):
 *   s[i,j] = scale * ( q1[b,i,h,:] . k1[b,j,h,:]  +  q2[b,i,h,:] . k2[b,j,h,:] )     (q2/k2 may be NULL)
 *   s[i,j] = -inf where key_mask[b,j] != 0;   p = softmax_j(s);   o[b,i,h,:] = sum_j p[i,j] v[b,j,h,:]
 * Row (b,i) of q1 starts at q1 + (b*Lq+i)*ldq, head h at column h*dh; same for the other operands
 * (rows of k1, k2 and v are (b*Lk+j)).  dh (per part) and the value head dim must both be 32.
 * The two-part form is the reference's per-head concat [content(32) ; positional(32)] of the box
 * decoder's cross attention (query_decoder.py:368-384) without materialising the concat.
 *   lse[B,H,Lq]     : log-sum-exp of each row (saved for backward)
 *   p_avg[B,Lq,Lk]  : optional (may be NULL); receives mean over heads of p (the `weights` output,
 *                     query_decoder.py:341,604; must be zero-filled by the caller)
 * bwd: dq1, dq2, dk1, dk2, dv get the gradients (fully overwritten); dp_avg (may be NULL) is the gradient
 *      flowing into p_avg.  delta[B,H,Lq] is caller-provided scratch.  o (may be NULL) is the forward output:
 *      with it the tensor-core kernel takes delta_i = dO_i . O_i instead of a second pass over the keys.
 * bf16 operands with Lq = Lk in [64, 256], a single score part and no p_avg (the spatial encoder's shape class)
 * run the tcgen05 kernels of attention_tc.cu; everything else runs the exact SIMT kernels.
 * ---------------------------------------------------------------------------------------------- */
STCAT_API int stcat_attention_fwd(const void* q1, const void* q2, int64_t ldq, const void* k1, const void* k2, int64_t ldk,
                        const void* v, int64_t ldv, void* o, int64_t ldo, int dtype, const uint8_t* key_mask,
                        float* lse, float* p_avg, int B, int H, int Lq, int Lk, int dh, float scale, void* stream);
STCAT_API int stcat_attention_bwd(const void* q1, const void* q2, int64_t ldq, const void* k1, const void* k2, int64_t ldk,
                        const void* v, int64_t ldv, const void* o, int64_t ldo, const void* d_o, int64_t lddo, int dtype,
                        const uint8_t* key_mask, const float* lse, const float* dp_avg, float* delta,
                        void* dq1, void* dq2, int64_t lddq, void* dk1, void* dk2, int64_t lddk, void* dv,
                        int64_t lddv, int B, int H, int Lq, int Lk, int dh, float scale, void* stream);

/* Train-mode dropout (nn.Dropout / F.dropout sites: modal_encoder.py:237-240; query_decoder.py:344,431-436,612,653-658;
 * attention.py:381; nn.MultiheadAttention(dropout=p)).  Counter-based and stateless: element idx of a site that drew
 * (seed, offset) is kept iff bits24(offset + idx + seed * 0x9E3779B97F4A7C15) >= p * 2^24 (bits24: the 64-bit counter folded to
 * 32 bits, lo ^ hi * 0x9E3779B1, then the multiply-xorshift mixer x ^= x>>16; x *= 0x21f0aaad; x ^= x>>15; x *= 0x735a2d97;
 * x ^= x>>15; top 24 bits -- stcat_b200/csrc/common.cuh drop_bits24), and
 * kept values are scaled by 1 / (keep probability).  The backward of a site is the same call on the gradient.
 *   stcat_dropout               : out[i] = keep(i) ? x[i] * scale : 0   (fp32 or bf16, in place allowed)
 *   stcat_attention_dropout_fwd : stcat_attention_fwd with dropout on the normalised probabilities, element index
 *                                 ((b*H + h)*Lq + i)*Lk + j; p_avg is the head average of the DROPPED probabilities
 *   stcat_attention_dropout_bwd : its backward (same seed / offset; `o` = the forward output, optional like stcat_attention_bwd's)
 * Kernel selection: the single-query, tcgen05 and mma.sync kernels apply the same mask for their shape classes, everything
 * else runs the generic SIMT kernels. */
STCAT_API int stcat_dropout(const void* x, void* out, int dtype, int64_t n, float p, uint64_t seed, uint64_t offset, void* stream);
/* Optional device-resident step counter (one uint64 in device memory, NULL = off, the default) that every dropout site --
 * stcat_dropout and the attention dropout entry points -- mixes into its seed when the kernel RUNS (not when it is
 * launched): seed' = seed + *counter * 0xD1B54A32D192ED03.  A training step captured into a CUDA graph increments the
 * counter once after its backward pass, so every replay draws fresh masks while forward and backward of one step still
 * agree.  Process-global; set it before capturing. */
STCAT_API int stcat_set_dropout_step(const void* counter);
STCAT_API int stcat_attention_dropout_fwd(const void* q1, const void* q2, int64_t ldq, const void* k1, const void* k2, int64_t ldk,
                                const void* v, int64_t ldv, void* o, int64_t ldo, int dtype, const uint8_t* key_mask,
                                float* lse, float* p_avg, int B, int H, int Lq, int Lk, int dh, float scale, float drop_p,
                                uint64_t seed, uint64_t offset, const void* keep_bits, int bits_wpr, void* stream);
STCAT_API int stcat_attention_dropout_bwd(const void* q1, const void* q2, int64_t ldq, const void* k1, const void* k2, int64_t ldk,
                                const void* v, int64_t ldv, const void* o, int64_t ldo, const void* d_o, int64_t lddo, int dtype,
                                const uint8_t* key_mask,
                                const float* lse, const float* dp_avg, float* delta, void* dq1, void* dq2, int64_t lddq,
                                void* dk1, void* dk2, int64_t lddk, void* dv, int64_t lddv, int B, int H, int Lq, int Lk, int dh,
                                float scale, float drop_p, uint64_t seed, uint64_t offset, const void* keep_bits, int bits_wpr,
                                void* stream);
/* keep_bits (optional, NULL = every kernel hashes per probability): the site's keep mask precomputed by stcat_dropout_bits with
 * rows = B*H*Lq, cols = Lk and bits_wpr >= Lk / 32 + 1 words per row (bit j of word w of row r = keep(r * Lk + 32 w + j), same
 * (p, seed, offset), same dropout step).  The tcgen05 kernels then read one word per 32 probabilities instead of hashing inside
 * their MUFU-bound loops; the generator is one HBM-light kernel that can run on a side stream under the projection GEMMs. */
STCAT_API int stcat_dropout_bits(void* bits, int64_t rows, int cols, int wpr, float p, uint64_t seed, uint64_t offset, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimizer-side step over a contiguous fp32 range of flat parameter / gradient / state buffers.
 *   stcat_sumsq      : *accum += sum_i x[i]^2   (device scalar; the caller zeroes it once per step and may add the
 *                      squared norms of parameters outside the flat buffer) -- the global gradient norm of
 *                      torch.nn.utils.clip_grad_norm_ (train_net.py:139-140) without a host sync
 *   stcat_adamw_step : per element, in this order (torch.optim.AdamW, engine/optimizer.py:44-46):
 *                        g *= min(1, max_norm / (sqrt(*total_sumsq) + 1e-6))        if max_norm > 0
 *                        p *= 1 - lr * weight_decay;  m += (g - m)(1 - beta1);  v = beta2 v + (1 - beta2) g^2
 *                        p -= lr / (1 - beta1^step) * m / (sqrt(v) / sqrt(1 - beta2^step) + eps)
 *                      then, optionally, ema = ema * ema_decay + (1 - ema_decay) * p (engine/optimizer.py:5-22) and
 *                      shadow_bf16 = bf16(p) (the GEMM-operand copy of the weights).  ema / shadow_bf16 / total_sumsq may
 *                      be NULL.  One pass over HBM (38 B per parameter). */
STCAT_API int stcat_sumsq(const float* x, int64_t n, float* accum, void* stream);
STCAT_API int stcat_adamw_step(float* p, const float* g, float* m, float* v, float* ema, void* shadow_bf16, int64_t n, float lr,
                               float beta1, float beta2, float eps, float weight_decay, int64_t step, const float* total_sumsq,
                               float max_norm, float ema_decay, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Either side of the hot path.
 *   pos_sine   : the backbone's image positional encoding, vision_model/position_encoding.py:70-94
 *                (PositionEmbeddingSine(num_pos_feats, normalize=True)): mask [n, H, W] uint8 (1 = padded) ->
 *                out [n, H, W, 2 * num_pos_feats] fp32, CHANNELS-LAST (the token assembly of the encoder reads it row by
 *                row; as an [n, 2F, H, W] tensor it is the permuted view).  Channels [0, F) = y, [F, 2F) = x.
 *   box_interp : engine/evaluate.py:20-38 linear_interp on the device: frame_ids [m] int64 ascending, boxes [m, 4] ->
 *                out [n_frames, 4] for the frames first .. first + n_frames - 1 (frames outside [frame_ids[0],
 *                frame_ids[m-1]] get -1).
 * ---------------------------------------------------------------------------------------------- */
STCAT_API int stcat_pos_sine(const uint8_t* mask, float* out, int n, int H, int W, int num_pos_feats, float temperature, float scale,
                             void* stream);
STCAT_API int stcat_box_interp(const int64_t* frame_ids, const float* boxes, int m, float* out, int64_t first, int n_frames,
                               void* stream);

/* Diagnostics (not on the product path): SM-clock timestamps at the phase boundaries of the tcgen05 spatial-attention
 * forward: CTA 0, its first 8 work items, 16 event slots per item (buf: 128 int64 in device memory; NULL = off). */
STCAT_API int stcat_debug_attn_trace(void* buf);
/* Same for the tcgen05 GEMM (stcat_linear_fwd with bf16 operands, 128 x 256 tiles): CTA 0, its first 8 tiles, 8 event
 * slots per tile (buf: 64 int64 in device memory; NULL = off). */
STCAT_API int stcat_debug_gemm_trace(void* buf);
/* Host-side launch counters of stcat_attention_fwd / _bwd (+ dropout variants) per kernel family since the library was
 * loaded: out5 = {single-query, tcgen05, mma.sync, shared-memory fp32, generic SIMT}.  Tests use it to assert that no
 * BASELINE shape is served by the generic kernels. */
STCAT_API int stcat_debug_attn_counts(long long* out5);

/* ------------------------------------------------------------------------------------------------
 * Element-wise helpers on [rows, cols] fp32 matrices (contiguous).
 *   add       : out = a + b  (q = k = src + pos, modal_encoder.py:225-226,235); out_bf16 optional copy
 *   relu_bwd  : dx = dy * (y > 0)                          (F.relu backward)
 *   cast_bf16 : fp32 -> bf16 operand copy (optionally transposed [cols, rows])
 * ---------------------------------------------------------------------------------------------- */
STCAT_API int stcat_add(const float* a, const float* b, float* out, void* out_bf16, int64_t n, void* stream);
STCAT_API int stcat_relu_bwd(const void* y, int y_dtype, void* dy_inout, int dy_dtype, int64_t n, void* stream);
STCAT_API int stcat_cast_bf16(const float* x, void* out_bf16, int64_t rows, int64_t cols, int transpose, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Layout glue either side of the encoder and between encoder and decoder, one launch each way (replaces the torch.cat /
 * transpose / expand / slice chains of modal_encoder.py:40-72 and query_decoder.py:83-120 and their autograd backward).
 * fp32 [n, S, d] is the frame-major token stream, S = 1 + HW + L (row 0 of a frame = its CLS slot); d % 4 == 0.
 *   token_assembly     : X[f] = [frame_cls ; vis[f]^T ; text[:, v(f)]], POS[f] = [local_pos ; vpos[f]^T ; 0] from
 *                        vis / vpos [n, d, HW] (NCHW feature maps), text [L, b, d], f2v [n] int64 (video of frame f; NULL
 *                        allowed iff b == 1); qk_op / x_op (bf16 [n, S, d], may be NULL) = bf16(X + POS), bf16(X): the GEMM
 *                        operands of the first spatial layer's q/k and v projections.
 *   token_assembly_bwd : from dX [n, S, d]: dvis [n, d, HW], dtext [L, b, d] (sum over the frames of each video, frames of a
 *                        video are consecutive: vid_start [b + 1] int64, NULL allowed iff b == 1), dcls [d] (sum over all
 *                        frames); every output is optional (NULL) and fully overwritten; sums run in frame order.
 *   mem_operands       : the decoder's operands of the encoder memory: mem_op = bf16(X[:, 1:]), pos_op = bf16(POS[:, 1:]),
 *                        mempos_op = bf16(X[:, 1:] + POS[:, 1:]), each [n (S - 1), d], and cls [n, d] = X[:, 0] (fp32).
 *                        pos_op / mempos_op / cls may be NULL; POS may be NULL when both of the former are.
 *   mem_operands_bwd   : dX [n, S, d] = [g_cls ; g_mem + g_mempos]; g_mem / g_mempos [n (S - 1), d] fp32 or bf16 (dtype codes),
 *                        g_cls [n, d] fp32; each may be NULL (= zero).  dX is fully overwritten.
 * ---------------------------------------------------------------------------------------------- */
STCAT_API int stcat_token_assembly(const float* vis, const float* vpos, const float* text, const int64_t* f2v, const float* frame_cls,
                                   const float* local_pos, float* X, float* POS, void* qk_op, void* x_op, int n, int d, int HW, int L,
                                   int b, void* stream);
STCAT_API int stcat_token_assembly_bwd(const float* dX, float* dvis, float* dtext, float* dcls, const int64_t* vid_start, int n, int d,
                                       int HW, int L, int b, void* stream);
STCAT_API int stcat_mem_operands(const float* X, const float* POS, void* mem_op, void* pos_op, void* mempos_op, float* cls, int n, int S,
                                 int d, void* stream);
STCAT_API int stcat_mem_operands_bwd(const void* g_mem, int g_mem_dtype, const void* g_mempos, int g_mempos_dtype, const float* g_cls,
                                     float* dX, int n, int S, int d, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Frame-CLS exchange between a spatial and a temporal encoder layer (modal_encoder.py:170-195), one un-padded video:
 *   cls_gather  : Y[1 + n, d] = [video ; X[:, r, :]] from the stream X [n, S, d] and the video token [1, d]; qk_op / y_op (bf16
 *                 [1 + n, d], may be NULL) = bf16(Y + pos), bf16(Y): the temporal layer's q/k and v operands (pos [1 + n, d]).
 *   cls_scatter : X[:, r, :] = Y[1:] in place; X_op (bf16 [n, S, d], may be NULL) receives the same rows; qk_next (bf16 [n, S, d],
 *                 may be NULL) receives bf16(Y[1:] + pos[:, r, :]) (pos fp32 [n, S, d]): the rows of the next spatial layer's q/k
 *                 operand, when that operand was projected early from the stream as it was before the temporal layer.
 * ---------------------------------------------------------------------------------------------- */
STCAT_API int stcat_cls_gather(const float* X, const float* video, const float* pos, float* Y, void* qk_op, void* y_op, int n, int S, int r,
                               int d, void* stream);
STCAT_API int stcat_cls_scatter(const float* Y, float* X, void* X_op, void* qk_next, const float* pos, int n, int S, int r, int d,
                                void* stream);

/* ------------------------------------------------------------------------------------------------
 * TemplateGenerator (query_decoder.py:441-475) as two small launches forward and three backward (the reference issues
 * 4 Linears + 2 tanh + mul + add + sigmoid, ~14 kernels forward and ~20 backward, on the decoder's dependent chain).
 * Weights W* are bf16 [d, d] (Wa: [q, d], q <= 8), biases fp32; GEMM operands (the video token, mod, the incoming
 * gradients) are rounded to bf16 exactly where the Linear entry points round theirs; accumulation is fp32.
 *   fwd: content [b, d] = Wc v + bc; gamma = tanh(Wg v + bg); beta = tanh(Wb v + bb)  (v = videos_cls [b, d]);
 *        mod[f] = gamma[v(f)] * frames_cls[f] + beta[v(f)] -> mod_op bf16 [n, d]; anchor [n, q] = sigmoid(Wa mod + ba);
 *        temp_query [n, d] (may be NULL) = content[v(f)], the temporal decoder's content query expanded to the frames.
 *   bwd: g_anchor [n, q] (gradient w.r.t. anchor), g_temp [n, d] or NULL (gradient w.r.t. content expanded to the
 *        frames).  Workspaces: dpq_op bf16 [n, q], dmod [n, d], dpre [3, b, d].  Outputs: d_frames_cls [n, d] and
 *        d_videos_cls [b, d] (overwritten); dWa [q, d], dba [q] and the optional dW{c,g,b} [d, d] / db{c,g,b} [d] are
 *        ACCUMULATED (+=).  f2v [n] / vid_start [b + 1] int64: NULL allowed iff b == 1.
 * ---------------------------------------------------------------------------------------------- */
STCAT_API int stcat_template_fwd(const float* videos_cls, const float* frames_cls, const int64_t* f2v, const void* Wc, const float* bc,
                                 const void* Wg, const float* bg, const void* Wb, const float* bb, const void* Wa, const float* ba,
                                 float* content, float* gamma, float* beta, void* mod_op, float* anchor, float* temp_query, int n, int b,
                                 int d, int q, void* stream);
STCAT_API int stcat_template_bwd(const float* g_anchor, const float* g_temp, const float* anchor, const float* videos_cls,
                                 const float* frames_cls, const int64_t* f2v, const int64_t* vid_start, const float* gamma,
                                 const float* beta, const void* mod_op, const void* Wc, const void* Wg, const void* Wb, const void* Wa,
                                 void* dpq_op, float* dmod, float* dpre, float* d_frames_cls, float* d_videos_cls, float* dWc, float* dbc,
                                 float* dWg, float* dbg, float* dWb, float* dbb, float* dWa, float* dba, int n, int b, int d, int q,
                                 void* stream);

/* ------------------------------------------------------------------------------------------------
 * Anchor-update chain between two box-decoder layers (query_decoder.py:205-219, 188-199):
 *   box_head_fwd : out[R,4] = sigmoid(W3 h + b3 + inverse_sigmoid(anchor, eps)) -- the last Linear of bbox_embed (h bf16 [R,K]
 *                  with leading dimension ldh, W3 bf16 [4,K]), the anchor refinement and, when sine != NULL, the sine
 *                  embedding of the refined anchor (anchor_sine of `out`; sine fp32 [R,512], sine_op optional bf16 copy) in
 *                  one launch.
 *   box_head_bwd : from g = d loss / d out: dd_op bf16 [R,4] = bf16(g out (1 - out)) (operand of the W3 weight gradient),
 *                  danchor [R,4] (optional, through the clamped logit) and dh bf16 [R,K] = (dd_op W3) * (h > 0): the data
 *                  gradient of the last Linear with the ReLU mask of its input h folded in.
 *   mul_cast     : out_f32[r,c] (may be NULL) = a[r,c] * b[r,c] and out_bf16 = its bf16 copy, c < cols; a has leading dimension
 *                  lda (query_sine = sine[:, :d] * query_scale(out) with its GEMM-operand copy in one launch); optionally
 *                  c_out_bf16[r,c] = bf16(c_in[r,c]) in the same launch (the operand copy of query_pos);
 *                  mul_cast_bwd: db = g * a (g fp32 or bf16).
 * ---------------------------------------------------------------------------------------------- */
STCAT_API int stcat_box_head_fwd(const void* h, int64_t ldh, const void* W, const float* bias, const float* anchor, float* out, float* sine,
                                 void* sine_op, int R, int K, float eps, void* stream);
STCAT_API int stcat_box_head_bwd(const float* g, const float* out, const float* anchor, const void* W, const void* h, int64_t ldh,
                                 void* dd_op, void* dh, float* danchor, int R, int K, float eps, void* stream);
STCAT_API int stcat_mul_cast(const float* a, int64_t lda, const float* b, float* out_f32, void* out_bf16, const float* c_in,
                             void* c_out_bf16, int64_t rows, int cols, void* stream);
STCAT_API int stcat_mul_cast_bwd(const void* g, int g_dtype, const float* a, int64_t lda, float* db, int64_t rows, int cols, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Anchor glue of the box decoder (query_decoder.py:188-219; net_utils.py:29-63), fp32, one launch each:
 *   anchor_sine : out[n,512] = gen_sineembed_for_position(anchor[n,4]) (order y,x,w,h; 128 dims per coordinate;
 *                 sin on even / cos on odd dims of 2 pi c / 10000^(2 floor(k/2)/128)); out_bf16 optional operand copy.
 *                 bwd: danchor[n,4] (fully written) from dy[n,512].
 *   box_refine  : out = sigmoid(delta + inverse_sigmoid(anchor, eps)) over n elements (anchor refinement :212-219 and
 *                 the box head pipeline.py:88-95); bwd: ddelta = g out (1 - out), danchor (may be NULL) likewise
 *                 through the clamped logit.
 * ---------------------------------------------------------------------------------------------- */
STCAT_API int stcat_anchor_sine_fwd(const float* anchor, float* out, void* out_bf16, int64_t n, void* stream);
STCAT_API int stcat_anchor_sine_bwd(const float* anchor, const float* dy, float* danchor, int64_t n, void* stream);
STCAT_API int stcat_box_refine_fwd(const float* delta, const float* anchor, float* out, int64_t n, float eps, void* stream);
STCAT_API int stcat_box_refine_bwd(const float* out, const float* anchor, const float* g, float* ddelta, float* danchor,
                         int64_t n, float eps, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Temporal start/end scoring (post_processor.py:30-53), one launch for b videos:
 *   score[v,i,j] = logsoftmax_t(sted[v,:,0])[i] + logsoftmax_t(sted[v,:,1])[j] + penalty(i,j)
 *   penalty = -1e32 where j <= i or i >= dur[v] or j >= dur[v]
 *   best[v] = flat index of the first maximum of score[v]   (start = best / t, end = best % t)
 * score (may be NULL) is [b,t,t]; durations is a device int32 array [b].
 * ---------------------------------------------------------------------------------------------- */
STCAT_API int stcat_sted_score(const float* sted, const int32_t* durations, float* score, int32_t* best, int b, int t,
                     void* stream);

/* ------------------------------------------------------------------------------------------------
 * VideoSTGLoss (criterion.py:26-207) for nl decoder layers in one launch: values AND gradients w.r.t. the predictions.
 *   coord [nl,n,4] predicted boxes (cx,cy,w,h); sted [nl,b,t,2]; act [nl,b,t] or NULL; attn [nl,b,t,t] or NULL.
 *   slice [K] int64 rows of coord with a ground-truth box, tgt_boxes [K,4]; time_mask [b,t] uint8 (1 = inside the
 *   clip); distrib [b,t,2] target start/end distributions; neg_f [b,t], nb_neg [b] (guided attention, :111-130);
 *   bce_weight, actioness [b,t] (:46-62).  coef5 = HOST array {bbox, giou, sted, guided_attn, actioness} weights.
 *   losses [nl,5]: unweighted values in that order.  d_* : gradient of sum_l sum_k coef_k * loss[l,k] w.r.t. the
 *   corresponding input (same shapes, fully overwritten; d_act / d_attn NULL iff act / attn NULL).
 * ---------------------------------------------------------------------------------------------- */
STCAT_API int stcat_stg_loss(const float* coord, const float* sted, const float* act, const float* attn, const int64_t* slice,
                   const float* tgt_boxes, const uint8_t* time_mask, const float* distrib, const float* neg_f,
                   const float* nb_neg, const float* bce_weight, const float* actioness, const float* coef5, float num_boxes,
                   int nl, int n, int b, int t, int K, float* losses, float* d_coord, float* d_sted, float* d_act,
                   float* d_attn, void* stream);

/* ------------------------------------------------------------------------------------------------
 * 2-D temporal proposal map (map2d_head.py:39-62, orphaned in the reference -- SURVEY.md 0-2):
 *   map[B,d,N,N]: cell (i,j) of the valid set = max_t x[B, i..j, d]; other cells 0.
 * x is [B, N, d] (already pooled to N positions); valid[N*N] uint8 is the reference's mask2d.
 * ---------------------------------------------------------------------------------------------- */
STCAT_API int stcat_map2d_pool(const float* x, const uint8_t* valid, float* map, int B, int N, int d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STCAT_B200_H_ */
