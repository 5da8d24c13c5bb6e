// C-ABI glue: version / error reporting and the stcat_linear_* entry points (dispatch between the
// exact-fp32 SIMT kernel and the bf16 tcgen05 kernel).
#include "common.cuh"
#include <string.h>

namespace stcat {

static thread_local char g_err[512] = "";
char* err_buf() { return g_err; }

int set_err(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int gemm_simt(const void* A, int64_t sam, int64_t sak, const void* B, int64_t sbn, int64_t sbk, int in_dtype,
              void* C, int64_t ldc, int out_dtype, const float* bias, int M, int N, int K, int relu,
              int accumulate, cudaStream_t st);
int colsum(const void* dy, int64_t ld, int dtype, float* db, int M, int N, int accumulate, cudaStream_t st);
int relu_mask_2d(const void* y, int64_t ldy, int y_dtype, void* dx, int64_t lddx, int dx_dtype, int M, int N, cudaStream_t st);

// gemm_tcgen05.cu.  Returns 1 if the shape/alignment is not handled by the tensor-core kernel
// (the caller then raises: there is no silent fallback for bf16 operands unless allow_simt is set).
int gemm_tc_supported(int M, int N, int K, int64_t lda, int64_t ldb, int64_t ldc, const void* A, const void* B,
                      const void* C, int a_mn_major, int b_mn_major);
int gemm_tc(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major, void* C,
            int64_t ldc, int out_dtype, const float* bias, int M, int N, int K, int relu, int accumulate,
            cudaStream_t st, const GemmEpilogue* epi = nullptr);

// gemm_tcgen05.cu: grouped launch
struct TcTerm { const void* A; int64_t lda; const void* B; int64_t ldb; const float* bias; int K; };
struct TcJob {
    TcTerm term[3];
    int nterms;
    void* C; int64_t ldc; int M, N;
    int relu, accumulate;
    GemmEpilogue epi;
};
int gemm_tc_group(const TcJob* jobs, int njobs, int a_mn_major, int b_mn_major, int out_dtype, cudaStream_t st);
int colsum_group(const void* const* dy, const int64_t* ld, float* const* db, const int* M, const int* N, int njobs, int dtype,
                 cudaStream_t st);

}  // namespace stcat

using namespace stcat;

namespace stcat {
static const uint64_t* g_dropout_step = nullptr;
const uint64_t* dropout_step_ptr() { return g_dropout_step; }
void set_dropout_step_ptr(const uint64_t* p) { g_dropout_step = p; }
}  // namespace stcat

namespace stcat { void gemm_tc_set_sm_limit(int n); }
extern "C" int stcat_set_gemm_sm_limit(int n) {
    stcat::gemm_tc_set_sm_limit(n);
    return 0;
}
namespace stcat { int g_sm_cap = 0; }
extern "C" int stcat_set_sm_cap(int n) {
    stcat::g_sm_cap = n > 0 ? n : 0;
    return 0;
}
extern "C" int stcat_set_dropout_step(const void* counter) {
    stcat::set_dropout_step_ptr((const uint64_t*)counter);
    return 0;
}
extern "C" int stcat_abi_version(void) { return STCAT_ABI_VERSION; }
extern "C" const char* stcat_last_error(void) { return err_buf(); }

extern "C" int stcat_device_arch(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return set_err(-(int)e, "cudaGetDevice: %s", cudaGetErrorString(e));
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    return major * 10 + minor;
}

static bool dtype_ok(int d) { return d == STCAT_F32 || d == STCAT_BF16; }

extern "C" int stcat_linear_fwd(const void* x, int64_t ldx, int x_dtype, const void* w, int64_t ldw, int w_dtype,
                                const float* bias, void* y, int64_t ldy, int y_dtype, int M, int N, int K, int relu,
                                int accumulate, void* stream) {
    STCAT_REQUIRE(x && w && y, STCAT_EINVAL, "linear_fwd: null pointer");
    STCAT_REQUIRE(M >= 0 && N > 0 && K > 0, STCAT_EINVAL, "linear_fwd: bad sizes M=%d N=%d K=%d", M, N, K);
    STCAT_REQUIRE(x_dtype == w_dtype && dtype_ok(x_dtype) && dtype_ok(y_dtype), STCAT_EINVAL, "linear_fwd: dtypes x=%d w=%d y=%d", x_dtype, w_dtype, y_dtype);
    STCAT_REQUIRE(ldx >= K && ldw >= K && ldy >= N, STCAT_EINVAL, "linear_fwd: leading dimension too small");
    if (M == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (x_dtype == STCAT_BF16 && gemm_tc_supported(M, N, K, ldx, ldw, ldy, x, w, y, 0, 0))
        return gemm_tc(x, ldx, 0, w, ldw, 0, y, ldy, y_dtype, bias, M, N, K, relu, accumulate, st);
    return gemm_simt(x, ldx, 1, w, ldw, 1, x_dtype, y, ldy, y_dtype, bias, M, N, K, relu, accumulate, st);
}

extern "C" int stcat_dropout(const void* x, void* out, int dtype, int64_t n, float p, uint64_t seed, uint64_t offset, void* stream);

extern "C" int stcat_linear_dropout_fwd(const void* x, int64_t ldx, int x_dtype, const void* w, int64_t ldw, int w_dtype,
                                        const float* bias, void* y, int64_t ldy, int y_dtype, int M, int N, int K, int relu,
                                        float p, uint64_t seed, uint64_t offset, void* stream) {
    STCAT_REQUIRE(x && w && y, STCAT_EINVAL, "linear_dropout_fwd: null pointer");
    STCAT_REQUIRE(M >= 0 && N > 0 && K > 0, STCAT_EINVAL, "linear_dropout_fwd: bad sizes M=%d N=%d K=%d", M, N, K);
    STCAT_REQUIRE(x_dtype == w_dtype && dtype_ok(x_dtype) && dtype_ok(y_dtype), STCAT_EINVAL, "linear_dropout_fwd: dtypes");
    STCAT_REQUIRE(ldx >= K && ldw >= K && ldy == N, STCAT_EINVAL, "linear_dropout_fwd: y must be contiguous (ldy == N), ldx / ldw >= K");
    STCAT_REQUIRE(p >= 0.f && p < 1.f, STCAT_EINVAL, "linear_dropout_fwd: p=%f", (double)p);
    if (M == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (p > 0.f && x_dtype == STCAT_BF16 && gemm_tc_supported(M, N, K, ldx, ldw, ldy, x, w, y, 0, 0)) {
        GemmEpilogue epi;  // the mask is drawn in the GEMM epilogue: no second pass over y
        epi.drop = make_drop(p, seed, offset);
        return gemm_tc(x, ldx, 0, w, ldw, 0, y, ldy, y_dtype, bias, M, N, K, relu, 0, st, &epi);
    }
    int rc = stcat_linear_fwd(x, ldx, x_dtype, w, ldw, w_dtype, bias, y, ldy, y_dtype, M, N, K, relu, 0, stream);
    if (rc || p == 0.f) return rc;
    return stcat_dropout(y, y, y_dtype, (int64_t)M * N, p, seed, offset, stream);
}

static int linear_bwd_data_impl(const void* dy, int64_t lddy, int dy_dtype, const void* w, int64_t ldw,
                                int w_dtype, void* dx, int64_t lddx, int dx_dtype, const void* relu_y, int64_t ldy,
                                int y_dtype, float* dbias, int M, int N, int K, int accumulate, float alpha, void* stream) {
    STCAT_REQUIRE(dy && w && dx, STCAT_EINVAL, "linear_bwd_data: null pointer");
    STCAT_REQUIRE(M >= 0 && N > 0 && K > 0, STCAT_EINVAL, "linear_bwd_data: bad sizes M=%d N=%d K=%d", M, N, K);
    STCAT_REQUIRE(dy_dtype == w_dtype && dtype_ok(dy_dtype) && dtype_ok(dx_dtype), STCAT_EINVAL, "linear_bwd_data: dtypes");
    STCAT_REQUIRE(lddy >= N && ldw >= K && lddx >= K, STCAT_EINVAL, "linear_bwd_data: leading dimension too small");
    if (M == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    STCAT_REQUIRE(!relu_y || (dtype_ok(y_dtype) && ldy >= K), STCAT_EINVAL, "linear_bwd_data: relu_y dtype / leading dimension");
    STCAT_REQUIRE(!(relu_y && accumulate), STCAT_ESHAPE, "linear_bwd_data: relu_y with accumulate is not supported");
    // dx[m,k] = sum_n dy[m,n] w[n,k]:  A = dy (contraction index n contiguous), B(k, n) = w[n*ldw + k] (MN-major)
    if (dy_dtype == STCAT_BF16 && gemm_tc_supported(M, K, N, lddy, ldw, lddx, dy, w, dx, 0, 1)) {
        const bool fuse = (relu_y || dbias) && !accumulate && K % 64 == 0 &&
                          (!relu_y || (y_dtype == STCAT_BF16 && ((uintptr_t)relu_y & 15) == 0 && ldy % 8 == 0));
        if (fuse) {
            GemmEpilogue epi;
            epi.relu_mask = relu_y; epi.ld_mask = ldy; epi.colsum = dbias; epi.alpha = alpha;
            return gemm_tc(dy, lddy, 0, w, ldw, 1, dx, lddx, dx_dtype, nullptr, M, K, N, 0, accumulate, st, &epi);
        }
        STCAT_REQUIRE(alpha == 1.f, STCAT_ESHAPE, "linear_bwd_data_scaled: alpha needs the fused tensor-core epilogue (bf16, K %% 64 == 0, relu_y or dbias, no accumulate)");
        int rc = gemm_tc(dy, lddy, 0, w, ldw, 1, dx, lddx, dx_dtype, nullptr, M, K, N, 0, accumulate, st);
        if (rc) return rc;
    } else {
        STCAT_REQUIRE(alpha == 1.f, STCAT_ESHAPE, "linear_bwd_data_scaled: alpha needs the fused tensor-core epilogue");
        int rc = gemm_simt(dy, lddy, 1, w, 1, ldw, dy_dtype, dx, lddx, dx_dtype, nullptr, M, K, N, 0, accumulate, st);
        if (rc) return rc;
    }
    // unfused tail (exact-fp32 path, or shapes the fused epilogue does not take)
    if (relu_y) {
        int rc = relu_mask_2d(relu_y, ldy, y_dtype, dx, lddx, dx_dtype, M, K, st);
        if (rc) return rc;
    }
    if (dbias) return colsum(dx, lddx, dx_dtype, dbias, M, K, 1, st);
    return 0;
}

extern "C" int stcat_linear_bwd_data(const void* dy, int64_t lddy, int dy_dtype, const void* w, int64_t ldw,
                                     int w_dtype, void* dx, int64_t lddx, int dx_dtype, const void* relu_y, int64_t ldy,
                                     int y_dtype, float* dbias, int M, int N, int K, int accumulate, void* stream) {
    return linear_bwd_data_impl(dy, lddy, dy_dtype, w, ldw, w_dtype, dx, lddx, dx_dtype, relu_y, ldy, y_dtype, dbias, M, N, K,
                                accumulate, 1.f, stream);
}

extern "C" int stcat_linear_bwd_data_scaled(const void* dy, int64_t lddy, int dy_dtype, const void* w, int64_t ldw,
                                            int w_dtype, void* dx, int64_t lddx, int dx_dtype, const void* relu_y, int64_t ldy,
                                            int y_dtype, float* dbias, int M, int N, int K, float alpha, void* stream) {
    STCAT_REQUIRE(relu_y != nullptr, STCAT_EINVAL, "linear_bwd_data_scaled: relu_y required");
    return linear_bwd_data_impl(dy, lddy, dy_dtype, w, ldw, w_dtype, dx, lddx, dx_dtype, relu_y, ldy, y_dtype, dbias, M, N, K, 0,
                                alpha, stream);
}

extern "C" int stcat_linear_bwd_weight(const void* dy, int64_t lddy, int dy_dtype, const void* x, int64_t ldx,
                                       int x_dtype, float* dw, int64_t lddw, float* db, int M, int N, int K,
                                       int accumulate, void* stream) {
    STCAT_REQUIRE(dy && x && dw, STCAT_EINVAL, "linear_bwd_weight: null pointer");
    STCAT_REQUIRE(M >= 0 && N > 0 && K > 0, STCAT_EINVAL, "linear_bwd_weight: bad sizes M=%d N=%d K=%d", M, N, K);
    STCAT_REQUIRE(dy_dtype == x_dtype && dtype_ok(dy_dtype), STCAT_EINVAL, "linear_bwd_weight: dtypes");
    STCAT_REQUIRE(lddy >= N && ldx >= K && lddw >= K, STCAT_EINVAL, "linear_bwd_weight: leading dimension too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (M == 0) {
        if (!accumulate) {
            cudaMemset2DAsync(dw, lddw * sizeof(float), 0, K * sizeof(float), N, st);
            if (db) cudaMemsetAsync(db, 0, N * sizeof(float), st);
        }
        return check_launch("linear_bwd_weight memset");
    }
    int rc;
    // dw[n,k] = sum_m dy[m,n] x[m,k]:  A(n, m) = dy[m*lddy + n], B(k, m) = x[m*ldx + k]  (both MN-major)
    if (dy_dtype == STCAT_BF16 && gemm_tc_supported(N, K, M, lddy, ldx, lddw, dy, x, dw, 1, 1))
        rc = gemm_tc(dy, lddy, 1, x, ldx, 1, dw, lddw, STCAT_F32, nullptr, N, K, M, 0, accumulate, st);
    else
        rc = gemm_simt(dy, 1, lddy, x, 1, ldx, dy_dtype, dw, lddw, STCAT_F32, nullptr, N, K, M, 0, accumulate, st);
    if (rc) return rc;
    if (db) return colsum(dy, lddy, dy_dtype, db, M, N, accumulate, st);
    return 0;
}

extern "C" int stcat_linear_group(int kind, int in_dtype, const stcat_linear_job* jobs, int njobs, void* stream) {
    STCAT_REQUIRE(jobs && njobs >= 1 && njobs <= 12, STCAT_EINVAL, "linear_group: njobs=%d (1..12)", njobs);
    STCAT_REQUIRE(kind >= 0 && kind <= 2 && dtype_ok(in_dtype), STCAT_EINVAL, "linear_group: kind=%d in_dtype=%d", kind, in_dtype);
    cudaStream_t st = (cudaStream_t)stream;
    const int a_mn = kind == 2, b_mn = kind != 0;
    bool all_tc = in_dtype == STCAT_BF16;
    for (int j = 0; j < njobs; ++j) {
        const stcat_linear_job& J = jobs[j];
        STCAT_REQUIRE(J.out && J.nterms >= 1 && J.nterms <= 3 && J.rows > 0 && J.cols > 0 && dtype_ok(J.out_dtype), STCAT_EINVAL,
                      "linear_group: job %d malformed", j);
        STCAT_REQUIRE(kind != 2 || (J.nterms == 1 && J.out_dtype == STCAT_F32), STCAT_EINVAL, "linear_group: bwd_weight jobs have one term and fp32 output");
        STCAT_REQUIRE(J.out_dtype == jobs[0].out_dtype, STCAT_EINVAL, "linear_group: jobs of one call share out_dtype");
        for (int t = 0; t < J.nterms; ++t) {
            const stcat_linear_term& T = J.term[t];
            STCAT_REQUIRE(T.a && T.b && T.k > 0, STCAT_EINVAL, "linear_group: job %d term %d malformed", j, t);
            all_tc = all_tc && gemm_tc_supported(J.rows, J.cols, T.k, T.lda, T.ldb, J.ldo, T.a, T.b, J.out, a_mn, b_mn);
        }
        if (J.relu && J.accumulate) all_tc = false;
    }
    if (all_tc) {
        TcJob tj[12];
        for (int j = 0; j < njobs; ++j) {
            const stcat_linear_job& J = jobs[j];
            tj[j] = TcJob();
            for (int t = 0; t < J.nterms; ++t)
                tj[j].term[t] = TcTerm{J.term[t].a, J.term[t].lda, J.term[t].b, J.term[t].ldb, kind == 0 ? J.term[t].bias : nullptr, J.term[t].k};
            tj[j].nterms = J.nterms;
            tj[j].C = J.out; tj[j].ldc = J.ldo; tj[j].M = J.rows; tj[j].N = J.cols;
            tj[j].relu = J.relu; tj[j].accumulate = J.accumulate;
        }
        int rc = gemm_tc_group(tj, njobs, a_mn, b_mn, jobs[0].out_dtype, st);
        if (rc) return rc;
    } else {
        // job by job, term by term (exact-fp32 SIMT kernel, or shapes the tensor-core kernel does not take)
        for (int j = 0; j < njobs; ++j) {
            const stcat_linear_job& J = jobs[j];
            for (int t = 0; t < J.nterms; ++t) {
                const stcat_linear_term& T = J.term[t];
                const int acc = J.accumulate || t > 0;
                const int relu = (t == J.nterms - 1) ? J.relu : 0;
                STCAT_REQUIRE(!(relu && J.nterms > 1 && J.out_dtype != STCAT_F32), STCAT_ESHAPE,
                              "linear_group: multi-term ReLU job needs fp32 output on the unfused path");
                int rc;
                const float* bias = kind == 0 ? T.bias : nullptr;
                if (in_dtype == STCAT_BF16 && gemm_tc_supported(J.rows, J.cols, T.k, T.lda, T.ldb, J.ldo, T.a, T.b, J.out, a_mn, b_mn) &&
                    !(relu && acc))
                    rc = gemm_tc(T.a, T.lda, a_mn, T.b, T.ldb, b_mn, J.out, J.ldo, J.out_dtype, bias, J.rows, J.cols, T.k, relu, acc, st);
                else if (kind == 0)
                    rc = gemm_simt(T.a, T.lda, 1, T.b, T.ldb, 1, in_dtype, J.out, J.ldo, J.out_dtype, bias, J.rows, J.cols, T.k, relu, acc, st);
                else if (kind == 1)
                    rc = gemm_simt(T.a, T.lda, 1, T.b, 1, T.ldb, in_dtype, J.out, J.ldo, J.out_dtype, nullptr, J.rows, J.cols, T.k, 0, acc, st);
                else
                    rc = gemm_simt(T.a, 1, T.lda, T.b, 1, T.ldb, in_dtype, J.out, J.ldo, STCAT_F32, nullptr, J.rows, J.cols, T.k, 0, acc, st);
                if (rc) return rc;
            }
        }
    }
    if (kind == 2) {
        const void* dy[12]; int64_t ld[12]; float* db[12]; int Ms[12], Ns[12];
        int n = 0;
        for (int j = 0; j < njobs; ++j)
            if (jobs[j].dbias) {
                dy[n] = jobs[j].term[0].a; ld[n] = jobs[j].term[0].lda; db[n] = jobs[j].dbias;
                Ms[n] = jobs[j].term[0].k; Ns[n] = jobs[j].rows; ++n;
            }
        if (n) return colsum_group(dy, ld, db, Ms, Ns, n, in_dtype, st);
    }
    return 0;
}
