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

}  // namespace stcat

using namespace stcat;

extern "C" int stcat_abi_version(void) { return 2; }
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

extern "C" int stcat_linear_bwd_data(const void* dy, int64_t lddy, int dy_dtype, const void* w, int64_t ldw,
                                     int w_dtype, void* dx, int64_t lddx, int dx_dtype, const void* relu_y, int64_t ldy,
                                     int y_dtype, float* dbias, int M, int N, int K, int accumulate, void* stream) {
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
            epi.relu_mask = relu_y; epi.ld_mask = ldy; epi.colsum = dbias;
            return gemm_tc(dy, lddy, 0, w, ldw, 1, dx, lddx, dx_dtype, nullptr, M, K, N, 0, accumulate, st, &epi);
        }
        int rc = gemm_tc(dy, lddy, 0, w, ldw, 1, dx, lddx, dx_dtype, nullptr, M, K, N, 0, accumulate, st);
        if (rc) return rc;
    } else {
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
