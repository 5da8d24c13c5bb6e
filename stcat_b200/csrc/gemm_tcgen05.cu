// bf16 tcgen05/TMA GEMM (placeholder until the tensor-core kernel lands: reports "unsupported" so
// the dispatcher uses the SIMT kernel for bf16 operands as well).
#include "common.cuh"

namespace stcat {

int gemm_tc_supported(int, int, int, int64_t, int64_t, int64_t, const void*, const void*, const void*, int, int) {
    return 0;
}

int gemm_tc(const void*, int64_t, int, const void*, int64_t, int, void*, int64_t, int, const float*, int, int, int,
            int, int, cudaStream_t) {
    return set_err(STCAT_ESHAPE, "gemm_tc: not built");
}

}  // namespace stcat
