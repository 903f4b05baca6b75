// tcgen05 GEMM -- placeholder until the TMA/tcgen05 kernel lands.
#include "gemm_epilogue.cuh"
namespace vct {
int gemm_tcgen05(const vct_gemm_args* a, cudaStream_t st) {
    (void)a; (void)st;
    set_error("vct_gemm: VCT_GEMM_TCGEN05 not built in this library");
    return VCT_ERR_INVALID;
}
}  // namespace vct
