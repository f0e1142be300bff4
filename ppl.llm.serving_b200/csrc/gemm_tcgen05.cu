// placeholder until the tcgen05 kernel lands (see gemm_mma.cu for the baseline path)
#include "common.cuh"
namespace b2llm {
bool gemm_tc_available() { return false; }
int32_t launch_gemm_tc(cudaStream_t, bool, const void*, const float*, const void*, const float*, int64_t, int, int, int,
                       void*, int64_t) {
    set_last_error("tcgen05 gemm not built");
    return B2LLM_ERR_UNSUPPORTED;
}
}  // namespace b2llm
