// ppl::kernel::llm::cuda::pmx::sample_topk_topp (EXTERNAL, ppl.llm.kernel.cuda) with the argument list of its
// call site, src/backends/cuda/post_processor.cc:135,190-193.  Forwards to b2llm_sample_topk_topp.
#ifndef B2LLM_SHIM_PPL_KERNEL_LLM_CUDA_PMX_SAMPLE_H_
#define B2LLM_SHIM_PPL_KERNEL_LLM_CUDA_PMX_SAMPLE_H_

#include "ppl/common/retcode.h"

#include <cuda_runtime.h>
#include <stdint.h>

namespace ppl { namespace kernel { namespace llm { namespace cuda { namespace pmx {

int64_t sample_topk_topp_get_workspace_size(int32_t batch, int32_t vocab_size, int32_t top_k_val);

ppl::common::RetCode sample_topk_topp(cudaStream_t stream,
                                      const float* logits,                // (batch, batch_stride)
                                      const float* temperatures_optional, // (batch) or nullptr -> 1
                                      const float* top_p_optional,        // (batch) or nullptr -> top_p_val
                                      const float* rnd_optional,          // (batch) or nullptr -> rnd_val
                                      int32_t batch, int32_t vocab_size, int32_t batch_stride, int32_t top_k_val,
                                      float top_p_val, float rnd_val, void* workspace,
                                      int32_t* output,  // (batch)
                                      float* logprobs_optional);

}}}}} // namespace ppl::kernel::llm::cuda::pmx

#endif
