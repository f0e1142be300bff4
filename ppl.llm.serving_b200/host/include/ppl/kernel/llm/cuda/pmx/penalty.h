// ppl::kernel::llm::cuda::pmx::apply_penalty (EXTERNAL, ppl.llm.kernel.cuda) with the argument list of its call
// site, src/backends/cuda/post_processor.cc:271-274.  Forwards to b2llm_apply_penalty.
#ifndef B2LLM_SHIM_PPL_KERNEL_LLM_CUDA_PMX_PENALTY_H_
#define B2LLM_SHIM_PPL_KERNEL_LLM_CUDA_PMX_PENALTY_H_

#include "ppl/common/retcode.h"

#include <cuda_runtime.h>
#include <stdint.h>

namespace ppl { namespace kernel { namespace llm { namespace cuda { namespace pmx {

ppl::common::RetCode apply_penalty(cudaStream_t stream, const float* logits_in, const float* temperatures,
                                   const float* repetition_penalties, const float* presence_penalties_optional,
                                   const float* frequency_penalties_optional, const int64_t* batch_slots,
                                   const int64_t* token_inputs, const int64_t* seqstarts, const int64_t* start_pos,
                                   int32_t batch, int32_t vocab_size, uint16_t* penalty_count_map, float* logits_out);

}}}}} // namespace ppl::kernel::llm::cuda::pmx

#endif
