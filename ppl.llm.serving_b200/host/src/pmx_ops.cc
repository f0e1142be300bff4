// ppl::kernel::llm::cuda::pmx::{sample_topk_topp, apply_penalty}: the two free functions the reference's
// CudaPostProcessor calls (src/backends/cuda/post_processor.cc:135,190-193,271-274), forwarded to the C ABI.
#include "b2llm.h"

#include "ppl/common/log.h"
#include "ppl/kernel/llm/cuda/pmx/penalty.h"
#include "ppl/kernel/llm/cuda/pmx/sample.h"

namespace ppl { namespace kernel { namespace llm { namespace cuda { namespace pmx {

int64_t sample_topk_topp_get_workspace_size(int32_t batch, int32_t vocab_size, int32_t top_k_val) {
    return b2llm_sample_topk_topp_get_workspace_size(batch, vocab_size, top_k_val);
}

ppl::common::RetCode sample_topk_topp(cudaStream_t stream, const float* logits, const float* temperatures_optional,
                                      const float* top_p_optional, const float* rnd_optional, int32_t batch,
                                      int32_t vocab_size, int32_t batch_stride, int32_t top_k_val, float top_p_val,
                                      float rnd_val, void* workspace, int32_t* output, float* logprobs_optional) {
    const int32_t rc = b2llm_sample_topk_topp((void*)stream, logits, temperatures_optional, top_p_optional, rnd_optional,
                                              batch, vocab_size, batch_stride, top_k_val, top_p_val, rnd_val, workspace,
                                              output, logprobs_optional);
    if (rc != B2LLM_OK) {
        LOG(ERROR) << "sample_topk_topp: " << b2llm_last_error();
    }
    return (ppl::common::RetCode)rc;
}

ppl::common::RetCode apply_penalty(cudaStream_t stream, const float* logits_in, const float* temperatures,
                                   const float* repetition_penalties, const float* presence_penalties_optional,
                                   const float* frequency_penalties_optional, const int64_t* batch_slots,
                                   const int64_t* token_inputs, const int64_t* seqstarts, const int64_t* start_pos,
                                   int32_t batch, int32_t vocab_size, uint16_t* penalty_count_map, float* logits_out) {
    const int32_t rc = b2llm_apply_penalty((void*)stream, logits_in, temperatures, repetition_penalties,
                                           presence_penalties_optional, frequency_penalties_optional, batch_slots,
                                           token_inputs, seqstarts, start_pos, batch, vocab_size, penalty_count_map,
                                           logits_out);
    if (rc != B2LLM_OK) {
        LOG(ERROR) << "apply_penalty: " << b2llm_last_error();
    }
    return (ppl::common::RetCode)rc;
}

}}}}} // namespace ppl::kernel::llm::cuda::pmx
