// ppl::common::InitNccl (EXTERNAL): one communicator per GPU of a single process (resource_manager.cc:393).
#ifndef B2LLM_SHIM_PPL_COMMON_NCCL_UTILS_H_
#define B2LLM_SHIM_PPL_COMMON_NCCL_UTILS_H_

#include "../retcode.h"

#ifdef PPLNN_CUDA_ENABLE_NCCL
#include "nccl.h"
#include <vector>

namespace ppl { namespace common {

RetCode InitNccl(uint32_t tensor_parallel_size, std::vector<ncclComm_t>* nccl_comm_list);

}} // namespace ppl::common
#endif

#endif
