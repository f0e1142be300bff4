// ppl::common::InitCudaEnv (EXTERNAL): bind the calling worker thread to a device (resource_manager.cc:218).
#ifndef B2LLM_SHIM_PPL_COMMON_CUDA_ENV_H_
#define B2LLM_SHIM_PPL_COMMON_CUDA_ENV_H_

#include "../retcode.h"

namespace ppl { namespace common {

RetCode InitCudaEnv(int device_id);

}} // namespace ppl::common

#endif
