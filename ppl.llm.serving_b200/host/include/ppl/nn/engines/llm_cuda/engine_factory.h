// ppl::nn::llm::cuda::EngineFactory (EXTERNAL) -- resource_manager.cc:67,248-265.
#ifndef B2LLM_SHIM_PPL_NN_ENGINES_LLM_CUDA_ENGINE_FACTORY_H_
#define B2LLM_SHIM_PPL_NN_ENGINES_LLM_CUDA_ENGINE_FACTORY_H_

#include "ppl/nn/engines/engine.h"
#include "ppl/nn/engines/llm_cuda/options.h"

namespace ppl { namespace nn { namespace llm { namespace cuda {

class EngineFactory final {
public:
    static Engine* Create(const EngineOptions&);
    static DeviceContext* CreateDeviceContext(const DeviceOptions&);
    static DeviceContext* CreateHostDeviceContext(const HostDeviceOptions&);
};

}}}} // namespace ppl::nn::llm::cuda

#endif
