// ppl::nn::Runtime (EXTERNAL): the model addressed by tensor index; Run() is the whole forward
// (llm_engine.cc:113-116).  This is the plugin boundary b2llm implements (host/src/pplnn_b200.cc).
#ifndef B2LLM_SHIM_PPL_NN_RUNTIME_RUNTIME_H_
#define B2LLM_SHIM_PPL_NN_RUNTIME_RUNTIME_H_

#include "ppl/nn/runtime/tensor.h"

namespace ppl { namespace nn {

class Runtime {
public:
    virtual ~Runtime() {}
    virtual uint32_t GetInputCount() const = 0;
    virtual Tensor* GetInputTensor(uint32_t idx) const = 0;
    virtual uint32_t GetOutputCount() const = 0;
    virtual Tensor* GetOutputTensor(uint32_t idx) const = 0;
    virtual uint32_t GetDeviceContextCount() const = 0;
    virtual DeviceContext* GetDeviceContext(uint32_t idx) const = 0;
    virtual ppl::common::RetCode Run() = 0;
    virtual ppl::common::RetCode Configure(uint32_t option, ...) = 0;
};

}} // namespace ppl::nn

#endif
