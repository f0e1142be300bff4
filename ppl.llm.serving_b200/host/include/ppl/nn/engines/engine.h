// ppl::nn::Engine (EXTERNAL): per-device backend object configured through varargs keys
// (resource_manager.cc:74-112,239; llm_engine.cc:114).
#ifndef B2LLM_SHIM_PPL_NN_ENGINES_ENGINE_H_
#define B2LLM_SHIM_PPL_NN_ENGINES_ENGINE_H_

#include "ppl/common/retcode.h"
#include "ppl/nn/common/device_context.h"

namespace ppl { namespace nn {

class Engine {
public:
    virtual ~Engine() {}
    virtual const char* GetName() const = 0;
    virtual ppl::common::RetCode Configure(uint32_t option, ...) = 0;
};

}} // namespace ppl::nn

#endif
