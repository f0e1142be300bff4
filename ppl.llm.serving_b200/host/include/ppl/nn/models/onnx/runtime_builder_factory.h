// ppl::nn::onnx::RuntimeBuilderFactory (EXTERNAL) -- resource_manager.cc:118.
#ifndef B2LLM_SHIM_PPL_NN_MODELS_ONNX_RUNTIME_BUILDER_FACTORY_H_
#define B2LLM_SHIM_PPL_NN_MODELS_ONNX_RUNTIME_BUILDER_FACTORY_H_

#include "ppl/nn/models/onnx/runtime_builder.h"

namespace ppl { namespace nn { namespace onnx {

class RuntimeBuilderFactory final {
public:
    static RuntimeBuilder* Create();
};

}}} // namespace ppl::nn::onnx

#endif
