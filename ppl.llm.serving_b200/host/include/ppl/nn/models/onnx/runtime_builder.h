// ppl::nn::onnx::RuntimeBuilder (EXTERNAL) -- resource_manager.cc:117-147.
// LoadModel() of the b2llm implementation reads a b2llm model-slice descriptor:
// the reference addresses the model only as `<model_dir>/model_slice_<rank>/model.onnx` and by tensor index,
// so the file's content is private to the runtime behind this interface.
#ifndef B2LLM_SHIM_PPL_NN_MODELS_ONNX_RUNTIME_BUILDER_H_
#define B2LLM_SHIM_PPL_NN_MODELS_ONNX_RUNTIME_BUILDER_H_

#include "ppl/nn/engines/engine.h"
#include "ppl/nn/runtime/runtime.h"

namespace ppl { namespace nn { namespace onnx {

class RuntimeBuilder {
public:
    struct Resources final {
        Engine** engines = nullptr;
        uint32_t engine_num = 0;
    };

    virtual ~RuntimeBuilder() {}
    virtual ppl::common::RetCode LoadModel(const char* model_file) = 0;
    virtual ppl::common::RetCode SetResources(const Resources&) = 0;
    virtual ppl::common::RetCode Preprocess() = 0;
    virtual Runtime* CreateRuntime() const = 0;
};

}}} // namespace ppl::nn::onnx

#endif
