// RuntimeBuilder::LoadModel (resource_manager.cc:124-131) on the CPU: both kinds of model_slice_<r>/model.onnx -- a
// ppl.pmx ONNX export and a b2llm model-slice descriptor -- must be recognised before any device work
// (SetResources / Preprocess need the engine).  usage: test_load_model <model.onnx>...; prints one line per file.
#include "ppl/common/retcode.h"
#include "ppl/nn/models/onnx/runtime_builder_factory.h"

#include <stdio.h>

#include <memory>

int main(int argc, char** argv) {
    int failed = 0;
    for (int i = 1; i < argc; ++i) {
        std::unique_ptr<ppl::nn::onnx::RuntimeBuilder> b(ppl::nn::onnx::RuntimeBuilderFactory::Create());
        const auto rc = b->LoadModel(argv[i]);
        printf("%s: %s\n", argv[i], ppl::common::GetRetCodeStr(rc));
        failed += rc != ppl::common::RC_SUCCESS;
        // Preprocess without SetResources must fail cleanly, not crash
        if (b->Preprocess() == ppl::common::RC_SUCCESS) {
            printf("Preprocess succeeded without an engine\n");
            return 3;
        }
    }
    return failed ? 1 : 0;
}
