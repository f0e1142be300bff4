// A ppl.pmx LLaMA export (docs/llama_guide.md:12-36 of the reference: `Export.py --fused_qkv 1 --fused_kvcache 1
// --auto_causal 1 --quantized_cache 1 --dynamic_batching 1`) read as what b2llm needs: the model dimensions and
// graph-level constants (norm eps, rope theta, KV-cache attributes), and this rank's weight shards.
//
// The reference hands the file to ppl.nn (`builder->LoadModel(model_path)`, resource_manager.cc:124-131), which
// executes the graph node by node.  b2llm runs ONE fixed forward, so the graph is not
// interpreted: the initializers are bound by the parameter names the exporting torch module gives them
// (`layers.<i>.attention.wqkv.weight`, ...), and the pmx nodes are read only for their attributes -- and to REFUSE
// graphs this forward does not implement (bias terms, ALiBi, partial rotary, rope scaling) instead of computing
// something else.
#ifndef B2_PMX_LLAMA_H_
#define B2_PMX_LLAMA_H_

#include "b2llm.h"
#include "onnx_model.h"

#include <functional>
#include <memory>
#include <string>
#include <vector>

namespace b2onnx {

struct LayerTensors {
    const Tensor* attn_norm = nullptr;
    const Tensor* wqkv = nullptr; // --fused_qkv 1
    const Tensor* wq = nullptr;   // --fused_qkv 0
    const Tensor* wk = nullptr;
    const Tensor* wv = nullptr;
    const Tensor* wo = nullptr;
    const Tensor* ffn_norm = nullptr;
    const Tensor* w1 = nullptr; // gate
    const Tensor* w2 = nullptr; // down
    const Tensor* w3 = nullptr; // up
};

class PmxLlama {
public:
    // parses <path> (…/model_slice_<rank>/model.onnx); false + *err when it is not a readable pmx LLaMA export
    bool Open(const std::string& path, std::string* err);

    const b2llm_model_desc& desc() const { return desc_; }       // FULL-model dimensions
    int tensor_parallel_size() const { return tp_; }              // of the export (inferred from the shard shapes)
    int rank() const { return rank_; }                            // from the directory name, -1 if it has none
    bool fused_qkv() const { return fused_qkv_; }
    const std::vector<std::string>& warnings() const { return warnings_; }
    const Model& model() const { return *model_; }

    // Hands every weight of this rank to `sink` as fp16 in the layouts of b2llm_engine_load_weight_shard
    // (embedding / lm_head assembled to their full [vocab, hidden] from all slices when the export is tensor
    // parallel).  `name` is the initializer it came from.  Stops at the first non-zero return of `sink`.
    using Sink = std::function<int32_t(int32_t kind, int32_t layer, const void* fp16, uint64_t num_elements, const char* name)>;
    int32_t ForEachWeight(const Sink& sink, std::string* err);

private:
    bool Interpret(std::string* err);
    bool BindNames(const Model& m, std::vector<LayerTensors>* layers, const Tensor** emb, const Tensor** norm,
                   const Tensor** head, std::string* err) const;
    const Model* Sibling(int rank, std::string* err);

    std::unique_ptr<Model> model_;
    std::vector<std::unique_ptr<Model>> siblings_;
    b2llm_model_desc desc_{};
    int tp_ = 1, rank_ = -1;
    bool fused_qkv_ = false;
    int emb_split_ = 0;  // 0 whole, 1 hidden (column) split, 2 vocab (row) split
    int head_split_ = 0; // 0 whole, 2 vocab (row) split
    std::vector<LayerTensors> layers_;
    const Tensor* embedding_ = nullptr;
    const Tensor* final_norm_ = nullptr;
    const Tensor* lm_head_ = nullptr;
    std::vector<std::string> warnings_;
};

// fp32 / bf16 -> fp16 (round to nearest even), for exports that are not already fp16
uint16_t FloatToHalf(float f);

} // namespace b2onnx
#endif
