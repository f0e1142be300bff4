// See pmx_llama.h.  Conventions of the export this file is written against (ppl.pmx model_zoo/llama, the exporter
// docs/llama_guide.md:12-36 points at; NOT verifiable offline -- no export and no ppl.pmx checkout exist in this
// image, so every assumption is checked against the tensor shapes and refused loudly when it does not hold):
//   * initializers carry the torch parameter names: tok_embeddings.weight, norm.weight, output.weight,
//     layers.<i>.{attention_norm,ffn_norm}.weight, layers.<i>.attention.{wqkv | wq,wk,wv}.weight,
//     layers.<i>.attention.wo.weight, layers.<i>.feed_forward.{w1 (gate), w2 (down), w3 (up)}.weight
//     (Hugging Face style names are accepted as aliases);
//   * each model_slice_<r> holds rank r's shard: column-parallel linears split by output rows (wqkv = local q heads,
//     then local k heads, then local v heads), row-parallel linears by input columns, norms whole, the embedding
//     split along hidden (ParallelEmbedding) or along vocab or whole, the lm head split along vocab or whole;
//   * pmx nodes (domain "pmx" / "pmx.dynamic_batching", later "opmx") carry the constants as attributes: RMSNorm.eps,
//     RotaryPositionEmbedding.{theta,rotary_dim,bypass_key,max_position_embeddings,scaling_type},
//     MultiHeadCacheAttention / KeyValueCache.{num_heads,num_kv_heads,head_dim,is_causal,is_alibi,quant_bit,quant_group,
//     cache_mode,cache_layout,page_size}, *ParallelLinear.{in_features,out_features,bias_term},
//     ParallelEmbedding.{num_embeddings,embedding_dims}.
// Where an attribute is absent the value comes from <model-dir>/params.json (src/common/config.cc:41-145), then from
// LLaMA-2 defaults, with a warning.
#include "pmx_llama.h"

#include <ctype.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <fstream>
#include <sstream>

namespace b2onnx {

uint16_t FloatToHalf(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7fffffffu;
    if (x >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | (x > 0x7f800000u ? 0x200u : 0)); // inf / nan
    if (x >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);                                    // rounds to inf
    if (x < 0x33000001u) return (uint16_t)sign;                                                 // rounds to zero
    int32_t exp = (int32_t)(x >> 23) - 127 + 15;
    uint32_t man = x & 0x7fffffu;
    if (exp <= 0) { // subnormal half
        man |= 0x800000u;
        const int shift = 14 - exp; // 13 + (1 - exp)
        const uint32_t q = man >> shift, rem = man & ((1u << shift) - 1), half = 1u << (shift - 1);
        return (uint16_t)(sign | (q + ((rem > half || (rem == half && (q & 1))) ? 1 : 0)));
    }
    const uint32_t q = ((uint32_t)exp << 10) | (man >> 13), rem = man & 0x1fffu;
    return (uint16_t)(sign | (q + ((rem > 0x1000u || (rem == 0x1000u && (q & 1))) ? 1 : 0)));
}

namespace {

std::string Lower(std::string s) {
    for (auto& c : s) c = (char)tolower((unsigned char)c);
    return s;
}

bool EndsWith(const std::string& s, const char* suffix) {
    const size_t n = strlen(suffix);
    return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}

// "key": <integer | true | false> in a flat JSON object (params.json has no nesting, config.cc:41-145)
bool JsonInt(const std::string& text, const char* key, int64_t* out) {
    const std::string pat = std::string("\"") + key + "\"";
    size_t p = text.find(pat);
    if (p == std::string::npos) return false;
    p = text.find(':', p + pat.size());
    if (p == std::string::npos) return false;
    ++p;
    while (p < text.size() && isspace((unsigned char)text[p])) ++p;
    if (text.compare(p, 4, "true") == 0) {
        *out = 1;
        return true;
    }
    if (text.compare(p, 5, "false") == 0) {
        *out = 0;
        return true;
    }
    char* end = nullptr;
    const long long v = strtoll(text.c_str() + p, &end, 10);
    if (end == text.c_str() + p) return false;
    *out = v;
    return true;
}

const Node* Consumer(const Model& m, const Tensor* t) {
    if (!t) return nullptr;
    for (const auto& n : m.nodes) {
        for (const auto& in : n.inputs) {
            if (in == t->name) return &n;
        }
    }
    return nullptr;
}

bool Is2D(const Tensor* t) {
    return t && t->dims.size() == 2 && t->dims[0] > 0 && t->dims[1] > 0;
}

std::string Shape(const Tensor* t) {
    std::ostringstream os;
    os << "[";
    for (size_t i = 0; t && i < t->dims.size(); ++i) os << (i ? ", " : "") << t->dims[i];
    os << "]";
    return os.str();
}

} // namespace

bool PmxLlama::BindNames(const Model& m, std::vector<LayerTensors>* layers, const Tensor** emb, const Tensor** norm,
                         const Tensor** head, std::string* err) const {
    *emb = *norm = *head = nullptr;
    layers->clear();
    for (const Tensor& t : m.initializers) {
        std::string n = t.name;
        if (n.rfind("model.", 0) == 0) n = n.substr(6);
        if (n.rfind("transformer.", 0) == 0) n = n.substr(12);
        if (EndsWith(n, ".bias") && (n.find("attention") != std::string::npos || n.find("feed_forward") != std::string::npos ||
                                     n.find("self_attn") != std::string::npos || n.find("mlp") != std::string::npos)) {
            *err = "initializer [" + t.name + "]: linear layers with a bias term are not supported (LLaMA has none)";
            return false;
        }
        if (n == "tok_embeddings.weight" || n == "embed_tokens.weight") {
            *emb = &t;
        } else if (n == "norm.weight") {
            *norm = &t;
        } else if (n == "output.weight" || n == "lm_head.weight") {
            *head = &t;
        } else if (n.rfind("layers.", 0) == 0) {
            char* end = nullptr;
            const long idx = strtol(n.c_str() + 7, &end, 10);
            if (end == n.c_str() + 7 || *end != '.' || idx < 0 || idx > 4096) continue;
            const std::string rest(end + 1);
            if ((size_t)idx >= layers->size()) layers->resize(idx + 1);
            LayerTensors& L = (*layers)[idx];
            if (rest == "attention_norm.weight" || rest == "input_layernorm.weight") L.attn_norm = &t;
            else if (rest == "ffn_norm.weight" || rest == "post_attention_layernorm.weight") L.ffn_norm = &t;
            else if (rest == "attention.wqkv.weight" || rest == "self_attn.qkv_proj.weight") L.wqkv = &t;
            else if (rest == "attention.wq.weight" || rest == "self_attn.q_proj.weight") L.wq = &t;
            else if (rest == "attention.wk.weight" || rest == "self_attn.k_proj.weight") L.wk = &t;
            else if (rest == "attention.wv.weight" || rest == "self_attn.v_proj.weight") L.wv = &t;
            else if (rest == "attention.wo.weight" || rest == "self_attn.o_proj.weight") L.wo = &t;
            else if (rest == "feed_forward.w1.weight" || rest == "mlp.gate_proj.weight") L.w1 = &t;
            else if (rest == "feed_forward.w2.weight" || rest == "mlp.down_proj.weight") L.w2 = &t;
            else if (rest == "feed_forward.w3.weight" || rest == "mlp.up_proj.weight") L.w3 = &t;
        }
    }
    if (!*emb || !*norm || !*head || layers->empty()) {
        *err = std::string("not a pmx LLaMA export: missing initializer ") +
            (!*emb ? "[tok_embeddings.weight]" : !*norm ? "[norm.weight]" : !*head ? "[output.weight]" : "[layers.<i>.*]");
        return false;
    }
    for (size_t i = 0; i < layers->size(); ++i) {
        const LayerTensors& L = (*layers)[i];
        const bool qkv = L.wqkv || (L.wq && L.wk && L.wv);
        if (!L.attn_norm || !L.ffn_norm || !qkv || !L.wo || !L.w1 || !L.w2 || !L.w3) {
            *err = "layer " + std::to_string(i) + ": incomplete weight set (need attention_norm, ffn_norm, wqkv or wq/wk/wv, wo, "
                   "feed_forward.w1/w2/w3)";
            return false;
        }
    }
    return true;
}

bool PmxLlama::Open(const std::string& path, std::string* err) {
    model_.reset(new Model());
    if (!model_->Load(path, err)) return false;
    // rank from ".../model_slice_<r>/model.onnx" (resource_manager.cc:280-286)
    rank_ = -1;
    const std::string& dir = model_->dir;
    const size_t p = dir.rfind("model_slice_");
    if (p != std::string::npos && dir.find('/', p) == std::string::npos) {
        char* end = nullptr;
        const long r = strtol(dir.c_str() + p + 12, &end, 10);
        if (end != dir.c_str() + p + 12 && *end == 0) rank_ = (int)r;
    }
    return Interpret(err);
}

bool PmxLlama::Interpret(std::string* err) {
    const Model& m = *model_;
    if (!BindNames(m, &layers_, &embedding_, &final_norm_, &lm_head_, err)) return false;

    std::string params;
    {
        std::ifstream ifs(m.dir + "/../params.json");
        if (ifs.is_open()) {
            std::stringstream ss;
            ss << ifs.rdbuf();
            params = ss.str();
        }
    }
    auto warn = [&](const std::string& w) { warnings_.push_back(w); };
    auto param = [&](const char* key, int64_t* out) { return !params.empty() && JsonInt(params, key, out); };

    // ---- graph-level constants from the pmx nodes
    b2llm_model_desc& d = desc_;
    d = b2llm_model_desc{};
    bool have_eps = false, have_theta = false, have_cache = false;
    int64_t a_heads = 0, a_kv_heads = 0, a_head_dim = 0, a_rotary_dim = 0, a_max_pos = 0;
    for (const Node& n : m.nodes) {
        const std::string op = Lower(n.op_type);
        if (op.find("rmsnorm") != std::string::npos) {
            if (n.Find("eps")) {
                const float eps = n.Float("eps", 1e-5f);
                if (have_eps && eps != d.norm_eps) {
                    *err = "RMSNorm nodes disagree on eps; one engine-wide eps is supported";
                    return false;
                }
                d.norm_eps = eps;
                have_eps = true;
            }
        } else if (op.find("rotary") != std::string::npos) {
            if (n.Find("theta")) {
                d.rope_theta = n.Float("theta", 10000.f);
                have_theta = true;
            }
            if (n.Int("bypass_key", 0) != 0) {
                *err = "RotaryPositionEmbedding.bypass_key = 1 is not supported";
                return false;
            }
            const std::string scaling = n.Str("scaling_type", "");
            if (!scaling.empty() && Lower(scaling) != "none") {
                *err = "RotaryPositionEmbedding.scaling_type [" + scaling + "] is not supported";
                return false;
            }
            a_max_pos = std::max<int64_t>(a_max_pos, n.Int("max_position_embeddings", 0));
            if (n.Int("rotary_dim", 0) > 0) a_rotary_dim = n.Int("rotary_dim", 0); // 0 = the whole head
        } else if (op.find("cacheattention") != std::string::npos || op.find("keyvaluecache") != std::string::npos ||
                   op.find("multiheadattention") != std::string::npos) {
            if (n.Int("is_alibi", 0) != 0) {
                *err = n.op_type + ".is_alibi = 1 is not supported (LLaMA uses rotary embeddings)";
                return false;
            }
            if (n.Find("is_causal") && n.Int("is_causal", 1) == 0) {
                *err = n.op_type + ".is_causal = 0 is not supported (export with --auto_causal 1)";
                return false;
            }
            if (n.Find("num_heads")) a_heads = n.Int("num_heads", 0);
            if (n.Find("num_kv_heads")) a_kv_heads = n.Int("num_kv_heads", 0);
            if (n.Find("head_dim")) a_head_dim = n.Int("head_dim", 0);
            if (n.Find("cache_layout") || n.Find("quant_bit") || n.Find("cache_mode")) {
                b2llm_model_desc c = d;
                c.cache_quant_bit = (int32_t)n.Int("quant_bit", 0);
                c.cache_quant_group = (int32_t)n.Int("quant_group", 8);
                c.cache_mode = (int32_t)n.Int("cache_mode", 0);
                c.cache_layout = (int32_t)n.Int("cache_layout", 0);
                c.page_size = (int32_t)n.Int("page_size", 128);
                if (have_cache && (c.cache_quant_bit != d.cache_quant_bit || c.cache_quant_group != d.cache_quant_group ||
                                   c.cache_mode != d.cache_mode || c.cache_layout != d.cache_layout || c.page_size != d.page_size)) {
                    *err = "KV-cache attributes differ between nodes (quant_bit / quant_group / cache_mode / cache_layout / page_size)";
                    return false;
                }
                d = c;
                have_cache = true;
            }
        } else if (op.find("linear") != std::string::npos) {
            if (n.Int("bias_term", 0) != 0) {
                *err = n.op_type + " [" + n.name + "] has bias_term = 1: not supported";
                return false;
            }
        }
    }
    if (!have_eps) {
        d.norm_eps = 1e-5f;
        warn("no RMSNorm.eps attribute in the graph: using 1e-5");
    }
    if (!have_theta) {
        d.rope_theta = 10000.f;
        warn("no RotaryPositionEmbedding.theta attribute in the graph: using 10000");
    }
    if (!have_cache) {
        int64_t v;
        bool all = true;
        all &= param("cache_quant_bit", &v); d.cache_quant_bit = all ? (int32_t)v : 8;
        d.cache_quant_group = param("cache_quant_group", &v) ? (int32_t)v : 8;
        all &= param("cache_layout", &v); d.cache_layout = all ? (int32_t)v : 0;
        all &= param("cache_mode", &v); d.cache_mode = all ? (int32_t)v : 0;
        d.page_size = param("page_size", &v) ? (int32_t)v : 128;
        warn(all ? "no KV-cache attributes in the graph: taken from params.json"
                 : "no KV-cache attributes in the graph and no params.json beside the slices: quant_bit 8, group 8, layout 0, mode 0");
    }

    // ---- dimensions from the shard shapes
    if (!final_norm_ || final_norm_->dims.size() != 1) {
        *err = "norm.weight must be 1-D, is " + Shape(final_norm_);
        return false;
    }
    const int64_t h = final_norm_->dims[0];
    const LayerTensors& L0 = layers_[0];
    fused_qkv_ = L0.wqkv != nullptr;
    if (!Is2D(L0.wo) || L0.wo->dims[0] != h || h % L0.wo->dims[1] != 0) {
        *err = "layers.0.attention.wo.weight " + Shape(L0.wo) + " is not [hidden = " + std::to_string(h) + ", hidden / tp]";
        return false;
    }
    tp_ = (int)(h / L0.wo->dims[1]);
    const int64_t q_rows = L0.wo->dims[1]; // local q heads * head_dim
    int64_t kv_rows;
    if (fused_qkv_) {
        if (!Is2D(L0.wqkv) || L0.wqkv->dims[1] != h || L0.wqkv->dims[0] <= q_rows || (L0.wqkv->dims[0] - q_rows) % 2) {
            *err = "layers.0.attention.wqkv.weight " + Shape(L0.wqkv) + " is not [(q + 2 kv) rows, hidden]";
            return false;
        }
        kv_rows = (L0.wqkv->dims[0] - q_rows) / 2;
    } else {
        if (!Is2D(L0.wq) || !Is2D(L0.wk) || L0.wq->dims[0] != q_rows || L0.wq->dims[1] != h || L0.wk->dims[1] != h) {
            *err = "layers.0.attention.wq/wk.weight shapes " + Shape(L0.wq) + " / " + Shape(L0.wk) + " do not match wo " + Shape(L0.wo);
            return false;
        }
        kv_rows = L0.wk->dims[0];
    }
    // head_dim: attribute, else params.json num_heads, else refuse
    int64_t D = a_head_dim, v = 0;
    if (D == 0 && param("num_heads", &v) && v > 0 && h % v == 0) D = h / v;
    if (D == 0) {
        *err = "cannot determine head_dim: no head_dim attribute on the attention nodes and no params.json beside the slices";
        return false;
    }
    if (a_rotary_dim && a_rotary_dim != D) {
        *err = "partial rotary embedding (rotary_dim " + std::to_string(a_rotary_dim) + " != head_dim " + std::to_string(D) + ") is not supported";
        return false;
    }
    if (q_rows % D || kv_rows % D || kv_rows == 0 || q_rows % kv_rows) {
        *err = "q rows " + std::to_string(q_rows) + " / kv rows " + std::to_string(kv_rows) + " per rank do not divide into heads of " + std::to_string(D);
        return false;
    }
    const int64_t nq_l = q_rows / D, nkv_l = kv_rows / D;
    if (a_heads && a_heads != nq_l && a_heads != nq_l * tp_) warn("attention num_heads attribute " + std::to_string(a_heads) + " matches neither the local (" + std::to_string(nq_l) + ") nor the global head count");
    if (a_kv_heads && a_kv_heads != nkv_l && a_kv_heads != nkv_l * tp_) warn("attention num_kv_heads attribute " + std::to_string(a_kv_heads) + " matches neither the local (" + std::to_string(nkv_l) + ") nor the global kv head count");
    if (!Is2D(L0.w1) || L0.w1->dims[1] != h) {
        *err = "layers.0.feed_forward.w1.weight " + Shape(L0.w1) + " is not [intermediate / tp, hidden]";
        return false;
    }
    const int64_t I_l = L0.w1->dims[0];
    d.hidden_dim = (int32_t)h;
    d.num_layers = (int32_t)layers_.size();
    d.num_heads = (int32_t)(nq_l * tp_);
    d.num_kv_heads = (int32_t)(nkv_l * tp_);
    d.intermediate_dim = (int32_t)(I_l * tp_);
    d.max_position = (int32_t)std::max<int64_t>(a_max_pos, 16384);
    d.quant_method = B2LLM_QUANT_NONE; // chosen by the engine options (--quant-method), not by the export

    // every layer: same shapes, fp16 / fp32 / bf16 payloads
    for (size_t i = 0; i < layers_.size(); ++i) {
        const LayerTensors& L = layers_[i];
        auto shape_is = [&](const Tensor* t, int64_t r, int64_t c) { return Is2D(t) && t->dims[0] == r && t->dims[1] == c; };
        bool ok = L.attn_norm->dims == std::vector<int64_t>{h} && L.ffn_norm->dims == std::vector<int64_t>{h} &&
            shape_is(L.wo, h, q_rows) && shape_is(L.w1, I_l, h) && shape_is(L.w3, I_l, h) && shape_is(L.w2, h, I_l);
        if (fused_qkv_) ok = ok && shape_is(L.wqkv, q_rows + 2 * kv_rows, h);
        else ok = ok && shape_is(L.wq, q_rows, h) && shape_is(L.wk, kv_rows, h) && shape_is(L.wv, kv_rows, h);
        if (!ok) {
            *err = "layer " + std::to_string(i) + ": weight shapes differ from layer 0's (or mix fused and split qkv)";
            return false;
        }
    }
    // ---- vocabulary and how embedding / lm head are partitioned
    int64_t V = 0;
    if (const Node* n = Consumer(m, lm_head_)) V = n->Int("out_features", 0);
    if (!V) {
        if (const Node* n = Consumer(m, embedding_)) V = n->Int("num_embeddings", 0);
    }
    if (!V && param("vocab_size", &v)) V = v;
    if (!Is2D(lm_head_) || lm_head_->dims[1] != h || !Is2D(embedding_)) {
        *err = "output.weight " + Shape(lm_head_) + " / tok_embeddings.weight " + Shape(embedding_) + " are not 2-D [*, hidden]";
        return false;
    }
    if (!V) {
        V = lm_head_->dims[0];
        if (tp_ > 1) warn("vocabulary size not stated (no out_features / num_embeddings attribute, no params.json): assuming output.weight holds the whole vocabulary");
    }
    if (lm_head_->dims[0] == V) head_split_ = 0;
    else if (lm_head_->dims[0] * tp_ == V) head_split_ = 2;
    else {
        *err = "output.weight " + Shape(lm_head_) + " is neither the whole vocabulary (" + std::to_string(V) + ") nor a 1/" + std::to_string(tp_) + " row slice of it";
        return false;
    }
    if (embedding_->dims[0] == V && embedding_->dims[1] == h) emb_split_ = 0;
    else if (embedding_->dims[0] == V && embedding_->dims[1] * tp_ == h) emb_split_ = 1;
    else if (embedding_->dims[0] * tp_ == V && embedding_->dims[1] == h) emb_split_ = 2;
    else {
        *err = "tok_embeddings.weight " + Shape(embedding_) + " is not [vocab, hidden], [vocab, hidden / tp] or [vocab / tp, hidden]";
        return false;
    }
    d.vocab_size = (int32_t)V;

    // ---- cross-check with params.json (what LLMEngine::Init shapes the KV tensors from, llm_engine.cc:118-169)
    if (!params.empty()) {
        const struct {
            const char* key;
            int64_t mine;
        } checks[] = {{"hidden_dim", d.hidden_dim}, {"num_heads", d.num_heads}, {"num_layers", d.num_layers},
                      {"intermediate_dim", d.intermediate_dim}, {"vocab_size", d.vocab_size},
                      {"cache_quant_bit", d.cache_quant_bit}, {"cache_quant_group", d.cache_quant_group},
                      {"cache_layout", d.cache_layout}, {"cache_mode", d.cache_mode}};
        for (const auto& c : checks) {
            if (param(c.key, &v) && v != c.mine) {
                *err = std::string("params.json says ") + c.key + " = " + std::to_string(v) + " but the exported graph has " + std::to_string(c.mine);
                return false;
            }
        }
        if (param("num_kv_heads", &v) && v != d.num_kv_heads) {
            *err = "params.json says num_kv_heads = " + std::to_string(v) + " but the exported graph has " + std::to_string(d.num_kv_heads);
            return false;
        }
        if (d.cache_mode == 1 && param("page_size", &v) && v != d.page_size) {
            *err = "params.json says page_size = " + std::to_string(v) + " but the exported graph has " + std::to_string(d.page_size);
            return false;
        }
    }
    if (rank_ >= tp_) {
        *err = "model_slice_" + std::to_string(rank_) + " but the shard shapes say tensor-parallel size " + std::to_string(tp_);
        return false;
    }
    return true;
}

const Model* PmxLlama::Sibling(int rank, std::string* err) {
    if (rank == rank_) return model_.get();
    if (rank_ < 0) {
        *err = "a tensor-parallel export must live in <model-dir>/model_slice_<rank>/ so that the other slices can be found";
        return nullptr;
    }
    if (siblings_.size() < (size_t)tp_) siblings_.resize(tp_);
    if (!siblings_[rank]) {
        std::unique_ptr<Model> m(new Model());
        if (!m->Load(model_->dir + "/../model_slice_" + std::to_string(rank) + "/model.onnx", err)) return nullptr;
        siblings_[rank] = std::move(m);
    }
    return siblings_[rank].get();
}

namespace {

// payload of `t` as fp16; converted into `scratch` when the export is fp32 / bf16
const uint16_t* AsHalf(const Tensor* t, std::vector<uint16_t>* scratch, std::string* err) {
    const uint64_t n = t->NumElements();
    if (t->data_type == DT_FLOAT16) return (const uint16_t*)t->data;
    scratch->resize(n);
    if (t->data_type == DT_FLOAT) {
        for (uint64_t i = 0; i < n; ++i) {
            float f;
            memcpy(&f, t->data + i * 4, 4);
            (*scratch)[i] = FloatToHalf(f);
        }
    } else if (t->data_type == DT_BFLOAT16) {
        for (uint64_t i = 0; i < n; ++i) {
            uint16_t b;
            memcpy(&b, t->data + i * 2, 2);
            const uint32_t u = (uint32_t)b << 16;
            float f;
            memcpy(&f, &u, 4);
            (*scratch)[i] = FloatToHalf(f);
        }
    } else {
        *err = "initializer [" + t->name + "]: data type " + std::to_string(t->data_type) + " is not fp16 / fp32 / bf16";
        return nullptr;
    }
    return scratch->data();
}

} // namespace

int32_t PmxLlama::ForEachWeight(const Sink& sink, std::string* err) {
    const b2llm_model_desc& d = desc_;
    const uint64_t h = d.hidden_dim, V = d.vocab_size;
    std::vector<uint16_t> scratch, full;
    int32_t rc = 0;
    auto put = [&](int32_t kind, int32_t layer, const Tensor* t) -> bool {
        const uint16_t* p = AsHalf(t, &scratch, err);
        if (!p) {
            rc = B2LLM_ERR_UNSUPPORTED;
            return false;
        }
        rc = sink(kind, layer, p, t->NumElements(), t->name.c_str());
        return rc == 0;
    };
    // lm head split along vocab (rows [r * V/tp, (r+1) * V/tp) per slice): b2llm's head is vocab-parallel too, the
    // rank's slice goes in as it is and the logits are all-gathered every step (llm_engine.cc:200).
    // embedding (and an lm head the engine keeps whole: vocab / tp not a multiple of 32, B2LLM_TP_HEAD=whole): whole on
    // every rank -- a row gather reads only the step's tokens, so nothing is gained by splitting it; assembled once, here.
    auto put_assembled = [&](int32_t kind, const Tensor* mine, int split) -> bool {
        if (split == 0 || tp_ == 1) return put(kind, 0, mine);
        // the engine's own rule for a vocab-parallel head (csrc/engine.cu: vocab / tp a multiple of 32, B2LLM_TP_HEAD != whole)
        const char* hs = getenv("B2LLM_TP_HEAD");
        if (kind == B2LLM_W_LM_HEAD && split == 2 && (V / tp_) % 32 == 0 && !(hs && hs[0] == 'w')) {
            if (put(kind, 0, mine)) return true;
            rc = 0;  // refused after all: assemble it below
            err->clear();
        }
        full.assign(V * h, 0);
        for (int r = 0; r < tp_; ++r) {
            const Model* m = Sibling(r, err);
            if (!m) {
                rc = B2LLM_ERR_INVALID_VALUE;
                return false;
            }
            std::vector<LayerTensors> ls;
            const Tensor *e = nullptr, *n = nullptr, *o = nullptr;
            if (!BindNames(*m, &ls, &e, &n, &o, err)) {
                rc = B2LLM_ERR_INVALID_VALUE;
                return false;
            }
            const Tensor* piece = kind == B2LLM_W_EMBEDDING ? e : o;
            if (piece->dims != mine->dims) {
                *err = "model_slice_" + std::to_string(r) + " [" + piece->name + "] " + Shape(piece) + " differs from this rank's " + Shape(mine);
                rc = B2LLM_ERR_INVALID_VALUE;
                return false;
            }
            const uint16_t* p = AsHalf(piece, &scratch, err);
            if (!p) {
                rc = B2LLM_ERR_UNSUPPORTED;
                return false;
            }
            if (split == 2) { // rows [r * V/tp, (r+1) * V/tp)
                memcpy(full.data() + (uint64_t)r * (V / tp_) * h, p, (V / tp_) * h * 2);
            } else { // columns [r * h/tp, (r+1) * h/tp)
                const uint64_t hl = h / tp_;
                for (uint64_t row = 0; row < V; ++row) memcpy(full.data() + row * h + r * hl, p + row * hl, hl * 2);
            }
        }
        rc = sink(kind, 0, full.data(), V * h, mine->name.c_str());
        return rc == 0;
    };
    if (!put_assembled(B2LLM_W_EMBEDDING, embedding_, emb_split_)) return rc;
    if (!put(B2LLM_W_FINAL_NORM, 0, final_norm_)) return rc;
    if (!put_assembled(B2LLM_W_LM_HEAD, lm_head_, head_split_)) return rc;
    full.clear();
    full.shrink_to_fit();
    std::vector<uint16_t> qkv;
    for (int32_t l = 0; l < d.num_layers; ++l) {
        const LayerTensors& L = layers_[l];
        if (!put(B2LLM_W_ATTN_NORM, l, L.attn_norm)) return rc;
        if (fused_qkv_) {
            if (!put(B2LLM_W_QKV, l, L.wqkv)) return rc;
        } else {
            qkv.clear();
            for (const Tensor* t : {L.wq, L.wk, L.wv}) {
                const uint16_t* p = AsHalf(t, &scratch, err);
                if (!p) return B2LLM_ERR_UNSUPPORTED;
                qkv.insert(qkv.end(), p, p + t->NumElements());
            }
            rc = sink(B2LLM_W_QKV, l, qkv.data(), qkv.size(), L.wq->name.c_str());
            if (rc) return rc;
        }
        if (!put(B2LLM_W_O, l, L.wo)) return rc;
        if (!put(B2LLM_W_FFN_NORM, l, L.ffn_norm)) return rc;
        if (!put(B2LLM_W_GATE, l, L.w1)) return rc;
        if (!put(B2LLM_W_UP, l, L.w3)) return rc;
        if (!put(B2LLM_W_DOWN, l, L.w2)) return rc;
    }
    return 0;
}

} // namespace b2onnx
