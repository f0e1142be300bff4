// b2pplnn_inspect_model: what RuntimeBuilder::LoadModel + Preprocess would hand to b2llm for a ppl.pmx ONNX export,
// as JSON, WITHOUT touching a device -- the CPU-side check of the model-slice loader (tests/test_pmx_onnx_cpu.py).
// Every weight goes through the same PmxLlama::ForEachWeight the runtime uses; instead of
// b2llm_engine_load_weight_shard the sink records kind, layer, element count and an FNV-1a hash of the fp16 bytes.
#include "pmx_llama.h"

#include <stdio.h>
#include <string.h>

#include <sstream>

namespace {

uint64_t Fnv1a(const void* p, uint64_t n) {
    const uint8_t* b = (const uint8_t*)p;
    uint64_t h = 1469598103934665603ull;
    for (uint64_t i = 0; i < n; ++i) {
        h ^= b[i];
        h *= 1099511628211ull;
    }
    return h;
}

std::string Escape(const std::string& s) {
    std::string o;
    for (char c : s) {
        if (c == '"' || c == '\\') o += '\\';
        if ((unsigned char)c < 0x20) {
            char buf[8];
            snprintf(buf, sizeof(buf), "\\u%04x", c);
            o += buf;
        } else {
            o += c;
        }
    }
    return o;
}

} // namespace

// returns 0 and a JSON object in `out` (NUL terminated, truncated to cap); 2 (RC_INVALID_VALUE) with
// {"error": "..."} when the file is not a loadable pmx LLaMA export
extern "C" __attribute__((visibility("default"))) int32_t b2pplnn_inspect_model(const char* path, char* out, uint64_t cap) {
    if (!path || !out || cap == 0) return 2;
    std::ostringstream js;
    b2onnx::PmxLlama pmx;
    std::string err;
    int32_t rc = 0;
    if (!pmx.Open(path, &err)) {
        js << "{\"error\": \"" << Escape(err) << "\"}";
        rc = 2;
    } else {
        const b2llm_model_desc& d = pmx.desc();
        char eps[32], theta[32];
        snprintf(eps, sizeof(eps), "%.9g", d.norm_eps);
        snprintf(theta, sizeof(theta), "%.9g", d.rope_theta);
        js << "{\"hidden_dim\": " << d.hidden_dim << ", \"intermediate_dim\": " << d.intermediate_dim << ", \"num_layers\": "
           << d.num_layers << ", \"num_heads\": " << d.num_heads << ", \"num_kv_heads\": " << d.num_kv_heads
           << ", \"vocab_size\": " << d.vocab_size << ", \"norm_eps\": " << eps << ", \"rope_theta\": " << theta
           << ", \"cache_quant_bit\": " << d.cache_quant_bit << ", \"cache_quant_group\": " << d.cache_quant_group
           << ", \"cache_layout\": " << d.cache_layout << ", \"cache_mode\": " << d.cache_mode << ", \"page_size\": "
           << d.page_size << ", \"max_position\": " << d.max_position << ", \"tensor_parallel_size\": "
           << pmx.tensor_parallel_size() << ", \"rank\": " << pmx.rank() << ", \"fused_qkv\": " << (pmx.fused_qkv() ? "true" : "false")
           << ", \"producer\": \"" << Escape(pmx.model().producer_name) << "\", \"nodes\": " << pmx.model().nodes.size()
           << ", \"initializers\": " << pmx.model().initializers.size() << ", \"warnings\": [";
        for (size_t i = 0; i < pmx.warnings().size(); ++i) js << (i ? ", " : "") << "\"" << Escape(pmx.warnings()[i]) << "\"";
        js << "], \"weights\": [";
        bool first = true;
        rc = pmx.ForEachWeight(
            [&](int32_t kind, int32_t layer, const void* fp16, uint64_t n, const char* name) {
                js << (first ? "" : ", ") << "{\"kind\": " << kind << ", \"layer\": " << layer << ", \"elements\": " << n
                   << ", \"name\": \"" << Escape(name) << "\", \"fnv1a\": \"" << std::hex << Fnv1a(fp16, n * 2) << std::dec << "\"}";
                first = false;
                return 0;
            },
            &err);
        js << "]";
        if (rc != 0) js << ", \"error\": \"" << Escape(err) << "\"";
        js << "}";
    }
    const std::string s = js.str();
    const uint64_t n = s.size() < cap - 1 ? s.size() : cap - 1;
    memcpy(out, s.data(), n);
    out[n] = 0;
    return rc;
}
