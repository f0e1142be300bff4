// The ppl.nn plugin surface of the reference, implemented over the b2llm C ABI (include/b2llm.h).
//
// ppl.llm.serving reaches its model ONLY through ppl::nn::{Engine, Runtime, Tensor, DeviceContext} and the
// onnx::RuntimeBuilder: 11 input tensors and 1 output addressed by index (llm_engine.h:124-138),
// `Runtime::Run()` = the whole forward (llm_engine.cc:113-116).  This file provides those classes so that the
// reference's own llm_engine.cc / resource_manager.cc / post_processor.cc / llm_generator.cc compile and link
// UNCHANGED against libb2llm.so instead of ppl.nn + ppl.llm.kernel.cuda.
//
//   EngineFactory::Create            -> B200Engine (options, NCCL communicator, Configure keys)
//   RuntimeBuilder::LoadModel        -> reads model_slice_<rank>/model.onnx: a ppl.pmx ONNX export (onnx_model.cc,
//                                       pmx_llama.cc) or a b2llm model-slice descriptor
//   RuntimeBuilder::CreateRuntime    -> b2llm_engine_create + weights (ONNX initializers, synthetic seed or fp16 blob)
//   Tensor::CopyFromHostAsync        -> cudaMemcpyAsync on the rank's stream (or host memcpy for the 3 scalars)
//   Runtime::Run                     -> b2llm_engine_reserve + b2llm_engine_bind_kv + b2llm_engine_forward
//
// No computation happens here and nothing falls back to the CPU: without a B200 every entry point returns a
// RetCode error.
#include "b2llm.h"
#include "pmx_llama.h"

#include "ppl/common/log.h"
#include "ppl/nn/engines/llm_cuda/engine_factory.h"
#include "ppl/nn/engines/llm_cuda/options.h"
#include "ppl/nn/models/onnx/runtime_builder.h"
#include "ppl/nn/models/onnx/runtime_builder_factory.h"
#include "ppl/nn/runtime/runtime.h"

#include <cuda_runtime.h>
#include <fcntl.h>
#include <stdarg.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#ifdef PPLNN_CUDA_ENABLE_NCCL
#include "nccl.h"
#endif

using namespace ppl::common;

namespace ppl { namespace nn {

namespace {

RetCode FromB2(int32_t rc, const char* what) {
    if (rc != B2LLM_OK) {
        LOG(ERROR) << what << " failed: " << b2llm_last_error();
    }
    return (RetCode)rc; // values are shared (ppl/common/retcode.h)
}

// ------------------------------------------------------------------------------------- device contexts
class CudaDeviceContext final : public DeviceContext {
public:
    CudaDeviceContext(int device_id, cudaStream_t stream) : device_id_(device_id), stream_(stream) {
        memcpy(type_.str, "cuda", 4);
    }
    const Type& GetType() const override {
        return type_;
    }
    RetCode Configure(uint32_t option, ...) override {
        if (option != llm::cuda::DEV_CONF_GET_STREAM) {
            LOG(ERROR) << "cuda device context: unknown option [" << option << "]";
            return RC_INVALID_VALUE;
        }
        va_list args;
        va_start(args, option);
        cudaStream_t* out = va_arg(args, cudaStream_t*);
        va_end(args);
        *out = stream_;
        return RC_SUCCESS;
    }
    cudaStream_t stream() const {
        return stream_;
    }
    int device_id() const {
        return device_id_;
    }

private:
    Type type_;
    int device_id_;
    cudaStream_t stream_;
};

class HostDeviceContext final : public DeviceContext {
public:
    HostDeviceContext() {
        memcpy(type_.str, "cpu", 3);
    }
    const Type& GetType() const override {
        return type_;
    }
    RetCode Configure(uint32_t, ...) override {
        return RC_UNSUPPORTED;
    }

private:
    Type type_;
};

bool IsHost(const DeviceContext* ctx) {
    return ctx && ctx->GetType().str[1] == 'p'; // "cpu"
}

// ------------------------------------------------------------------------------------- engine
class B200Engine final : public Engine {
public:
    explicit B200Engine(const llm::cuda::EngineOptions& o)
        : options_(o), device_ctx_(new CudaDeviceContext((int)o.device_id, o.runtime_stream)) {}

    const char* GetName() const override {
        return "llm_cuda(b2llm sm_100a)";
    }

    RetCode Configure(uint32_t option, ...) override {
        va_list args;
        va_start(args, option);
        RetCode rc = RC_SUCCESS;
        switch (option) {
            // The three algorithm switches select among Ampere decode-attention kernels upstream; b2llm has one
            // tensor-core split-KV kernel for every case, so they are accepted and recorded only.
            case llm::cuda::ENGINE_CONF_DECODING_SHM_MHA:
            case llm::cuda::ENGINE_CONF_DECODING_INF_MHA:
            case llm::cuda::ENGINE_CONF_DECODING_INF_GQA:
            case llm::cuda::ENGINE_CONF_DECODING_ATTN_TPB:
            case llm::cuda::ENGINE_CONF_GRAPH_FUSION:
                conf_[option] = va_arg(args, uint32_t);
                break;
            case llm::cuda::ENGINE_CONF_DECODING_ATTN_SPLIT_K: {
                const uint32_t v = va_arg(args, uint32_t);
                if (v > 2) {
                    LOG(ERROR) << "ENGINE_CONF_DECODING_ATTN_SPLIT_K must be 0, 1 or 2";
                    rc = RC_INVALID_VALUE;
                } else {
                    conf_[option] = v;
                    if (b2_) rc = FromB2(b2llm_engine_configure(b2_, B2LLM_CONF_DECODING_ATTN_SPLIT_K, v), "configure split-k");
                }
                break;
            }
            case llm::cuda::ENGINE_CONF_SET_TP_NCCL_COMM:
                nccl_comm_ = va_arg(args, void*);
                break;
            case llm::cuda::ENGINE_CONF_CACHE_PREFILL:
                cache_prefill_ = va_arg(args, uint32_t) != 0;
                break;
            default:
                LOG(ERROR) << "engine Configure: unknown option [" << option << "]";
                rc = RC_INVALID_VALUE;
        }
        va_end(args);
        return rc;
    }

    const llm::cuda::EngineOptions& options() const {
        return options_;
    }
    void* nccl_comm() const {
        return nccl_comm_;
    }
    bool cache_prefill() const {
        return cache_prefill_;
    }
    uint32_t conf(uint32_t key, uint32_t dflt) const {
        auto it = conf_.find(key);
        return it == conf_.end() ? dflt : it->second;
    }
    CudaDeviceContext* device_context() const {
        return device_ctx_.get();
    }
    void attach(b2llm_engine* e) {
        b2_ = e;
    }

private:
    llm::cuda::EngineOptions options_;
    std::unique_ptr<CudaDeviceContext> device_ctx_;
    std::map<uint32_t, uint32_t> conf_;
    void* nccl_comm_ = nullptr;
    bool cache_prefill_ = false;
    b2llm_engine* b2_ = nullptr; // owned by the runtime
};

// ------------------------------------------------------------------------------------- tensors
class B200Tensor final : public Tensor {
public:
    B200Tensor(const char* name, datatype_t dt) : name_(name) {
        shape_.SetDataType(dt);
        shape_.Reshape({0});
    }
    ~B200Tensor() override {
        if (owned_) cudaFree(owned_);
    }
    const char* GetName() const override {
        return name_.c_str();
    }
    TensorShape* GetShape() const override {
        return &shape_;
    }
    RetCode SetDeviceContext(DeviceContext* ctx) override {
        ctx_ = ctx;
        return RC_SUCCESS;
    }
    DeviceContext* GetDeviceContext() const override {
        return ctx_;
    }
    void SetBufferPtr(void* p) override {
        external_ = p;
    }
    void* GetBufferPtr() const override {
        if (external_) return external_;
        return IsHost(ctx_) ? (void*)host_.data() : owned_;
    }
    // ppl.nn returns the buffer to its pool; here the allocation is kept for the next step (same effect for the
    // caller: the tensor holds no data until the next CopyFromHostAsync / Run)
    void FreeBuffer() override {
        valid_bytes_ = 0;
    }
    RetCode ReallocBuffer() override {
        return Ensure(shape_.CalcBytesIncludingPadding());
    }

    RetCode CopyFromHostAsync(const void* src) override {
        const uint64_t bytes = shape_.CalcBytesIncludingPadding();
        if (IsHost(ctx_)) {
            host_.resize(bytes);
            memcpy(host_.data(), src, bytes);
            valid_bytes_ = bytes;
            return RC_SUCCESS;
        }
        if (external_) {
            LOG(ERROR) << "tensor [" << name_ << "]: CopyFromHost on a caller-owned buffer";
            return RC_INVALID_VALUE;
        }
        auto rc = Ensure(bytes);
        if (rc != RC_SUCCESS) return rc;
        if (bytes) {
            const cudaError_t err = cudaMemcpyAsync(owned_, src, bytes, cudaMemcpyHostToDevice, Stream());
            if (err != cudaSuccess) {
                LOG(ERROR) << "tensor [" << name_ << "]: cudaMemcpyAsync H2D failed: " << cudaGetErrorString(err);
                return RC_DEVICE_MEMORY_ERROR;
            }
        }
        valid_bytes_ = bytes;
        return RC_SUCCESS;
    }
    RetCode CopyFromHost(const void* src) override {
        auto rc = CopyFromHostAsync(src);
        if (rc == RC_SUCCESS && !IsHost(ctx_)) {
            if (cudaStreamSynchronize(Stream()) != cudaSuccess) rc = RC_DEVICE_RUNTIME_ERROR;
        }
        return rc;
    }
    RetCode CopyToHostAsync(void* dst) const override {
        const uint64_t bytes = shape_.CalcBytesIncludingPadding();
        if (IsHost(ctx_)) {
            memcpy(dst, host_.data(), std::min<uint64_t>(bytes, host_.size()));
            return RC_SUCCESS;
        }
        const void* src = GetBufferPtr();
        if (!src && bytes) return RC_INVALID_VALUE;
        if (bytes && cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, Stream()) != cudaSuccess) {
            return RC_DEVICE_MEMORY_ERROR;
        }
        return RC_SUCCESS;
    }
    RetCode CopyToHost(void* dst) const override {
        auto rc = CopyToHostAsync(dst);
        if (rc == RC_SUCCESS && !IsHost(ctx_)) {
            if (cudaStreamSynchronize(Stream()) != cudaSuccess) rc = RC_DEVICE_RUNTIME_ERROR;
        }
        return rc;
    }
    RetCode ConvertToHost(void* dst, const TensorShape& dst_desc) const override {
        if (dst_desc.GetDataType() != shape_.GetDataType()) {
            LOG(ERROR) << "tensor [" << name_ << "]: ConvertToHost with a data type conversion is not supported";
            return RC_UNSUPPORTED;
        }
        return CopyToHost(dst);
    }

    // runtime-side helpers
    int64_t Dim(uint32_t i) const {
        return i < shape_.GetDimCount() ? shape_.GetDim(i) : 0;
    }
    int64_t HostScalar() const {
        int64_t v = 0;
        if (host_.size() >= sizeof(v)) memcpy(&v, host_.data(), sizeof(v));
        return v;
    }
    bool HasData() const {
        return external_ || valid_bytes_ > 0;
    }
    void PointAt(void* p) { // output tensor: buffer owned by the engine
        external_ = p;
    }

private:
    cudaStream_t Stream() const {
        auto* c = dynamic_cast<CudaDeviceContext*>(ctx_);
        return c ? c->stream() : (cudaStream_t)0;
    }
    RetCode Ensure(uint64_t bytes) {
        if (bytes <= capacity_) return RC_SUCCESS;
        if (owned_) {
            cudaStreamSynchronize(Stream()); // a kernel of the previous step may still read the old buffer
            cudaFree(owned_);
            owned_ = nullptr;
            capacity_ = 0;
        }
        const uint64_t want = bytes + bytes / 2 + 256;
        if (cudaMalloc(&owned_, want) != cudaSuccess) {
            cudaGetLastError();
            LOG(ERROR) << "tensor [" << name_ << "]: cudaMalloc of " << want << " bytes failed";
            return RC_OUT_OF_MEMORY;
        }
        capacity_ = want;
        return RC_SUCCESS;
    }

    std::string name_;
    mutable TensorShape shape_;
    DeviceContext* ctx_ = nullptr;
    void* external_ = nullptr;
    void* owned_ = nullptr;
    uint64_t capacity_ = 0, valid_bytes_ = 0;
    std::vector<char> host_;
};

// ------------------------------------------------------------------------------------- model slice descriptor
struct SliceDesc {
    b2llm_model_desc d{};
    int tp = 1, rank = 0;
    std::string weights; // "synthetic:<seed>", "file:<path relative to the descriptor>" or "onnx" (pmx below)
    std::string dir;
    std::shared_ptr<b2onnx::PmxLlama> pmx; // the parsed ppl.pmx export when model.onnx is a real ONNX file
};

// model.onnx of a ppl.pmx export (docs/llama_guide.md:12-36): dimensions and constants from the graph, weights
// from its initializers
bool ParsePmxOnnx(const char* path, SliceDesc* out) {
    std::shared_ptr<b2onnx::PmxLlama> pmx(new b2onnx::PmxLlama());
    std::string err;
    if (!pmx->Open(path, &err)) {
        LOG(ERROR) << "model [" << path << "]: " << err;
        return false;
    }
    for (const auto& w : pmx->warnings()) LOG(WARNING) << "model [" << path << "]: " << w;
    out->d = pmx->desc();
    out->tp = pmx->tensor_parallel_size();
    out->rank = pmx->rank() < 0 ? 0 : pmx->rank();
    out->weights = "onnx";
    out->dir = pmx->model().dir;
    out->pmx = pmx;
    LOG(INFO) << "model [" << path << "]: pmx LLaMA export, producer [" << pmx->model().producer_name << "], "
              << out->d.num_layers << " layers, hidden " << out->d.hidden_dim << ", heads " << out->d.num_heads << "/"
              << out->d.num_kv_heads << ", tensor-parallel slice " << out->rank << " of " << out->tp;
    return true;
}

bool ParseSlice(const char* path, SliceDesc* out) {
    std::ifstream ifs(path);
    if (!ifs.is_open()) {
        LOG(ERROR) << "model slice [" << path << "]: cannot open";
        return false;
    }
    char magic[20] = {0};
    ifs.read(magic, 19);
    if (strncmp(magic, "b2llm-model-slice 1", 19) != 0) {
        ifs.close();
        return ParsePmxOnnx(path, out); // a protobuf: the ppl.pmx export the reference's docs describe
    }
    ifs.seekg(0);
    std::string line;
    std::getline(ifs, line);
    std::map<std::string, std::string> kv;
    while (std::getline(ifs, line)) {
        const auto hash = line.find('#');
        if (hash != std::string::npos) line.resize(hash);
        const auto eq = line.find('=');
        if (eq == std::string::npos) continue;
        auto trim = [](std::string s) {
            const auto b = s.find_first_not_of(" \t\r");
            const auto e = s.find_last_not_of(" \t\r");
            return b == std::string::npos ? std::string() : s.substr(b, e - b + 1);
        };
        kv[trim(line.substr(0, eq))] = trim(line.substr(eq + 1));
    }
    auto geti = [&](const char* k, int32_t* v, bool required) {
        auto it = kv.find(k);
        if (it == kv.end()) {
            if (required) LOG(ERROR) << "model slice [" << path << "]: missing key [" << k << "]";
            return !required;
        }
        *v = (int32_t)atoll(it->second.c_str());
        return true;
    };
    b2llm_model_desc& d = out->d;
    d.norm_eps = 1e-5f;
    d.rope_theta = 10000.f;
    d.cache_quant_bit = 8;
    d.cache_quant_group = 8;
    d.max_position = 4096;
    bool ok = geti("hidden_dim", &d.hidden_dim, true) & geti("intermediate_dim", &d.intermediate_dim, true) &
        geti("num_layers", &d.num_layers, true) & geti("num_heads", &d.num_heads, true) &
        geti("vocab_size", &d.vocab_size, true) & geti("cache_layout", &d.cache_layout, true) &
        geti("cache_mode", &d.cache_mode, true);
    if (!ok) return false;
    d.num_kv_heads = d.num_heads;
    geti("num_kv_heads", &d.num_kv_heads, false);
    geti("cache_quant_bit", &d.cache_quant_bit, false);
    geti("cache_quant_group", &d.cache_quant_group, false);
    geti("page_size", &d.page_size, false);
    geti("max_position", &d.max_position, false);
    int32_t tp = 1, rank = 0;
    geti("tensor_parallel_size", &tp, false);
    geti("rank", &rank, false);
    out->tp = tp;
    out->rank = rank;
    if (kv.count("norm_eps")) d.norm_eps = (float)atof(kv["norm_eps"].c_str());
    if (kv.count("rope_theta")) d.rope_theta = (float)atof(kv["rope_theta"].c_str());
    out->weights = kv.count("weights") ? kv["weights"] : "synthetic:45568";
    const std::string p(path);
    const auto slash = p.rfind('/');
    out->dir = slash == std::string::npos ? "." : p.substr(0, slash);
    return true;
}

// fp16 blob: embedding, final_norm, lm_head, then per layer attn_norm, wqkv, wo, ffn_norm, wgate, wup, wdown --
// FULL (unsharded) tensors; b2llm_engine_load_weight keeps this rank's slice.
RetCode LoadWeightBlob(b2llm_engine* e, const b2llm_model_desc& d, const std::string& file) {
    const int fd = open(file.c_str(), O_RDONLY);
    if (fd < 0) {
        LOG(ERROR) << "weights [" << file << "]: cannot open";
        return RC_NOT_FOUND;
    }
    struct stat st;
    fstat(fd, &st);
    const uint64_t h = d.hidden_dim, D = h / d.num_heads, V = d.vocab_size, I = d.intermediate_dim;
    const uint64_t nqkv = (uint64_t)(d.num_heads + 2 * d.num_kv_heads) * D;
    const uint64_t per_layer = h + nqkv * h + h * h + h + 3 * I * h;
    const uint64_t total = V * h + h + V * h + (uint64_t)d.num_layers * per_layer;
    if ((uint64_t)st.st_size != total * 2) {
        LOG(ERROR) << "weights [" << file << "]: expected " << total * 2 << " bytes, file has " << st.st_size;
        close(fd);
        return RC_INVALID_VALUE;
    }
    void* m = mmap(nullptr, st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return RC_OTHER_ERROR;
    const uint16_t* p = (const uint16_t*)m;
    int32_t rc = B2LLM_OK;
    auto put = [&](int32_t kind, int32_t layer, uint64_t n) {
        if (rc == B2LLM_OK) rc = b2llm_engine_load_weight(e, kind, layer, p, n);
        p += n;
    };
    put(B2LLM_W_EMBEDDING, 0, V * h);
    put(B2LLM_W_FINAL_NORM, 0, h);
    put(B2LLM_W_LM_HEAD, 0, V * h);
    for (int l = 0; l < d.num_layers; ++l) {
        put(B2LLM_W_ATTN_NORM, l, h);
        put(B2LLM_W_QKV, l, nqkv * h);
        put(B2LLM_W_O, l, h * h);
        put(B2LLM_W_FFN_NORM, l, h);
        put(B2LLM_W_GATE, l, I * h);
        put(B2LLM_W_UP, l, I * h);
        put(B2LLM_W_DOWN, l, h * I);
    }
    munmap(m, st.st_size);
    return FromB2(rc, "b2llm_engine_load_weight");
}

// ------------------------------------------------------------------------------------- runtime
enum { IN_TOKEN_IDS = 0, IN_ATTN_MASK, IN_SEQ_STARTS, IN_KV_STARTS, IN_CACHE_INDICES, IN_DECODING_BATCHES, IN_START_POS,
       IN_MAX_SEQ_LEN, IN_MAX_KV_LEN, IN_KV_CACHE, IN_KV_SCALE, IN_COUNT };

class B200Runtime final : public Runtime {
public:
    B200Runtime(B200Engine* engine, const SliceDesc& slice, b2llm_engine* b2) : engine_(engine), slice_(slice), b2_(b2) {
        static const struct {
            const char* name;
            datatype_t dt;
        } defs[IN_COUNT] = {{"token_ids", DATATYPE_INT64},      {"attn_mask", DATATYPE_FLOAT16},  {"seq_starts", DATATYPE_INT64},
                            {"kv_starts", DATATYPE_INT64},      {"cache_indices", DATATYPE_INT64}, {"decoding_batches", DATATYPE_INT64},
                            {"start_pos", DATATYPE_INT64},      {"max_seq_len", DATATYPE_INT64},  {"max_kv_len", DATATYPE_INT64},
                            {"kv_cache", DATATYPE_INT8},        {"kv_scale", DATATYPE_FLOAT16}};
        for (int i = 0; i < IN_COUNT; ++i)
            inputs_.emplace_back(new B200Tensor(defs[i].name, (i == IN_KV_CACHE && slice_.d.cache_quant_bit == 0) ? DATATYPE_FLOAT16 : defs[i].dt));
        for (int i : {IN_DECODING_BATCHES, IN_MAX_SEQ_LEN, IN_MAX_KV_LEN}) inputs_[i]->GetShape()->ReshapeAsScalar();
        logits_.reset(new B200Tensor("logits", DATATYPE_FLOAT32));
        logits_->GetShape()->Reshape({0, slice_.d.vocab_size});
        engine_->attach(b2_);
        b2llm_engine_configure(b2_, B2LLM_CONF_DECODING_ATTN_SPLIT_K,
                               engine_->conf(llm::cuda::ENGINE_CONF_DECODING_ATTN_SPLIT_K, 1));
    }
    ~B200Runtime() override {
        engine_->attach(nullptr);
        b2llm_engine_destroy(b2_);
    }

    uint32_t GetInputCount() const override {
        return slice_.d.cache_quant_bit > 0 ? IN_COUNT : IN_COUNT - 1;
    }
    Tensor* GetInputTensor(uint32_t idx) const override {
        return idx < GetInputCount() ? inputs_[idx].get() : nullptr;
    }
    uint32_t GetOutputCount() const override {
        return 1;
    }
    Tensor* GetOutputTensor(uint32_t idx) const override {
        return idx == 0 ? logits_.get() : nullptr;
    }
    uint32_t GetDeviceContextCount() const override {
        return 1;
    }
    DeviceContext* GetDeviceContext(uint32_t idx) const override {
        return idx == 0 ? engine_->device_context() : nullptr;
    }
    RetCode Configure(uint32_t, ...) override {
        return RC_UNSUPPORTED;
    }

    RetCode Run() override {
        const b2llm_model_desc& d = slice_.d;
        B200Tensor* tok = inputs_[IN_TOKEN_IDS].get();
        B200Tensor* sp = inputs_[IN_START_POS].get();
        B200Tensor* idx = inputs_[IN_CACHE_INDICES].get();
        B200Tensor* kv = inputs_[IN_KV_CACHE].get();
        B200Tensor* ks = inputs_[IN_KV_SCALE].get();
        for (int i : {IN_TOKEN_IDS, IN_SEQ_STARTS, IN_KV_STARTS, IN_START_POS, IN_CACHE_INDICES}) {
            if (!inputs_[i]->HasData()) {
                LOG(ERROR) << "Run: input [" << inputs_[i]->GetName() << "] has not been set";
                return RC_INVALID_VALUE;
            }
        }
        // cache_quant_bit 0: fp16 cache, input 10 (kv_scale) does not exist (llm_engine.h:134-136, 145-147)
        if (!kv->GetBufferPtr() || (d.cache_quant_bit > 0 && !ks->GetBufferPtr())) {
            LOG(ERROR) << "Run: kv_cache / kv_scale buffers have not been bound (Tensor::SetBufferPtr)";
            return RC_INVALID_VALUE;
        }
        // kv_cache shape (LLMEngine::Init, llm_engine.cc:118-169) carries kv_cache_max_tokens at a layout-dependent axis
        static const uint32_t token_axis[4] = {0, 1, 2, 3};
        const uint64_t max_tokens = (uint64_t)kv->Dim(token_axis[d.cache_layout]);
        if (max_tokens == 0) {
            LOG(ERROR) << "Run: kv_cache tensor has no shape (LLMEngine::Init not called?)";
            return RC_INVALID_VALUE;
        }
        if (kv->GetBufferPtr() != bound_kv_ || ks->GetBufferPtr() != bound_ks_ || max_tokens != bound_tokens_) {
            auto rc = FromB2(b2llm_engine_bind_kv(b2_, kv->GetBufferPtr(), ks->GetBufferPtr(), max_tokens), "b2llm_engine_bind_kv");
            if (rc != RC_SUCCESS) return rc;
            bound_kv_ = kv->GetBufferPtr();
            bound_ks_ = ks->GetBufferPtr();
            bound_tokens_ = max_tokens;
        }
        b2llm_step st{};
        st.token_ids = (const int64_t*)tok->GetBufferPtr();
        st.seq_starts = (const int64_t*)inputs_[IN_SEQ_STARTS]->GetBufferPtr();
        st.kv_starts = (const int64_t*)inputs_[IN_KV_STARTS]->GetBufferPtr();
        st.cache_indices = (const int64_t*)idx->GetBufferPtr();
        st.start_pos = (const int64_t*)sp->GetBufferPtr();
        st.num_tokens = tok->Dim(0);
        st.batch = sp->Dim(0);
        st.decoding_batches = inputs_[IN_DECODING_BATCHES]->HostScalar();
        st.max_seq_len = inputs_[IN_MAX_SEQ_LEN]->HostScalar();
        st.max_kv_len = inputs_[IN_MAX_KV_LEN]->HostScalar();
        st.max_pages = d.cache_mode == 1 ? idx->Dim(1) : 0;
        st.cache_prefill = engine_->cache_prefill() ? 1 : 0;
        auto rc = FromB2(b2llm_engine_reserve(b2_, std::max<int64_t>(st.num_tokens, 1), std::max<int64_t>(st.batch, 1)),
                         "b2llm_engine_reserve");
        if (rc != RC_SUCCESS) return rc;
        float* logits = nullptr;
        int64_t stride = 0;
        rc = FromB2(b2llm_engine_forward(b2_, &st, &logits, &stride), "b2llm_engine_forward");
        if (rc != RC_SUCCESS) return rc;
        logits_->GetShape()->Reshape({st.batch, stride});
        logits_->PointAt(logits);
        return RC_SUCCESS;
    }

private:
    B200Engine* engine_;
    SliceDesc slice_;
    b2llm_engine* b2_;
    std::vector<std::unique_ptr<B200Tensor>> inputs_;
    std::unique_ptr<B200Tensor> logits_;
    void* bound_kv_ = nullptr;
    void* bound_ks_ = nullptr;
    uint64_t bound_tokens_ = 0;
};

// ------------------------------------------------------------------------------------- builder
class B200RuntimeBuilder final : public onnx::RuntimeBuilder {
public:
    RetCode LoadModel(const char* model_file) override {
        loaded_ = ParseSlice(model_file, &slice_);
        return loaded_ ? RC_SUCCESS : RC_INVALID_VALUE;
    }
    RetCode SetResources(const Resources& r) override {
        if (r.engine_num != 1 || !r.engines || !dynamic_cast<B200Engine*>(r.engines[0])) {
            LOG(ERROR) << "RuntimeBuilder::SetResources: exactly one llm_cuda engine expected";
            return RC_INVALID_VALUE;
        }
        engine_ = static_cast<B200Engine*>(r.engines[0]);
        return RC_SUCCESS;
    }
    RetCode Preprocess() override {
        if (!loaded_ || !engine_) return RC_INVALID_VALUE;
        b2llm_model_desc d = slice_.d;
        const auto& o = engine_->options();
        d.quant_method = o.quant_method == llm::cuda::QUANT_METHOD_ONLINE_I8I8 ? B2LLM_QUANT_ONLINE_I8I8 : B2LLM_QUANT_NONE;
        d.max_tokens_per_step = 256; // grown on demand by Run() (b2llm_engine_reserve)
        d.max_running_batch = 16;
        int tp = 1, rank = 0;
        void* comm = engine_->nccl_comm();
#ifdef PPLNN_CUDA_ENABLE_NCCL
        if (comm) {
            ncclCommCount((ncclComm_t)comm, &tp);
            ncclCommUserRank((ncclComm_t)comm, &rank);
        }
#endif
        if (tp != slice_.tp || rank != slice_.rank) {
            LOG(ERROR) << "model slice is rank " << slice_.rank << " of " << slice_.tp << " but the engine is rank " << rank
                       << " of " << tp << " (--tensor-parallel-size must match the export, as in docs/llama_guide.md:60-73)";
            return RC_INVALID_VALUE;
        }
        b2llm_engine* e = nullptr;
        auto rc = FromB2(b2llm_engine_create(&d, rank, tp, tp > 1 ? comm : nullptr, (void*)o.runtime_stream, &e),
                         "b2llm_engine_create");
        if (rc != RC_SUCCESS) return rc;
        if (slice_.weights.rfind("synthetic:", 0) == 0) {
            rc = FromB2(b2llm_engine_random_init(e, strtoull(slice_.weights.c_str() + 10, nullptr, 0)), "b2llm_engine_random_init");
        } else if (slice_.weights.rfind("file:", 0) == 0) {
            rc = LoadWeightBlob(e, d, slice_.dir + "/" + slice_.weights.substr(5));
        } else if (slice_.pmx) {
            std::string err;
            const int32_t brc = slice_.pmx->ForEachWeight(
                [&](int32_t kind, int32_t layer, const void* fp16, uint64_t n, const char*) {
                    return b2llm_engine_load_weight_shard(e, kind, layer, fp16, n);
                },
                &err);
            if (brc != B2LLM_OK && !err.empty()) LOG(ERROR) << "model [" << slice_.pmx->model().path << "]: " << err;
            rc = FromB2(brc, "loading the ONNX initializers (b2llm_engine_load_weight_shard)");
            slice_.pmx.reset(); // unmap the export; the weights now live in HBM
        } else {
            LOG(ERROR) << "model slice: unknown weights source [" << slice_.weights << "]";
            rc = RC_INVALID_VALUE;
        }
        if (rc != RC_SUCCESS) {
            b2llm_engine_destroy(e);
            return rc;
        }
        b2_ = e;
        slice_.d = d;
        return RC_SUCCESS;
    }
    Runtime* CreateRuntime() const override {
        if (!b2_) return nullptr;
        auto* rt = new B200Runtime(engine_, slice_, b2_);
        b2_ = nullptr; // ownership moved
        return rt;
    }
    ~B200RuntimeBuilder() override {
        if (b2_) b2llm_engine_destroy(b2_);
    }

private:
    SliceDesc slice_;
    bool loaded_ = false;
    B200Engine* engine_ = nullptr;
    mutable b2llm_engine* b2_ = nullptr;
};

} // namespace

// ------------------------------------------------------------------------------------- factories
namespace llm { namespace cuda {

Engine* EngineFactory::Create(const EngineOptions& options) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || (int)options.device_id >= n) {
        cudaGetLastError();
        LOG(ERROR) << "EngineFactory::Create: device [" << options.device_id << "] not available (" << n
                   << " CUDA device(s)); b2llm has no CPU path";
        return nullptr;
    }
    if (options.cublas_layout_hint != CUBLAS_LAYOUT_DEFAULT) {
        LOG(WARNING) << "--cublas-layout-hint is an Ampere cuBLASLt IMMA layout knob; ignored by the tcgen05 GEMMs";
    }
    return new B200Engine(options);
}

DeviceContext* EngineFactory::CreateDeviceContext(const DeviceOptions& o) {
    return new CudaDeviceContext((int)o.device_id, o.stream);
}

DeviceContext* EngineFactory::CreateHostDeviceContext(const HostDeviceOptions&) {
    return new HostDeviceContext();
}

}} // namespace llm::cuda

namespace onnx {
RuntimeBuilder* RuntimeBuilderFactory::Create() {
    return new B200RuntimeBuilder();
}
} // namespace onnx

}} // namespace ppl::nn
