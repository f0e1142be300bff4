// The engine behind the C ABI: weights, activation buffers, staged step inputs and the per-step
// forward (the B200-native replacement of ppl::nn::Runtime::Run(), src/engine/llm_engine.cc:113-116).
//
// Per layer (decode and prefill share the code; T = tokens of the step):
//   rmsnorm+quant -> W8A8 qkv GEMM -> rope + int8 KV append -> attention -> quant ->
//   W8A8 o_proj GEMM (+residual | allreduce) -> rmsnorm+quant -> W8A8 gate_up GEMM with SwiGLU
//   epilogue -> quant -> W8A8 down GEMM (+residual | allreduce)
// then last-token gather -> rmsnorm -> fp16 lm_head GEMM -> fp32 logits [batch, vocab].
#include <dlfcn.h>
#include <math.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "gemm_epilogue.cuh"

namespace b2llm {

thread_local std::string g_last_error;
thread_local int64_t g_launch_count = 0;
void set_last_error(const std::string& msg) { g_last_error = msg; }

namespace {

// ---- NCCL through dlopen: no link-time dependency; the process' already-loaded libnccl is reused
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_allgather_fn)(const void*, void*, size_t, int, void*, cudaStream_t);
nccl_allreduce_fn g_nccl_allreduce = nullptr;
nccl_allgather_fn g_nccl_allgather = nullptr;
std::once_flag g_nccl_once;
void load_nccl() {
    std::call_once(g_nccl_once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) g_nccl_allreduce = (nccl_allreduce_fn)dlsym(h, "ncclAllReduce");
        if (h) g_nccl_allgather = (nccl_allgather_fn)dlsym(h, "ncclAllGather");
    });
}
constexpr int kNcclFloat16 = 6, kNcclFloat32 = 7, kNcclSum = 0;

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    bool owned = true;
    void view(void* ptr, size_t n) {  // a window of memory somebody else owns (the TP communication buffer)
        release();
        p = ptr;
        bytes = n;
        owned = false;
    }
    int32_t ensure(size_t n) {
        if (n <= bytes) return B2LLM_OK;
        if (p && owned) cudaFree(p);
        owned = true;
        p = nullptr;
        bytes = 0;
        if (cudaMalloc(&p, n) != cudaSuccess) {
            set_last_error("cudaMalloc of " + std::to_string(n) + " bytes failed");
            cudaGetLastError();
            return B2LLM_ERR_OUT_OF_MEMORY;
        }
        bytes = n;
        return B2LLM_OK;
    }
    void release() {
        if (p && owned) cudaFree(p);
        p = nullptr;
        bytes = 0;
        owned = true;
    }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

struct Linear {  // [N, K] weight of this rank; int8 + per-row scale, or fp16
    DevBuf w, scale;
    int N = 0, K = 0;
};

struct Layer {
    DevBuf attn_norm, ffn_norm;
    Linear qkv, o, gate_up, down;
    DevBuf gate_stage, up_stage;  // fp16 halves of gate_up awaiting their partner
    bool have_gate = false, have_up = false;
};

}  // namespace
}  // namespace b2llm

using namespace b2llm;

struct b2llm_engine {
    b2llm_model_desc d;
    int rank = 0, tp = 1;
    void* comm = nullptr;
    cudaStream_t stream = nullptr;
    // local (per rank) sizes
    int D = 0, nq = 0, nkv = 0, inter = 0, nqkv = 0;
    b2llm_kv_geom geom{};
    std::vector<Layer> layers;
    DevBuf embedding, final_norm, lm_head;
    DevBuf rope_cos, rope_sin;
    // activations
    DevBuf x, a8, a_s, qkv, attn, act, b8, b_s, tmp, y16, xl, yl, logits, attn_ws, w16_scratch;
    // tp > 1: the lm head is vocab-parallel -- this rank holds rows [rank * head_rows, (rank + 1) * head_rows) of
    // output.weight, computes its column block of the logits and the blocks are all-gathered (SURVEY 8(e);
    // the reference reads the gathered logits on rank 0, llm_engine.cc:200)
    bool head_split = false;
    int head_rows = 0;
    DevBuf logits_part, logits_gather;
    // tp > 1: fused residual join over NVLink peer memory (tp_join.cu).  tmp / x / a8 / a_s / y16 are windows of `comm`,
    // which every peer maps; B2LLM_TP_JOIN=nccl keeps ncclAllReduce + separate norm kernels (the cross-check path)
    bool tp_fused = false;
    bool comm_mapped = false;
    DevBuf cbuf;                       // the peer-mapped communication buffer
    std::vector<void*> comm_retired;   // outgrown buffers: peers may still map them, freed with the engine
    std::vector<void*> ipc_opened;
    TpLayout comm_layout{};
    TpPeers peers{};
    uint32_t epoch = 0;
    // staged inputs
    DevBuf in_tokens, in_seq_starts, in_kv_starts, in_start_pos, in_cache_idx;
    b2llm_step staged{};
    bool page_table_staged = false;  // cache_mode 1: a page table uploaded by set_inputs is still on the device
    void* kv_cache = nullptr;
    void* kv_scale = nullptr;
    int64_t last_launches = 0;
    int64_t last_tokens = 0;
    int64_t last_batch = 0;
    int attn_impl = 0, gemm_impl = 0;
    int split_k = 1;  // B2LLM_CONF_DECODING_ATTN_SPLIT_K
    int64_t cap_tokens = 0, cap_batch = 0;  // current capacity of the activation buffers (b2llm_engine_reserve)
    // optional per-kernel-class timing with CUDA events on the engine stream (bench.py roofline)
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    std::vector<std::pair<int, std::pair<size_t, size_t>>> ev_spans;  // (class, (begin, end))
    cudaEvent_t next_event() {
        if (ev_used == ev_pool.size()) {
            cudaEvent_t ev;
            cudaEventCreate(&ev);
            ev_pool.push_back(ev);
        }
        return ev_pool[ev_used++];
    }
};

namespace {
struct Span {  // RAII: records begin / end events around a kernel class when profiling is on
    b2llm_engine* e;
    int cls;
    size_t b = 0;
    Span(b2llm_engine* e_, int cls_) : e(e_), cls(cls_) {
        if (!e->profiling) return;
        b = e->ev_used;
        cudaEventRecord(e->next_event(), e->stream);
    }
    ~Span() {
        if (!e->profiling) return;
        const size_t en = e->ev_used;
        cudaEventRecord(e->next_event(), e->stream);
        e->ev_spans.push_back({cls, {b, en}});
    }
};
}  // namespace

namespace {

int32_t alloc_linear(Linear& L, int N, int K, int quant_method) {
    L.N = N;
    L.K = K;
    if (quant_method == B2LLM_QUANT_W4A16) {  // packed nibbles + one fp16 scale per 128 K-elements
        int32_t rc = L.w.ensure((size_t)N * K / 2);
        if (rc) return rc;
        return L.scale.ensure((size_t)N * (K / 128) * sizeof(__half));
    }
    const bool i8 = quant_method == B2LLM_QUANT_ONLINE_I8I8;
    int32_t rc = L.w.ensure((size_t)N * K * (i8 ? 1 : 2));
    if (rc) return rc;
    if (i8) rc = L.scale.ensure((size_t)N * sizeof(float));
    return rc;
}

// fp16 [N, K] on device -> the Linear (quantise or copy)
int32_t finalize_linear(b2llm_engine* e, Linear& L, const __half* w16) {
    if (e->d.quant_method == B2LLM_QUANT_ONLINE_I8I8)
        return launch_quant_weight(e->stream, w16, L.N, L.K, L.w.as<int8_t>(), L.scale.as<float>());
    if (e->d.quant_method == B2LLM_QUANT_W4A16)
        return launch_quant_weight_w4(e->stream, w16, L.N, L.K, L.w.as<uint8_t>(), L.scale.as<__half>());
    B2_CHECK_CUDA(cudaMemcpyAsync(L.w.p, w16, (size_t)L.N * L.K * 2, cudaMemcpyDeviceToDevice, e->stream));
    return B2LLM_OK;
}

int32_t gemm(b2llm_engine* e, const void* a, const float* a_scale, const Linear& L, int64_t M, int epi, void* out,
             int64_t ldc) {
    const bool i8 = e->d.quant_method == B2LLM_QUANT_ONLINE_I8I8;
    Span span(e, 1);
    if (e->d.quant_method == B2LLM_QUANT_W4A16) {
        // fused kernel: packed nibbles -> smem -> converter warps -> tcgen05 (0.5 B of HBM traffic per weight)
        if (e->gemm_impl == 0 || e->gemm_impl == 2) {
            const int32_t rcf = launch_gemm_w4a16(e->stream, a, L.w.as<uint8_t>(), L.scale.p, M, L.N, L.K, epi, out, ldc);
            if (rcf != B2LLM_ERR_UNSUPPORTED) return rcf;
        }
        // fallback / cross-check (B2LLM_GEMM_IMPL=1 or 3): expand the int4 weight to its fp16 operand in a scratch buffer,
        // then the fp16 GEMM (4.5 B of traffic per weight)
        __half* w16 = e->w16_scratch.as<__half>();
        int32_t rc = launch_dequant_w4(e->stream, L.w.as<uint8_t>(), L.scale.as<__half>(), L.N, L.K, w16);
        if (rc) return rc;
        if (e->gemm_impl != 1 && gemm_tc_available()) {
            rc = launch_gemm_tc(e->stream, false, a, nullptr, w16, nullptr, M, L.N, L.K, epi, out, ldc);
            if (rc != B2LLM_ERR_UNSUPPORTED) return rc;
        }
        return launch_gemm_mma(e->stream, false, a, nullptr, w16, nullptr, M, L.N, L.K, epi, out, ldc);
    }
    if (e->gemm_impl != 1 && gemm_tc_available()) {
        const int32_t rc = launch_gemm_tc(e->stream, i8, a, a_scale, L.w.p, L.scale.as<float>(), M, L.N, L.K, epi, out, ldc);
        if (rc != B2LLM_ERR_UNSUPPORTED) return rc;
    }
    return launch_gemm_mma(e->stream, i8, a, a_scale, L.w.p, L.scale.as<float>(), M, L.N, L.K, epi, out, ldc);
}

int32_t allreduce_half(b2llm_engine* e, __half* buf, size_t count) {
    load_nccl();
    B2_REQUIRE(g_nccl_allreduce != nullptr, B2LLM_ERR_UNSUPPORTED, "tensor parallel: libnccl.so.2 not loadable");
    B2_REQUIRE(e->comm != nullptr, B2LLM_ERR_INVALID_VALUE, "tensor parallel: no NCCL communicator given");
    Span span(e, 3);
    const int r = g_nccl_allreduce(buf, buf, count, kNcclFloat16, kNcclSum, e->comm, e->stream);
    B2_REQUIRE(r == 0, B2LLM_ERR_DEVICE, "ncclAllReduce failed with code " + std::to_string(r));
    return B2LLM_OK;
}

// (re)publish this rank's communication buffer and map the peers' (collective: every rank of the group calls it in the
// same step -- buffer sizes follow the step sizes, which are identical on all ranks, llm_engine.cc:179)
int32_t map_comm(b2llm_engine* e) {
    load_nccl();
    B2_REQUIRE(g_nccl_allgather != nullptr && e->comm != nullptr, B2LLM_ERR_UNSUPPORTED,
               "tensor parallel: ncclAllGather / communicator unavailable");
    for (void* m : e->ipc_opened) cudaIpcCloseMemHandle(m);
    e->ipc_opened.clear();
    B2_CHECK_CUDA(cudaMemsetAsync(e->cbuf.p, 0, 1024, e->stream));  // flags, CTA counter, fault word
    const int32_t rc = tp_comm_exchange(e->stream, e->comm, g_nccl_allgather, e->rank, e->tp, e->cbuf.p, &e->peers, &e->ipc_opened);
    // every rank must take the same path: agree on the outcome (a rank without peer access / IPC would otherwise run
    // ncclAllReduce against peers spinning in the fused kernel).  MIN over the ranks of "mapping succeeded".
    int ok = rc == B2LLM_OK ? 1 : 0;
    if (const char* t = getenv("B2LLM_TP_TEST_FAIL_MAP")) {  // test hook: pretend THIS rank could not map its peers
        if (atoi(t) == e->rank) ok = 0;
    }
    int* flag = reinterpret_cast<int*>((uint8_t*)e->cbuf.p + 768);
    B2_REQUIRE(g_nccl_allreduce != nullptr, B2LLM_ERR_UNSUPPORTED, "tensor parallel: ncclAllReduce unavailable");
    B2_CHECK_CUDA(cudaMemcpyAsync(flag, &ok, sizeof(int), cudaMemcpyHostToDevice, e->stream));
    const int nr = g_nccl_allreduce(flag, flag, 1, 2 /* ncclInt32 */, 3 /* ncclMin */, e->comm, e->stream);
    B2_REQUIRE(nr == 0, B2LLM_ERR_DEVICE, "ncclAllReduce failed with code " + std::to_string(nr));
    B2_CHECK_CUDA(cudaMemcpyAsync(&ok, flag, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    B2_CHECK_CUDA(cudaStreamSynchronize(e->stream));
    if (!ok) {
        // some rank could not map its peers: the whole group falls back to ncclAllReduce + separate kernels for good
        // (the activation buffers stay where they are: windows of this rank's own communication buffer)
        fprintf(stderr, "b2llm: rank %d: peer mapping of the tensor-parallel buffers failed on some rank (%s) -- using ncclAllReduce\n",
                e->rank, rc ? g_last_error.c_str() : "another rank");
        e->tp_fused = false;
        return B2LLM_OK;
    }
    e->epoch = 0;
    e->comm_mapped = true;
    return B2LLM_OK;
}

int32_t tp_join(b2llm_engine* e, int mode, bool bcast_x, const __half* gamma, int64_t T) {
    Span span(e, 3);
    return launch_tp_join(e->stream, e->peers, e->comm_layout, mode, bcast_x, gamma, e->d.norm_eps, T, e->d.hidden_dim, ++e->epoch);
}

}  // namespace

extern "C" const char* b2llm_version(void) { return "b2llm 0.2 (sm_100a)"; }
extern "C" const char* b2llm_last_error(void) { return g_last_error.c_str(); }

extern "C" int32_t b2llm_rope_table(int32_t max_position, int32_t head_dim, float theta, float* cos_host,
                                    float* sin_host) {
    B2_REQUIRE(max_position > 0 && head_dim > 0 && head_dim % 2 == 0 && cos_host && sin_host, B2LLM_ERR_INVALID_VALUE,
               "rope_table: bad arguments");
    const int half = head_dim / 2;
    for (int i = 0; i < half; ++i) {
        const double inv_freq = pow((double)theta, -2.0 * (double)i / (double)head_dim);
        for (int p = 0; p < max_position; ++p) {
            const double ang = (double)p * inv_freq;
            cos_host[(size_t)p * half + i] = (float)cos(ang);
            sin_host[(size_t)p * half + i] = (float)sin(ang);
        }
    }
    return B2LLM_OK;
}

extern "C" int32_t b2llm_engine_create(const b2llm_model_desc* desc, int32_t rank, int32_t tp, void* nccl_comm,
                                       void* stream, b2llm_engine** out) {
    B2_REQUIRE(desc && out, B2LLM_ERR_INVALID_VALUE, "engine_create: null argument");
    *out = nullptr;
    const b2llm_model_desc& d = *desc;
    B2_REQUIRE(tp >= 1 && rank >= 0 && rank < tp, B2LLM_ERR_INVALID_VALUE, "engine_create: bad rank / tp");
    B2_REQUIRE(d.num_heads > 0 && d.hidden_dim % d.num_heads == 0, B2LLM_ERR_INVALID_VALUE, "bad num_heads");
    B2_REQUIRE(d.num_kv_heads > 0 && d.num_heads % d.num_kv_heads == 0, B2LLM_ERR_INVALID_VALUE, "bad num_kv_heads");
    B2_REQUIRE(d.num_heads % tp == 0 && d.num_kv_heads % tp == 0 && d.intermediate_dim % tp == 0, B2LLM_ERR_INVALID_VALUE,
               "heads / kv heads / intermediate_dim must divide by tensor_parallel_size");
    // the two cache modes the reference accepts (llm_generator.cc:131-136): int8 with an fp16 scale per 8 elements, or
    // plain fp16 without a scale tensor (bit 0, group 1)
    B2_REQUIRE((d.cache_quant_bit == 8 && d.cache_quant_group == 8) || (d.cache_quant_bit == 0 && d.cache_quant_group == 1),
               B2LLM_ERR_UNSUPPORTED, "KV cache: (cache_quant_bit 8, cache_quant_group 8) or (cache_quant_bit 0, cache_quant_group 1)");
    B2_REQUIRE(d.cache_layout >= 0 && d.cache_layout <= 3, B2LLM_ERR_INVALID_VALUE, "cache_layout must be 0..3");
    B2_REQUIRE(d.cache_mode == 0 || (d.cache_mode == 1 && d.page_size > 0), B2LLM_ERR_INVALID_VALUE,
               "cache_mode must be 0, or 1 with page_size > 0");
    B2_REQUIRE(d.quant_method == B2LLM_QUANT_NONE || d.quant_method == B2LLM_QUANT_ONLINE_I8I8 ||
                   d.quant_method == B2LLM_QUANT_W4A16,
               B2LLM_ERR_UNSUPPORTED, "quant_method must be none, online_i8i8 or w4a16");
    B2_REQUIRE(d.quant_method != B2LLM_QUANT_W4A16 ||
                   (d.hidden_dim % 128 == 0 && (d.intermediate_dim / tp) % 128 == 0 && (d.hidden_dim / tp) % 128 == 0),
               B2LLM_ERR_UNSUPPORTED, "w4a16: every (per-rank) reduction dimension must be a multiple of the group size 128");
    B2_REQUIRE(d.max_tokens_per_step > 0 && d.max_running_batch > 0 && d.max_position > 0, B2LLM_ERR_INVALID_VALUE,
               "max_tokens_per_step / max_running_batch / max_position must be positive");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_last_error("no CUDA device: b2llm has no CPU fallback");
        return B2LLM_ERR_DEVICE;
    }
    int dev = 0;
    B2_CHECK_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    B2_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
    B2_REQUIRE(prop.major == 10, B2LLM_ERR_DEVICE, "b2llm kernels are built for sm_100a only");

    auto* e = new b2llm_engine();
    e->d = d;
    e->rank = rank;
    e->tp = tp;
    e->comm = nccl_comm;
    e->stream = (cudaStream_t)stream;
    e->D = d.hidden_dim / d.num_heads;
    e->nq = d.num_heads / tp;
    e->nkv = d.num_kv_heads / tp;
    e->inter = d.intermediate_dim / tp;
    e->nqkv = (e->nq + 2 * e->nkv) * e->D;
    e->geom.num_layers = d.num_layers;
    e->geom.num_kv_heads = e->nkv;
    e->geom.head_dim = e->D;
    e->geom.quant_group = d.cache_quant_group;
    e->geom.cache_layout = d.cache_layout;
    e->geom.cache_mode = d.cache_mode;
    e->geom.page_size = d.page_size;
    e->head_rows = d.vocab_size;
    if (tp > 1 && d.vocab_size % tp == 0 && (d.vocab_size / tp) % 32 == 0) {
        const char* hs = getenv("B2LLM_TP_HEAD");  // "whole": every rank keeps (and multiplies by) the whole lm head
        if (!(hs && hs[0] == 'w')) {
            e->head_split = true;
            e->head_rows = d.vocab_size / tp;
        }
    }
    if ((tp == 2 || tp == 4 || tp == 8) && d.hidden_dim % 8 == 0 && d.hidden_dim <= 8192) {
        const char* js = getenv("B2LLM_TP_JOIN");
        e->tp_fused = !(js && js[0] == 'n');
    }
    if (const char* s = getenv("B2LLM_ATTN_IMPL")) e->attn_impl = atoi(s);
    if (const char* s = getenv("B2LLM_GEMM_IMPL")) e->gemm_impl = atoi(s);

    const bool i8 = d.quant_method == B2LLM_QUANT_ONLINE_I8I8;
    const int h = d.hidden_dim;
    int32_t rc = B2LLM_OK;
    auto chk = [&](int32_t r) { if (rc == B2LLM_OK) rc = r; };
    e->layers.resize(d.num_layers);
    for (auto& L : e->layers) {
        chk(L.attn_norm.ensure((size_t)h * 2));
        chk(L.ffn_norm.ensure((size_t)h * 2));
        chk(alloc_linear(L.qkv, e->nqkv, h, d.quant_method));
        chk(alloc_linear(L.o, h, e->nq * e->D, d.quant_method));
        chk(alloc_linear(L.gate_up, 2 * e->inter, h, d.quant_method));
        chk(alloc_linear(L.down, h, e->inter, d.quant_method));
    }
    if (d.quant_method == B2LLM_QUANT_W4A16) {
        const size_t big = std::max((size_t)std::max(e->nqkv, 2 * e->inter) * h, (size_t)h * std::max(e->inter, e->nq * e->D));
        chk(e->w16_scratch.ensure(big * 2));
    }
    chk(e->embedding.ensure((size_t)d.vocab_size * h * 2));
    chk(e->final_norm.ensure((size_t)h * 2));
    chk(e->lm_head.ensure((size_t)e->head_rows * h * 2));
    const size_t half = e->D / 2;
    chk(e->rope_cos.ensure((size_t)d.max_position * half * 4));
    chk(e->rope_sin.ensure((size_t)d.max_position * half * 4));
    chk(b2llm_engine_reserve(e, d.max_tokens_per_step, d.max_running_batch));
    if (rc == B2LLM_OK) {
        std::vector<float> c((size_t)d.max_position * half), s((size_t)d.max_position * half);
        b2llm_rope_table(d.max_position, e->D, d.rope_theta, c.data(), s.data());
        if (cudaMemcpy(e->rope_cos.p, c.data(), c.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(e->rope_sin.p, s.data(), s.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
            set_last_error("rope table upload failed");
            rc = B2LLM_ERR_DEVICE_MEMORY;
        }
    }
    if (rc != B2LLM_OK) {
        b2llm_engine_destroy(e);
        return rc;
    }
    *out = e;
    return B2LLM_OK;
}

// (re)size the activation / staging buffers for steps of up to `max_tokens` tokens and `max_batch` sequences.
// The reference gives these limits to the generator, not to the runtime (ppl.nn sizes its buffers per step),
// so the ppl::nn::Runtime adapter calls this lazily; growth synchronises the stream first.
extern "C" int32_t b2llm_engine_reserve(b2llm_engine* e, int64_t max_tokens, int64_t max_batch) {
    B2_REQUIRE(e && max_tokens > 0 && max_batch > 0, B2LLM_ERR_INVALID_VALUE, "reserve: bad arguments");
    if (max_tokens <= e->cap_tokens && max_batch <= e->cap_batch) return B2LLM_OK;
    if (e->cap_tokens > 0) B2_CHECK_CUDA(cudaStreamSynchronize(e->stream));
    const b2llm_model_desc& d = e->d;
    const bool i8 = d.quant_method == B2LLM_QUANT_ONLINE_I8I8;
    const int h = d.hidden_dim;
    // growth is geometric (a ramp of prefill sizes must not reallocate every step) and drops the contents: inputs staged
    // by set_inputs are gone, so set_inputs has to follow (b2llm_engine_run refuses to run on stale staging)
    auto grow = [](int64_t want, int64_t cap) { return (size_t)(want <= cap ? cap : std::max<int64_t>(want, cap + cap / 2)); };
    const size_t T = grow(max_tokens, e->cap_tokens), B = grow(max_batch, e->cap_batch);
    e->staged = b2llm_step{};
    e->page_table_staged = false;
    int32_t rc = B2LLM_OK;
    auto chk = [&](int32_t r) { if (rc == B2LLM_OK) rc = r; };
    const size_t amax_cols = (size_t)(h > e->nq * e->D ? h : e->nq * e->D);
    if (e->tp_fused) {
        // one allocation every peer maps; the old one (if any) is retired, not freed: peers may still hold its mapping
        const TpLayout L = tp_layout((int64_t)T, h, (int)amax_cols);
        if (e->cbuf.p) e->comm_retired.push_back(e->cbuf.p);
        e->cbuf.p = nullptr;
        e->cbuf.bytes = 0;
        chk(e->cbuf.ensure(L.total));
        if (rc == B2LLM_OK) {
            uint8_t* base = (uint8_t*)e->cbuf.p;
            e->comm_layout = L;
            e->comm_mapped = false;
            e->tmp.view(base + L.partial, T * h * 2);
            e->x.view(base + L.x, T * h * 2);
            e->a8.view(base + L.q, T * amax_cols);
            e->a_s.view(base + L.qscale, T * 4);
            e->y16.view(base + L.y, T * h * 2);
        }
    } else {
        chk(e->x.ensure(T * h * 2));
        chk(e->a8.ensure(T * amax_cols));
        chk(e->a_s.ensure(T * 4));
    }
    chk(e->qkv.ensure(T * e->nqkv * 2));
    chk(e->attn.ensure(T * e->nq * e->D * 2));
    chk(e->act.ensure(T * e->inter * 2));
    chk(e->b8.ensure(T * e->inter));
    chk(e->b_s.ensure(T * 4));
    if (e->tp > 1 && !e->tp_fused) chk(e->tmp.ensure(T * h * 2));
    if ((!i8 || e->tp > 1) && !e->tp_fused) chk(e->y16.ensure(T * h * 2));
    chk(e->xl.ensure(B * h * 2));
    chk(e->yl.ensure(B * h * 2));
    chk(e->logits.ensure(B * (size_t)d.vocab_size * 4));
    if (e->head_split) {
        chk(e->logits_part.ensure(B * (size_t)e->head_rows * 4));
        chk(e->logits_gather.ensure(B * (size_t)d.vocab_size * 4));
    }
    chk(e->attn_ws.ensure((size_t)attention_workspace_bytes(B, e->nq, e->D)));
    chk(e->in_tokens.ensure(T * 8));
    chk(e->in_seq_starts.ensure((B + 1) * 8));
    chk(e->in_kv_starts.ensure((B + 1) * 8));
    chk(e->in_start_pos.ensure(B * 8));
    chk(e->in_cache_idx.ensure(B * 8));
    if (rc == B2LLM_OK) {
        e->cap_tokens = (int64_t)T;
        e->cap_batch = (int64_t)B;
    }
    return rc;
}

extern "C" int32_t b2llm_engine_configure(b2llm_engine* e, int32_t key, int64_t value) {
    B2_REQUIRE(e, B2LLM_ERR_INVALID_VALUE, "configure: null engine");
    switch (key) {
        case B2LLM_CONF_DECODING_ATTN_SPLIT_K:
            B2_REQUIRE(value >= 0 && value <= 2, B2LLM_ERR_INVALID_VALUE, "configure: split-k must be 0, 1 or 2");
            e->split_k = (int)value;
            return B2LLM_OK;
        case B2LLM_CONF_ATTN_IMPL:
            B2_REQUIRE(value >= 0 && value <= 2, B2LLM_ERR_INVALID_VALUE, "configure: attention impl must be 0, 1 or 2");
            e->attn_impl = (int)value;
            return B2LLM_OK;
        case B2LLM_CONF_GEMM_IMPL:
            B2_REQUIRE(value >= 0 && value <= 3, B2LLM_ERR_INVALID_VALUE, "configure: gemm impl must be 0..3");
            e->gemm_impl = (int)value;
            return B2LLM_OK;
        default:
            set_last_error("configure: unknown key " + std::to_string(key));
            return B2LLM_ERR_UNSUPPORTED;
    }
}

extern "C" int32_t b2llm_engine_destroy(b2llm_engine* e) {
    if (!e) return B2LLM_OK;
    if (e->stream) cudaStreamSynchronize(e->stream); else cudaDeviceSynchronize();
    for (auto& L : e->layers) {
        for (DevBuf* b : {&L.attn_norm, &L.ffn_norm, &L.qkv.w, &L.qkv.scale, &L.o.w, &L.o.scale, &L.gate_up.w,
                          &L.gate_up.scale, &L.down.w, &L.down.scale, &L.gate_stage, &L.up_stage})
            b->release();
    }
    for (cudaEvent_t ev : e->ev_pool) cudaEventDestroy(ev);
    for (void* m : e->ipc_opened) cudaIpcCloseMemHandle(m);
    for (void* m : e->comm_retired) cudaFree(m);
    e->cbuf.release();
    for (DevBuf* b : {&e->embedding, &e->final_norm, &e->lm_head, &e->rope_cos, &e->rope_sin, &e->x, &e->a8, &e->a_s,
                      &e->qkv, &e->attn, &e->act, &e->b8, &e->b_s, &e->tmp, &e->y16, &e->xl, &e->yl, &e->logits,
                      &e->attn_ws, &e->w16_scratch, &e->logits_part, &e->logits_gather, &e->in_tokens, &e->in_seq_starts, &e->in_kv_starts, &e->in_start_pos,
                      &e->in_cache_idx})
        b->release();
    delete e;
    return B2LLM_OK;
}

extern "C" int32_t b2llm_engine_kv_bytes_per_token(const b2llm_engine* e, uint64_t* cache_bytes, uint64_t* scale_bytes) {
    B2_REQUIRE(e && cache_bytes && scale_bytes, B2LLM_ERR_INVALID_VALUE, "null argument");
    // resource_manager.cc:381-388: sizeof(int8 | fp16) per cached element; scale bytes only when cache_quant_bit > 0
    const bool kv16 = e->d.cache_quant_bit == 0;
    *cache_bytes = (uint64_t)e->d.num_layers * 2 * e->nkv * e->D * (kv16 ? 2 : 1);
    *scale_bytes = kv16 ? 0 : (uint64_t)e->d.num_layers * 2 * e->nkv * e->D / e->d.cache_quant_group * 2;
    return B2LLM_OK;
}

extern "C" int32_t b2llm_engine_bind_kv(b2llm_engine* e, void* kv_cache_device, void* kv_scale_device,
                                        uint64_t kv_cache_max_tokens) {
    B2_REQUIRE(e && kv_cache_device && kv_cache_max_tokens > 0, B2LLM_ERR_INVALID_VALUE, "bind_kv: null / empty KV memory");
    B2_REQUIRE(kv_scale_device || e->d.cache_quant_bit == 0, B2LLM_ERR_INVALID_VALUE,
               "bind_kv: the int8 cache needs its scale memory (runtime input 10, llm_engine.h:134-136)");
    e->kv_cache = kv_cache_device;
    e->kv_scale = kv_scale_device;
    e->geom.max_tokens = kv_cache_max_tokens;
    return B2LLM_OK;
}

// ------------------------------------------------------------------------------------ weights
namespace {

// copy a [rows, cols] window of a host fp16 matrix with `full_cols` columns to a dense device buffer
int32_t upload_window(b2llm_engine* e, const __half* host, int64_t full_cols, int64_t row0, int64_t rows, int64_t col0,
                      int64_t cols, __half* dst) {
    B2_CHECK_CUDA(cudaMemcpy2DAsync(dst, (size_t)cols * 2, host + row0 * full_cols + col0, (size_t)full_cols * 2,
                                    (size_t)cols * 2, (size_t)rows, cudaMemcpyHostToDevice, e->stream));
    return B2LLM_OK;
}

int32_t finish_gate_up(b2llm_engine* e, Layer& L) {
    if (!(L.have_gate && L.have_up)) return B2LLM_OK;
    DevBuf inter16;
    int32_t rc = inter16.ensure((size_t)2 * e->inter * e->d.hidden_dim * 2);
    if (rc) return rc;
    rc = launch_interleave_rows(e->stream, L.gate_stage.as<__half>(), L.up_stage.as<__half>(), e->inter, e->d.hidden_dim,
                                inter16.as<__half>());
    if (rc == B2LLM_OK) rc = finalize_linear(e, L.gate_up, inter16.as<__half>());
    cudaStreamSynchronize(e->stream);
    inter16.release();
    L.gate_stage.release();
    L.up_stage.release();
    L.have_gate = L.have_up = false;
    return rc;
}

}  // namespace

namespace {
// `sharded`: the host tensor already is this rank's tensor-parallel shard (what a ppl.pmx export stores per
// model_slice_<rank>); otherwise it is the full tensor and the rank's window is cut out here.
int32_t load_weight_impl(b2llm_engine* e, int32_t kind, int32_t layer, const void* host_fp16, uint64_t num_elements,
                         bool sharded) {
    B2_REQUIRE(e && host_fp16, B2LLM_ERR_INVALID_VALUE, "load_weight: null argument");
    const b2llm_model_desc& d = e->d;
    const __half* src = (const __half*)host_fp16;
    const int64_t h = d.hidden_dim, D = e->D;
    const int r = sharded ? 0 : e->rank;
    const int64_t NQ = sharded ? e->nq : d.num_heads, NKV = sharded ? e->nkv : d.num_kv_heads;
    const int64_t I = sharded ? e->inter : d.intermediate_dim;
    auto expect = [&](uint64_t n) -> bool {
        if (num_elements != n) {
            set_last_error("load_weight: expected " + std::to_string(n) + " elements, got " + std::to_string(num_elements));
            return false;
        }
        return true;
    };
    if (kind == B2LLM_W_EMBEDDING) {
        if (!expect((uint64_t)d.vocab_size * h)) return B2LLM_ERR_INVALID_VALUE;
        B2_CHECK_CUDA(cudaMemcpyAsync(e->embedding.p, src, num_elements * 2, cudaMemcpyHostToDevice, e->stream));
        B2_CHECK_CUDA(cudaStreamSynchronize(e->stream));
        return B2LLM_OK;
    }
    if (kind == B2LLM_W_LM_HEAD) {
        // vocab-parallel head (tp > 1): the whole [vocab, h] tensor (this rank's row window is cut out) or, from
        // load_weight_shard, the rank's own [vocab / tp, h] slice exactly as model_slice_<rank> stores it
        const uint64_t whole = (uint64_t)d.vocab_size * h, part = (uint64_t)e->head_rows * h;
        if (sharded && e->head_split && num_elements == part) {
            B2_CHECK_CUDA(cudaMemcpyAsync(e->lm_head.p, src, part * 2, cudaMemcpyHostToDevice, e->stream));
        } else {
            if (!expect(whole)) return B2LLM_ERR_INVALID_VALUE;
            const __half* win = src + (e->head_split ? (uint64_t)e->rank * part : 0);
            B2_CHECK_CUDA(cudaMemcpyAsync(e->lm_head.p, win, part * 2, cudaMemcpyHostToDevice, e->stream));
        }
        B2_CHECK_CUDA(cudaStreamSynchronize(e->stream));
        return B2LLM_OK;
    }
    if (kind == B2LLM_W_FINAL_NORM) {
        if (!expect((uint64_t)h)) return B2LLM_ERR_INVALID_VALUE;
        B2_CHECK_CUDA(cudaMemcpyAsync(e->final_norm.p, src, h * 2, cudaMemcpyHostToDevice, e->stream));
        B2_CHECK_CUDA(cudaStreamSynchronize(e->stream));
        return B2LLM_OK;
    }
    B2_REQUIRE(layer >= 0 && layer < d.num_layers, B2LLM_ERR_INVALID_VALUE, "load_weight: bad layer");
    Layer& L = e->layers[layer];
    int32_t rc = B2LLM_OK;
    DevBuf stage;
    switch (kind) {
        case B2LLM_W_ATTN_NORM:
        case B2LLM_W_FFN_NORM: {
            if (!expect((uint64_t)h)) return B2LLM_ERR_INVALID_VALUE;
            DevBuf& dst = kind == B2LLM_W_ATTN_NORM ? L.attn_norm : L.ffn_norm;
            B2_CHECK_CUDA(cudaMemcpyAsync(dst.p, src, h * 2, cudaMemcpyHostToDevice, e->stream));
            break;
        }
        case B2LLM_W_QKV: {
            if (!expect((uint64_t)((NQ + 2 * NKV) * D * h))) return B2LLM_ERR_INVALID_VALUE;
            if ((rc = stage.ensure((size_t)e->nqkv * h * 2))) return rc;
            __half* s16 = stage.as<__half>();
            rc = upload_window(e, src, h, (int64_t)r * e->nq * D, e->nq * D, 0, h, s16);
            if (!rc) rc = upload_window(e, src, h, NQ * D + (int64_t)r * e->nkv * D, e->nkv * D, 0, h, s16 + (int64_t)e->nq * D * h);
            if (!rc) rc = upload_window(e, src, h, (NQ + NKV) * D + (int64_t)r * e->nkv * D, e->nkv * D, 0, h,
                                        s16 + (int64_t)(e->nq + e->nkv) * D * h);
            if (!rc) rc = finalize_linear(e, L.qkv, s16);
            break;
        }
        case B2LLM_W_O: {
            const int64_t full = NQ * D;
            if (!expect((uint64_t)(h * full))) return B2LLM_ERR_INVALID_VALUE;
            if ((rc = stage.ensure((size_t)h * e->nq * D * 2))) return rc;
            rc = upload_window(e, src, full, 0, h, (int64_t)r * e->nq * D, e->nq * D, stage.as<__half>());
            if (!rc) rc = finalize_linear(e, L.o, stage.as<__half>());
            break;
        }
        case B2LLM_W_GATE:
        case B2LLM_W_UP: {
            if (!expect((uint64_t)I * h)) return B2LLM_ERR_INVALID_VALUE;
            DevBuf& dst = kind == B2LLM_W_GATE ? L.gate_stage : L.up_stage;
            if ((rc = dst.ensure((size_t)e->inter * h * 2))) return rc;
            rc = upload_window(e, src, h, (int64_t)r * e->inter, e->inter, 0, h, dst.as<__half>());
            (kind == B2LLM_W_GATE ? L.have_gate : L.have_up) = true;
            if (!rc) rc = finish_gate_up(e, L);
            break;
        }
        case B2LLM_W_DOWN: {
            const int64_t full = I;
            if (!expect((uint64_t)(h * full))) return B2LLM_ERR_INVALID_VALUE;
            if ((rc = stage.ensure((size_t)h * e->inter * 2))) return rc;
            rc = upload_window(e, src, full, 0, h, (int64_t)r * e->inter, e->inter, stage.as<__half>());
            if (!rc) rc = finalize_linear(e, L.down, stage.as<__half>());
            break;
        }
        default:
            set_last_error("load_weight: unknown kind");
            return B2LLM_ERR_INVALID_VALUE;
    }
    cudaError_t err = cudaStreamSynchronize(e->stream);
    stage.release();
    if (rc == B2LLM_OK && err != cudaSuccess) {
        set_last_error(std::string("load_weight: ") + cudaGetErrorString(err));
        rc = B2LLM_ERR_DEVICE;
    }
    return rc;
}
}  // namespace

extern "C" int32_t b2llm_engine_load_weight(b2llm_engine* e, int32_t kind, int32_t layer, const void* host_fp16,
                                            uint64_t num_elements) {
    return load_weight_impl(e, kind, layer, host_fp16, num_elements, false);
}

extern "C" int32_t b2llm_engine_load_weight_shard(b2llm_engine* e, int32_t kind, int32_t layer, const void* host_fp16,
                                                  uint64_t num_elements) {
    return load_weight_impl(e, kind, layer, host_fp16, num_elements, true);
}

extern "C" int32_t b2llm_engine_random_init(b2llm_engine* e, uint64_t seed) {
    B2_REQUIRE(e, B2LLM_ERR_INVALID_VALUE, "random_init: null engine");
    const b2llm_model_desc& d = e->d;
    const int64_t h = d.hidden_dim, D = e->D;
    const int r = e->rank;
    cudaStream_t s = e->stream;
    // tensor ids and std: oracle/weights.py
    const float STD_EMBED = 1.0f, STD_W = 0.02f, STD_NORM = 0.02f;
    int32_t rc = launch_synth_fp16(s, seed, 1, (uint64_t)d.vocab_size * h, STD_EMBED, 0.f, e->embedding.as<__half>());
    if (!rc) rc = launch_synth_fp16(s, seed, 2, (uint64_t)h, STD_NORM, 1.f, e->final_norm.as<__half>());
    if (!rc) rc = launch_synth_fp16_2d(s, seed, 3, e->head_rows, h, e->head_split ? (int64_t)e->rank * e->head_rows : 0, 0, h, STD_W,
                                       0.f, e->lm_head.as<__half>());
    if (rc) return rc;
    DevBuf stage, stage2, stage3;
    const size_t big = (size_t)(2 * e->inter > e->nqkv ? 2 * e->inter : e->nqkv) * h;
    if ((rc = stage.ensure(big * 2))) return rc;
    if ((rc = stage2.ensure((size_t)e->inter * h * 2))) return rc;
    if ((rc = stage3.ensure((size_t)e->inter * h * 2))) return rc;
    __half* s16 = stage.as<__half>();
    for (int l = 0; l < d.num_layers && !rc; ++l) {
        Layer& L = e->layers[l];
        const uint64_t t0 = 16 + (uint64_t)l * 8;
        rc = launch_synth_fp16(s, seed, t0 + 0, (uint64_t)h, STD_NORM, 1.f, L.attn_norm.as<__half>());
        if (!rc) rc = launch_synth_fp16(s, seed, t0 + 3, (uint64_t)h, STD_NORM, 1.f, L.ffn_norm.as<__half>());
        // qkv: three row windows of the full [(NQ + 2 NKV) D, h]
        const int64_t NQ = d.num_heads, NKV = d.num_kv_heads;
        if (!rc) rc = launch_synth_fp16_2d(s, seed, t0 + 1, e->nq * D, h, (int64_t)r * e->nq * D, 0, h, STD_W, 0.f, s16);
        if (!rc) rc = launch_synth_fp16_2d(s, seed, t0 + 1, e->nkv * D, h, NQ * D + (int64_t)r * e->nkv * D, 0, h, STD_W, 0.f,
                                           s16 + (int64_t)e->nq * D * h);
        if (!rc) rc = launch_synth_fp16_2d(s, seed, t0 + 1, e->nkv * D, h, (NQ + NKV) * D + (int64_t)r * e->nkv * D, 0, h,
                                           STD_W, 0.f, s16 + (int64_t)(e->nq + e->nkv) * D * h);
        if (!rc) rc = finalize_linear(e, L.qkv, s16);
        // o: column window of [h, NQ D]
        if (!rc) rc = launch_synth_fp16_2d(s, seed, t0 + 2, h, e->nq * D, 0, (int64_t)r * e->nq * D, NQ * D, STD_W, 0.f, s16);
        if (!rc) rc = finalize_linear(e, L.o, s16);
        // gate / up: row windows of [I, h], interleaved
        if (!rc) rc = launch_synth_fp16_2d(s, seed, t0 + 4, e->inter, h, (int64_t)r * e->inter, 0, h, STD_W, 0.f, stage2.as<__half>());
        if (!rc) rc = launch_synth_fp16_2d(s, seed, t0 + 5, e->inter, h, (int64_t)r * e->inter, 0, h, STD_W, 0.f, stage3.as<__half>());
        if (!rc) rc = launch_interleave_rows(s, stage2.as<__half>(), stage3.as<__half>(), e->inter, (int)h, s16);
        if (!rc) rc = finalize_linear(e, L.gate_up, s16);
        // down: column window of [h, I]
        if (!rc) rc = launch_synth_fp16_2d(s, seed, t0 + 6, h, e->inter, 0, (int64_t)r * e->inter, d.intermediate_dim, STD_W, 0.f, s16);
        if (!rc) rc = finalize_linear(e, L.down, s16);
    }
    cudaError_t err = cudaStreamSynchronize(s);
    stage.release();
    stage2.release();
    stage3.release();
    if (rc == B2LLM_OK && err != cudaSuccess) {
        set_last_error(std::string("random_init: ") + cudaGetErrorString(err));
        rc = B2LLM_ERR_DEVICE;
    }
    return rc;
}

// ------------------------------------------------------------------------------------ the step
extern "C" int32_t b2llm_engine_set_inputs(b2llm_engine* e, const int64_t* token_ids, int64_t num_tokens,
                                           const int64_t* seq_starts, const int64_t* kv_starts, const int64_t* start_pos,
                                           int64_t batch, const int64_t* cache_indices_or_page_list, int64_t max_pages,
                                           int64_t decoding_batches, int64_t max_seq_len, int64_t max_kv_len,
                                           int32_t req_list_changed) {
    B2_REQUIRE(e && token_ids && seq_starts && kv_starts && start_pos, B2LLM_ERR_INVALID_VALUE, "set_inputs: null pointer");
    B2_REQUIRE(num_tokens >= 0 && num_tokens <= e->cap_tokens, B2LLM_ERR_INVALID_VALUE,
               "set_inputs: num_tokens exceeds max_tokens_per_step");
    B2_REQUIRE(batch >= 0 && batch <= e->cap_batch, B2LLM_ERR_INVALID_VALUE,
               "set_inputs: batch exceeds max_running_batch");
    cudaStream_t s = e->stream;
    B2_CHECK_CUDA(cudaMemcpyAsync(e->in_tokens.p, token_ids, num_tokens * 8, cudaMemcpyHostToDevice, s));
    B2_CHECK_CUDA(cudaMemcpyAsync(e->in_seq_starts.p, seq_starts, (batch + 1) * 8, cudaMemcpyHostToDevice, s));
    B2_CHECK_CUDA(cudaMemcpyAsync(e->in_kv_starts.p, kv_starts, (batch + 1) * 8, cudaMemcpyHostToDevice, s));
    B2_CHECK_CUDA(cudaMemcpyAsync(e->in_start_pos.p, start_pos, batch * 8, cudaMemcpyHostToDevice, s));
    if (e->d.cache_mode == 0) {
        B2_REQUIRE(cache_indices_or_page_list, B2LLM_ERR_INVALID_VALUE, "set_inputs: cache_indices required");
        B2_CHECK_CUDA(cudaMemcpyAsync(e->in_cache_idx.p, cache_indices_or_page_list, batch * 8, cudaMemcpyHostToDevice, s));
    } else if (req_list_changed) {
        B2_REQUIRE(cache_indices_or_page_list && max_pages > 0, B2LLM_ERR_INVALID_VALUE, "set_inputs: page_list required");
        const size_t bytes = (size_t)batch * max_pages * 8;
        if (bytes > e->in_cache_idx.bytes) {
            B2_CHECK_CUDA(cudaStreamSynchronize(s));  // the old table may still be in use
            const int32_t rc = e->in_cache_idx.ensure(bytes + bytes / 2);
            if (rc) return rc;
        }
        B2_CHECK_CUDA(cudaMemcpyAsync(e->in_cache_idx.p, cache_indices_or_page_list, bytes, cudaMemcpyHostToDevice, s));
        e->staged.max_pages = max_pages;
        e->page_table_staged = true;
    } else {
        B2_REQUIRE(e->page_table_staged, B2LLM_ERR_INVALID_VALUE,
                   "set_inputs: req_list_changed == 0 but no page table is staged (first step, or b2llm_engine_reserve grew "
                   "the buffers since)");
    }
    e->staged.token_ids = e->in_tokens.as<int64_t>();
    e->staged.seq_starts = e->in_seq_starts.as<int64_t>();
    e->staged.kv_starts = e->in_kv_starts.as<int64_t>();
    e->staged.start_pos = e->in_start_pos.as<int64_t>();
    e->staged.cache_indices = e->in_cache_idx.as<int64_t>();
    e->staged.num_tokens = num_tokens;
    e->staged.batch = batch;
    e->staged.decoding_batches = decoding_batches;
    e->staged.max_seq_len = max_seq_len;
    e->staged.max_kv_len = max_kv_len;
    return B2LLM_OK;
}

extern "C" int32_t b2llm_engine_staged_inputs(b2llm_engine* e, const int64_t** token_ids, const int64_t** seq_starts,
                                              const int64_t** start_pos) {
    B2_REQUIRE(e, B2LLM_ERR_INVALID_VALUE, "null engine");
    if (token_ids) *token_ids = e->in_tokens.as<int64_t>();
    if (seq_starts) *seq_starts = e->in_seq_starts.as<int64_t>();
    if (start_pos) *start_pos = e->in_start_pos.as<int64_t>();
    return B2LLM_OK;
}

extern "C" int32_t b2llm_engine_run(b2llm_engine* e, int32_t cache_prefill, float** logits_device, int64_t* logits_stride) {
    B2_REQUIRE(e, B2LLM_ERR_INVALID_VALUE, "null engine");
    B2_REQUIRE(e->staged.token_ids != nullptr, B2LLM_ERR_INVALID_VALUE,
               "run: no staged inputs (b2llm_engine_set_inputs must follow b2llm_engine_reserve growth)");
    e->staged.cache_prefill = cache_prefill;
    return b2llm_engine_forward(e, &e->staged, logits_device, logits_stride);
}

extern "C" int32_t b2llm_engine_forward(b2llm_engine* e, const b2llm_step* st, float** logits_device,
                                        int64_t* logits_stride) {
    B2_REQUIRE(e && st, B2LLM_ERR_INVALID_VALUE, "forward: null argument");
    B2_REQUIRE(e->kv_cache && (e->kv_scale || e->d.cache_quant_bit == 0), B2LLM_ERR_INVALID_VALUE,
               "forward: KV memory not bound (b2llm_engine_bind_kv)");
    const b2llm_model_desc& d = e->d;
    const int64_t T = st->num_tokens, B = st->batch;
    B2_REQUIRE(T >= 0 && T <= e->cap_tokens && B >= 0 && B <= e->cap_batch, B2LLM_ERR_INVALID_VALUE,
               "forward: step exceeds max_tokens_per_step / max_running_batch");
    B2_REQUIRE(st->decoding_batches >= 0 && st->decoding_batches <= B, B2LLM_ERR_INVALID_VALUE, "forward: bad decoding_batches");
    B2_REQUIRE(st->max_kv_len <= d.max_position, B2LLM_ERR_INVALID_VALUE, "forward: max_kv_len exceeds max_position");
    if (logits_device) *logits_device = e->logits.as<float>();
    if (logits_stride) *logits_stride = d.vocab_size;
    if (T == 0 || B == 0) return B2LLM_OK;

    const int64_t launches0 = g_launch_count;
    cudaStream_t s = e->stream;
    const bool i8 = d.quant_method == B2LLM_QUANT_ONLINE_I8I8;
    const int h = d.hidden_dim;
    __half* x = e->x.as<__half>();
    int32_t rc = launch_embedding(s, st->token_ids, e->embedding.as<__half>(), T, h, d.vocab_size, x);
    if (rc) return rc;

    AttnArgs aa{};
    aa.qkv = e->qkv.as<__half>();
    aa.step = st;
    aa.num_heads = e->nq;
    aa.geom = e->geom;
    aa.kv_cache = (const int8_t*)e->kv_cache;
    aa.kv_scale = (const __half*)e->kv_scale;
    aa.workspace = e->attn_ws.p;
    aa.out = e->attn.as<__half>();
    aa.split_k = e->split_k;
    const int64_t decode_tokens = st->decoding_batches;  // one token per decoding sequence, placed first
    const bool tp = e->tp > 1;
    const __half* pending_skip = nullptr;  // tp > 1, NCCL path: all-reduced projection output not yet added to x
    bool fused = tp && e->tp_fused;
    if (fused && !e->comm_mapped) {
        if ((rc = map_comm(e))) return rc;
        fused = e->tp_fused;  // false if the group agreed to fall back to NCCL
    }
    const int join_mode = i8 ? 1 : 2;      // what the fused join leaves for the next block: int8 + scale, or fp16
    bool have_norm = false;                 // the fused join already wrote norm(x) of the upcoming block into a8 / y16

    for (int l = 0; l < d.num_layers; ++l) {
        Layer& L = e->layers[l];
        // ---- attention block
        const void* lin_in = i8 ? e->a8.p : e->y16.p;
        if (!have_norm) {
            if (i8)
                rc = launch_rmsnorm_quant(s, x, pending_skip, L.attn_norm.as<__half>(), d.norm_eps, T, h, e->a8.as<int8_t>(),
                                          e->a_s.as<float>(), nullptr);
            else
                rc = launch_rmsnorm_quant(s, x, pending_skip, L.attn_norm.as<__half>(), d.norm_eps, T, h, nullptr, nullptr,
                                          e->y16.as<__half>());
            if (rc) return rc;
        }
        pending_skip = nullptr;
        have_norm = false;
        if ((rc = gemm(e, lin_in, e->a_s.as<float>(), L.qkv, T, EPI_F16, e->qkv.p, e->nqkv))) return rc;
        if ((rc = launch_rope_kv_append(s, e->qkv.as<__half>(), st, e->nq, e->geom, l, e->rope_cos.as<float>(),
                                        e->rope_sin.as<float>(), (int8_t*)e->kv_cache, (__half*)e->kv_scale)))
            return rc;
        aa.layer = l;
        {
            Span span(e, 0);
            if (e->attn_impl == 1 || e->D != 128) {
                if ((rc = launch_attention_simple(s, aa, 0, T))) return rc;
            } else {
                if ((rc = launch_attention_decode_mma(s, aa))) return rc;
                if (decode_tokens < T && (rc = launch_attention_prefill(s, aa))) return rc;
            }
        }
        if (i8) {
            if ((rc = launch_quant_rows(s, e->attn.as<__half>(), T, e->nq * e->D, e->a8.as<int8_t>(), e->a_s.as<float>()))) return rc;
            lin_in = e->a8.p;
        } else {
            lin_in = e->attn.p;
        }
        if (!tp) {
            if ((rc = gemm(e, lin_in, e->a_s.as<float>(), L.o, T, EPI_RESIDUAL, x, h))) return rc;
        } else {
            if ((rc = gemm(e, lin_in, e->a_s.as<float>(), L.o, T, EPI_F16, e->tmp.p, h))) return rc;
            if (fused) {  // all-reduce + residual + ffn RMSNorm + quant, one kernel over peer memory
                if ((rc = tp_join(e, join_mode, false, L.ffn_norm.as<__half>(), T))) return rc;
                have_norm = true;
            } else {
                if ((rc = allreduce_half(e, e->tmp.as<__half>(), (size_t)T * h))) return rc;
                pending_skip = e->tmp.as<__half>();
            }
        }
        // ---- feed-forward block
        lin_in = i8 ? e->a8.p : e->y16.p;
        if (!have_norm) {
            if (i8)
                rc = launch_rmsnorm_quant(s, x, pending_skip, L.ffn_norm.as<__half>(), d.norm_eps, T, h, e->a8.as<int8_t>(),
                                          e->a_s.as<float>(), nullptr);
            else
                rc = launch_rmsnorm_quant(s, x, pending_skip, L.ffn_norm.as<__half>(), d.norm_eps, T, h, nullptr, nullptr,
                                          e->y16.as<__half>());
            if (rc) return rc;
        }
        pending_skip = nullptr;
        have_norm = false;
        if ((rc = gemm(e, lin_in, e->a_s.as<float>(), L.gate_up, T, EPI_SWIGLU, e->act.p, e->inter))) return rc;
        if (i8) {
            if ((rc = launch_quant_rows(s, e->act.as<__half>(), T, e->inter, e->b8.as<int8_t>(), e->b_s.as<float>()))) return rc;
            lin_in = e->b8.p;
        } else {
            lin_in = e->act.p;
        }
        if (!tp) {
            if ((rc = gemm(e, lin_in, e->b_s.as<float>(), L.down, T, EPI_RESIDUAL, x, h))) return rc;
        } else {
            if ((rc = gemm(e, lin_in, e->b_s.as<float>(), L.down, T, EPI_F16, e->tmp.p, h))) return rc;
            if (fused) {
                // ... + the NEXT layer's attention RMSNorm + quant; the last join of the step only completes the residual
                // stream and gives every rank all of its rows (the lm head needs the last token of every sequence)
                const bool last = l + 1 == d.num_layers;
                if ((rc = tp_join(e, last ? 0 : join_mode, last, last ? nullptr : e->layers[l + 1].attn_norm.as<__half>(), T)))
                    return rc;
                have_norm = !last;
            } else {
                if ((rc = allreduce_half(e, e->tmp.as<__half>(), (size_t)T * h))) return rc;
                pending_skip = e->tmp.as<__half>();
            }
        }
    }
    if (pending_skip) {  // fold the last all-reduced output into x (norm output unused)
        if ((rc = launch_rmsnorm_quant(s, x, pending_skip, e->final_norm.as<__half>(), d.norm_eps, T, h, nullptr, nullptr,
                                       e->y16.as<__half>())))
            return rc;
    }
    // ---- head: only the sampler's rank needs logits (llm_engine.cc:200), every rank computes them so
    // that any rank can be asked; rank != 0 callers may ignore the result
    if ((rc = launch_gather_rows(s, x, st->seq_starts, B, h, e->xl.as<__half>()))) return rc;
    if ((rc = launch_rmsnorm_quant(s, e->xl.as<__half>(), nullptr, e->final_norm.as<__half>(), d.norm_eps, B, h, nullptr,
                                   nullptr, e->yl.as<__half>())))
        return rc;
    {
        // vocab-parallel: this rank's column block [B, head_rows] -> all-gather [tp, B, head_rows] -> [B, vocab]
        float* gemm_out = e->head_split ? e->logits_part.as<float>() : e->logits.as<float>();
        const int64_t gemm_ld = e->head_rows;
        int32_t r2 = B2LLM_ERR_UNSUPPORTED;
        {
            Span span(e, 2);
            if (e->gemm_impl != 1 && gemm_tc_available())
                r2 = launch_gemm_tc(s, false, e->yl.p, nullptr, e->lm_head.p, nullptr, B, e->head_rows, h, EPI_F32, gemm_out, gemm_ld);
            if (r2 == B2LLM_ERR_UNSUPPORTED)
                r2 = launch_gemm_mma(s, false, e->yl.p, nullptr, e->lm_head.p, nullptr, B, e->head_rows, h, EPI_F32, gemm_out, gemm_ld);
        }
        if (r2) return r2;
        if (e->head_split) {
            Span span(e, 3);
            load_nccl();
            B2_REQUIRE(g_nccl_allgather != nullptr && e->comm != nullptr, B2LLM_ERR_UNSUPPORTED,
                       "tensor parallel: ncclAllGather / communicator unavailable");
            const int r = g_nccl_allgather(e->logits_part.p, e->logits_gather.p, (size_t)B * e->head_rows, kNcclFloat32, e->comm, s);
            B2_REQUIRE(r == 0, B2LLM_ERR_DEVICE, "ncclAllGather failed with code " + std::to_string(r));
            if ((rc = launch_interleave_blocks(s, e->logits_gather.as<float>(), e->tp, B, e->head_rows, e->logits.as<float>())))
                return rc;
        }
    }
    e->last_launches = g_launch_count - launches0;
    e->last_tokens = T;
    e->last_batch = B;
    return B2LLM_OK;
}

extern "C" int32_t b2llm_engine_profile(b2llm_engine* e, int32_t enable) {
    B2_REQUIRE(e, B2LLM_ERR_INVALID_VALUE, "null engine");
    e->profiling = enable != 0;
    e->ev_used = 0;
    e->ev_spans.clear();
    return B2LLM_OK;
}

extern "C" int32_t b2llm_engine_profile_read(b2llm_engine* e, double* ms_by_class, int64_t* count_by_class,
                                             int32_t num_classes) {
    B2_REQUIRE(e && ms_by_class && count_by_class && num_classes > 0, B2LLM_ERR_INVALID_VALUE, "profile_read: bad arguments");
    B2_CHECK_CUDA(cudaStreamSynchronize(e->stream));
    for (int i = 0; i < num_classes; ++i) { ms_by_class[i] = 0.0; count_by_class[i] = 0; }
    for (auto& sp : e->ev_spans) {
        if (sp.first >= num_classes) continue;
        float ms = 0.f;
        B2_CHECK_CUDA(cudaEventElapsedTime(&ms, e->ev_pool[sp.second.first], e->ev_pool[sp.second.second]));
        ms_by_class[sp.first] += ms;
        count_by_class[sp.first] += 1;
    }
    e->ev_used = 0;
    e->ev_spans.clear();
    return B2LLM_OK;
}

extern "C" int64_t b2llm_engine_last_launch_count(const b2llm_engine* e) { return e ? e->last_launches : 0; }

extern "C" int32_t b2llm_engine_tp_join_stats(b2llm_engine* e, double* out8) {
    B2_REQUIRE(e && out8, B2LLM_ERR_INVALID_VALUE, "tp_join_stats: null argument");
    for (int i = 0; i < 8; ++i) out8[i] = 0.0;
    if (!e->tp_fused || !e->cbuf.p) return B2LLM_OK;
    unsigned long long h[8];
    B2_CHECK_CUDA(cudaStreamSynchronize(e->stream));
    B2_CHECK_CUDA(cudaMemcpy(h, (uint8_t*)e->cbuf.p + e->comm_layout.flags + 512, sizeof(h), cudaMemcpyDeviceToHost));
    B2_CHECK_CUDA(cudaMemset((uint8_t*)e->cbuf.p + e->comm_layout.flags + 512, 0, sizeof(h)));
    for (int i = 0; i < 8; ++i) out8[i] = (double)h[i];
    return B2LLM_OK;
}

extern "C" int32_t b2llm_engine_debug_read(b2llm_engine* e, int32_t what, void* host_dst, uint64_t bytes) {
    B2_REQUIRE(e && host_dst, B2LLM_ERR_INVALID_VALUE, "debug_read: null argument");
    const void* src = nullptr;
    uint64_t avail = 0;
    switch (what) {
        case 0: src = e->x.p; avail = (uint64_t)e->last_tokens * e->d.hidden_dim * 2; break;
        case 1: src = e->qkv.p; avail = (uint64_t)e->last_tokens * e->nqkv * 2; break;
        case 2: src = e->attn.p; avail = (uint64_t)e->last_tokens * e->nq * e->D * 2; break;
        case 3: src = e->logits.p; avail = (uint64_t)e->last_batch * e->d.vocab_size * 4; break;
        default: set_last_error("debug_read: unknown item"); return B2LLM_ERR_INVALID_VALUE;
    }
    B2_REQUIRE(bytes <= avail, B2LLM_ERR_INVALID_VALUE, "debug_read: too many bytes requested");
    B2_CHECK_CUDA(cudaStreamSynchronize(e->stream));
    B2_CHECK_CUDA(cudaMemcpy(host_dst, src, bytes, cudaMemcpyDeviceToHost));
    return B2LLM_OK;
}

// ------------------------------------------------------------------------------------ op-level ABI
extern "C" int32_t b2llm_op_rmsnorm_quant(void* stream, void* x_fp16, const void* skip_fp16, const void* gamma_fp16,
                                          float eps, int64_t rows, int32_t hidden, int8_t* q_out, float* scale_out,
                                          void* y_out_fp16) {
    B2_REQUIRE(x_fp16 && gamma_fp16, B2LLM_ERR_INVALID_VALUE, "rmsnorm_quant: null pointer");
    return launch_rmsnorm_quant((cudaStream_t)stream, (__half*)x_fp16, (const __half*)skip_fp16, (const __half*)gamma_fp16,
                                eps, rows, hidden, q_out, scale_out, (__half*)y_out_fp16);
}

extern "C" int32_t b2llm_op_quant_rows(void* stream, const void* x_fp16, int64_t rows, int32_t cols, int8_t* q_out,
                                       float* scale_out) {
    B2_REQUIRE(x_fp16 && q_out && scale_out, B2LLM_ERR_INVALID_VALUE, "quant_rows: null pointer");
    return launch_quant_rows((cudaStream_t)stream, (const __half*)x_fp16, rows, cols, q_out, scale_out);
}

extern "C" int32_t b2llm_op_gemm_w8a8(void* stream, const int8_t* a, const float* a_scale, const int8_t* w,
                                      const float* w_scale, int64_t M, int32_t N, int32_t K, int32_t epilogue,
                                      void* out_fp16, int32_t impl) {
    B2_REQUIRE(a && a_scale && w && w_scale && out_fp16, B2LLM_ERR_INVALID_VALUE, "gemm_w8a8: null pointer");
    B2_REQUIRE(epilogue >= 0 && epilogue <= 2, B2LLM_ERR_INVALID_VALUE, "gemm_w8a8: epilogue must be 0, 1 or 2");
    const int64_t ldc = epilogue == EPI_SWIGLU ? N / 2 : N;
    if (impl >= 2 || (impl == 0 && gemm_tc_available())) {
        const int32_t rc = launch_gemm_tc((cudaStream_t)stream, true, a, a_scale, w, w_scale, M, N, K, epilogue, out_fp16, ldc,
                                          impl == 2 ? 0 : (impl == 3 ? 2 : -1));
        if (rc != B2LLM_ERR_UNSUPPORTED || impl >= 2) return rc;
    }
    return launch_gemm_mma((cudaStream_t)stream, true, a, a_scale, w, w_scale, M, N, K, epilogue, out_fp16, ldc);
}

extern "C" int32_t b2llm_op_gemm_f16(void* stream, const void* a_fp16, const void* w_fp16, int64_t M, int32_t N,
                                     int32_t K, int32_t epilogue, void* out, int64_t ldc, int32_t impl) {
    B2_REQUIRE(a_fp16 && w_fp16 && out, B2LLM_ERR_INVALID_VALUE, "gemm_f16: null pointer");
    B2_REQUIRE(epilogue >= 0 && epilogue <= 3, B2LLM_ERR_INVALID_VALUE, "gemm_f16: epilogue must be 0..3");
    if (ldc <= 0) ldc = epilogue == EPI_SWIGLU ? N / 2 : N;
    if (impl >= 2 || (impl == 0 && gemm_tc_available())) {
        const int32_t rc = launch_gemm_tc((cudaStream_t)stream, false, a_fp16, nullptr, w_fp16, nullptr, M, N, K, epilogue, out, ldc,
                                          impl == 2 ? 0 : (impl == 3 ? 2 : -1));
        if (rc != B2LLM_ERR_UNSUPPORTED || impl >= 2) return rc;
    }
    return launch_gemm_mma((cudaStream_t)stream, false, a_fp16, nullptr, w_fp16, nullptr, M, N, K, epilogue, out, ldc);
}

extern "C" int32_t b2llm_op_rope_kv_append(void* stream, void* qkv_fp16, const b2llm_step* step, int32_t num_heads,
                                           const b2llm_kv_geom* geom, int32_t layer, const float* rope_cos,
                                           const float* rope_sin, void* kv_cache, void* kv_scale) {
    B2_REQUIRE(qkv_fp16 && step && geom && rope_cos && rope_sin && kv_cache && (kv_scale || geom->quant_group == 1),
               B2LLM_ERR_INVALID_VALUE, "rope_kv_append: null pointer");
    return launch_rope_kv_append((cudaStream_t)stream, (__half*)qkv_fp16, step, num_heads, *geom, layer, rope_cos, rope_sin,
                                 (int8_t*)kv_cache, (__half*)kv_scale);
}

extern "C" int64_t b2llm_attention_workspace_size(int64_t batch, int32_t num_heads, int32_t head_dim) {
    return attention_workspace_bytes(batch, num_heads, head_dim);
}

extern "C" int32_t b2llm_attention_decode_plan(int64_t decoding_batches, int32_t num_heads, int32_t num_kv_heads, int64_t max_kv_len,
                                               int32_t* nsplit, int32_t* warps) {
    B2_REQUIRE(nsplit && warps && decoding_batches > 0 && num_kv_heads > 0 && num_heads % num_kv_heads == 0 && max_kv_len > 0,
               B2LLM_ERR_INVALID_VALUE, "attention_decode_plan: bad arguments");
    const int gq = num_heads / num_kv_heads;
    const int G = gq == 1 ? 1 : (gq <= 4 ? 4 : 8);
    const int chunks = (gq + G - 1) / G;
    int n = 1, w = 1;
    attention_decode_plan((int64_t)num_kv_heads * chunks * decoding_batches, decoding_batches, max_kv_len, &n, &w);
    *nsplit = n;
    *warps = w;
    return B2LLM_OK;
}

extern "C" int32_t b2llm_op_attention(void* stream, const void* qkv_fp16, const b2llm_step* step, int32_t num_heads,
                                      const b2llm_kv_geom* geom, int32_t layer, const void* kv_cache, const void* kv_scale,
                                      void* workspace, void* out_fp16, int32_t impl) {
    B2_REQUIRE(qkv_fp16 && step && geom && kv_cache && (kv_scale || geom->quant_group == 1) && out_fp16, B2LLM_ERR_INVALID_VALUE,
               "attention: null pointer");
    AttnArgs aa{};
    aa.qkv = (const __half*)qkv_fp16;
    aa.step = step;
    aa.num_heads = num_heads;
    aa.geom = *geom;
    aa.layer = layer;
    aa.kv_cache = (const int8_t*)kv_cache;
    aa.kv_scale = (const __half*)kv_scale;
    aa.workspace = workspace;
    aa.out = (__half*)out_fp16;
    cudaStream_t s = (cudaStream_t)stream;
    if (impl >= 3 && impl <= 5) aa.loader = impl == 3 ? 2 : (impl == 4 ? 0 : 1);
    const int prefill_which = impl == 6 ? 1 : (impl == 8 ? 0 : -1);  // tcgen05 / mma.sync prefill kernel forced
    if (impl == 1 || (impl == 0 && geom->head_dim != 128)) return launch_attention_simple(s, aa, 0, step->num_tokens);
    B2_REQUIRE(workspace, B2LLM_ERR_INVALID_VALUE, "attention: workspace required");
    int32_t rc = launch_attention_decode_mma(s, aa);
    if (rc) return rc;
    if (step->decoding_batches >= step->batch) return B2LLM_OK;
    return launch_attention_prefill(s, aa, prefill_which);
}

extern "C" int32_t b2llm_op_synth_fp16(void* stream, uint64_t seed, uint64_t tensor_id, uint64_t num_elements, float std,
                                       float mean, void* out_fp16) {
    B2_REQUIRE(out_fp16, B2LLM_ERR_INVALID_VALUE, "synth_fp16: null pointer");
    return launch_synth_fp16((cudaStream_t)stream, seed, tensor_id, num_elements, std, mean, (__half*)out_fp16);
}

extern "C" int32_t b2llm_op_quant_weight(void* stream, const void* w_fp16, int32_t N, int32_t K, int8_t* q_out,
                                         float* scale_out) {
    B2_REQUIRE(w_fp16 && q_out && scale_out, B2LLM_ERR_INVALID_VALUE, "quant_weight: null pointer");
    return launch_quant_weight((cudaStream_t)stream, (const __half*)w_fp16, N, K, q_out, scale_out);
}

extern "C" int32_t b2llm_op_quant_weight_w4(void* stream, const void* w_fp16, int32_t N, int32_t K, uint8_t* packed_out,
                                            void* scale_out_fp16) {
    B2_REQUIRE(w_fp16 && packed_out && scale_out_fp16, B2LLM_ERR_INVALID_VALUE, "quant_weight_w4: null pointer");
    return launch_quant_weight_w4((cudaStream_t)stream, (const __half*)w_fp16, N, K, packed_out, (__half*)scale_out_fp16);
}

extern "C" int32_t b2llm_op_dequant_w4(void* stream, const uint8_t* packed, const void* scale_fp16, int32_t N, int32_t K,
                                       void* w_out_fp16) {
    B2_REQUIRE(packed && scale_fp16 && w_out_fp16, B2LLM_ERR_INVALID_VALUE, "dequant_w4: null pointer");
    return launch_dequant_w4((cudaStream_t)stream, packed, (const __half*)scale_fp16, N, K, (__half*)w_out_fp16);
}

extern "C" int32_t b2llm_op_gemm_w4a16(void* stream, const void* a_fp16, const uint8_t* packed, const void* scale_fp16, int64_t M,
                                       int32_t N, int32_t K, int32_t epilogue, void* out_fp16) {
    B2_REQUIRE(a_fp16 && packed && scale_fp16 && out_fp16, B2LLM_ERR_INVALID_VALUE, "gemm_w4a16: null pointer");
    B2_REQUIRE(epilogue >= 0 && epilogue <= 2, B2LLM_ERR_INVALID_VALUE, "gemm_w4a16: epilogue must be 0, 1 or 2");
    const int64_t ldc = epilogue == EPI_SWIGLU ? N / 2 : N;
    return launch_gemm_w4a16((cudaStream_t)stream, a_fp16, packed, scale_fp16, M, N, K, epilogue, out_fp16, ldc);
}
