/*
 * b2llm.h -- C ABI of the B200-native (sm_100a) implementation of ppl.llm.serving's batched
 * LLaMA decode hot path.  Plain pointers and sizes only; no C++ / torch types.
 *
 * This is the boundary a ppl.llm.serving maintainer binds instead of ppl.nn's llm_cuda engine
 * and ppl.llm.kernel.cuda's pmx operators.  Each entry point cites the reference interface it
 * replaces (paths relative to the ppl.llm.serving tree).  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every function returns int32_t: 0 (B2LLM_OK == ppl::common::RC_SUCCESS) or a RetCode-valued
 *     error (src/engine/llm_engine.cc:171-236 error convention); nothing throws;
 *   - all device work is enqueued on the cudaStream_t given at engine creation (or passed to an
 *     op-level call) and is asynchronous unless stated otherwise
 *     (reference: one stream per rank, src/backends/cuda/resource_manager.cc:224-232);
 *   - "device" pointers are CUDA device pointers, "host" pointers ordinary host memory;
 *   - the library never falls back to a CPU path: without a usable sm_100 device every call that
 *     needs one fails with B2LLM_ERR_DEVICE.
 */
#ifndef B2LLM_H_
#define B2LLM_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define B2LLM_API
#else
#define B2LLM_API __attribute__((visibility("default")))
#endif

/* values match ppl::common::RetCode as used by the reference (src/engine/llm_engine.cc,
 * src/backends/cuda/post_processor.cc) */
enum {
    B2LLM_OK = 0,
    B2LLM_ERR_OTHER = 1,          /* RC_OTHER_ERROR          */
    B2LLM_ERR_INVALID_VALUE = 2,  /* RC_INVALID_VALUE        */
    B2LLM_ERR_OUT_OF_MEMORY = 3,  /* RC_OUT_OF_MEMORY        */
    B2LLM_ERR_UNSUPPORTED = 4,    /* RC_UNSUPPORTED          */
    B2LLM_ERR_DEVICE = 5,         /* RC_DEVICE_RUNTIME_ERROR */
    B2LLM_ERR_DEVICE_MEMORY = 6   /* RC_DEVICE_MEMORY_ERROR  */
};

/* quant_method: src/backends/cuda/resource_manager.cc:49-56 accepts "none" and "online_i8i8" */
enum {
    B2LLM_QUANT_NONE = 0,
    B2LLM_QUANT_ONLINE_I8I8 = 1,
    /* W4A16: NOT selectable in the reference at this commit (anything but the two above is rejected,
     * resource_manager.cc:49-56), so its semantics are builder-defined: symmetric int4 weights, group 128
     * along K, fp16 scales, fp16 activations, operand value fp16(q * scale).  Reachable through the C ABI only. */
    B2LLM_QUANT_W4A16 = 2
};

/* Mirror of ppl::llm::ModelConfig (src/common/config.h:64-84, parsed by src/common/config.cc:31-148)
 * plus what the reference keeps in the exported graph rather than params.json (norm eps, rope
 * theta: config.h:74 "norm_eps // not used") and the scheduler limits the engine sizes its
 * activation buffers from (src/common/config.h:45-60). */
typedef struct b2llm_model_desc {
    int32_t hidden_dim;
    int32_t intermediate_dim;
    int32_t num_layers;
    int32_t num_heads;
    int32_t num_kv_heads;
    int32_t vocab_size;
    float norm_eps;            /* 1e-5 for LLaMA-2 */
    float rope_theta;          /* 10000 */
    int32_t cache_quant_bit;   /* 8: int8 KV + fp16 scale per group (the only quantised mode), or 0: fp16 KV, no scale tensor
                                  (llm_generator.cc:131-136, resource_manager.cc:381-388) */
    int32_t cache_quant_group; /* 8 with cache_quant_bit 8; 1 with cache_quant_bit 0 */
    int32_t cache_layout;      /* 0..3, llm_engine.cc:118-169 */
    int32_t cache_mode;        /* 0 contiguous index, 1 page table */
    int32_t page_size;         /* tokens per page when cache_mode == 1 */
    int32_t quant_method;      /* B2LLM_QUANT_* */
    int32_t max_position;      /* rope table length (>= max total tokens per request) */
    int32_t max_tokens_per_step; /* activation buffer rows (GeneratorConfig::max_tokens_per_step) */
    int32_t max_running_batch;   /* logits rows (GeneratorConfig::max_running_batch) */
    int32_t reserved[3];
} b2llm_model_desc;

/* One forward step over a ragged batch: the device-resident form of ppl::llm::ModelInput
 * (src/engine/llm_engine.h:40-60) as SetInputTask uploads it (src/engine/llm_engine.cc:29-111),
 * i.e. runtime inputs 0,2,3,4,6 on the device and 5,7,8 as host scalars (llm_engine.h:124-147). */
typedef struct b2llm_step {
    const int64_t* token_ids;     /* device [num_tokens]           input 0 */
    const int64_t* seq_starts;    /* device [batch + 1]            input 2 */
    const int64_t* kv_starts;     /* device [batch + 1]            input 3 */
    const int64_t* cache_indices; /* device [batch] (cache_mode 0) or [batch, max_pages] page
                                     begin-token indices, INT64_MAX padded (cache_mode 1)  input 4 */
    const int64_t* start_pos;     /* device [batch]                input 6 */
    int64_t num_tokens;
    int64_t batch;
    int64_t decoding_batches;     /* host scalar, input 5: sequences [0, decoding_batches) decode */
    int64_t max_seq_len;          /* host scalar, input 7 */
    int64_t max_kv_len;           /* host scalar, input 8 */
    int64_t max_pages;            /* row length of the page table */
    int32_t cache_prefill;        /* ENGINE_CONF_CACHE_PREFILL (llm_engine.cc:114): a prefill sequence
                                     with start_pos > 0 attends to its cached prefix */
    int32_t reserved;
} b2llm_step;

typedef struct b2llm_engine b2llm_engine;

/* weight kinds for b2llm_engine_load_weight */
enum {
    B2LLM_W_EMBEDDING = 0,  /* fp16 [vocab, hidden]                              (layer ignored) */
    B2LLM_W_FINAL_NORM = 1, /* fp16 [hidden]                                                      */
    B2LLM_W_LM_HEAD = 2,    /* fp16 [vocab, hidden]                                               */
    B2LLM_W_ATTN_NORM = 3,  /* fp16 [hidden]                                                      */
    B2LLM_W_QKV = 4,        /* fp16 [(nq + 2 nkv) * head_dim, hidden]  rows: q heads, k heads, v heads */
    B2LLM_W_O = 5,          /* fp16 [hidden, nq * head_dim]                                       */
    B2LLM_W_FFN_NORM = 6,   /* fp16 [hidden]                                                      */
    B2LLM_W_GATE = 7,       /* fp16 [intermediate, hidden]                                        */
    B2LLM_W_UP = 8,         /* fp16 [intermediate, hidden]                                        */
    B2LLM_W_DOWN = 9        /* fp16 [hidden, intermediate]                                        */
};

B2LLM_API const char* b2llm_version(void);
B2LLM_API const char* b2llm_last_error(void); /* thread-local text of the last failure */

/* ---- engine life cycle ------------------------------------------------------------------
 * replaces: llm_cuda EngineFactory::Create + RuntimeBuilder::{LoadModel,Preprocess,CreateRuntime}
 * (src/backends/cuda/resource_manager.cc:43-177, 213-371).  `rank`/`tp` select the tensor-parallel
 * slice (heads / tp, kv heads / tp, intermediate / tp; resource_manager.cc:280-286,
 * llm_engine.cc:124); `nccl_comm` is the rank's ncclComm_t (ENGINE_CONF_SET_TP_NCCL_COMM,
 * resource_manager.cc:238-244) or NULL when tp == 1; `stream` is the rank's cudaStream_t. */
B2LLM_API int32_t b2llm_engine_create(const b2llm_model_desc* desc, int32_t rank, int32_t tp, void* nccl_comm,
                                      void* stream, b2llm_engine** out);
B2LLM_API int32_t b2llm_engine_destroy(b2llm_engine* e);

/* (re)size activation / staging buffers for steps of up to max_tokens tokens and max_batch sequences.  The
 * reference gives these limits to the generator (GeneratorConfig, src/common/config.h:45-60), not to the
 * runtime -- ppl.nn sizes buffers per step -- so the ppl::nn::Runtime adapter (host/src/pplnn_b200.cc) calls
 * this lazily from Run().  Never shrinks; growth synchronises the engine stream. */
B2LLM_API int32_t b2llm_engine_reserve(b2llm_engine* e, int64_t max_tokens, int64_t max_batch);

/* replaces ppl::nn::Engine::Configure for the keys the reference sets (resource_manager.cc:74-112):
 * key = B2LLM_CONF_*, value as documented per key.  Unknown keys -> B2LLM_ERR_UNSUPPORTED. */
enum {
    B2LLM_CONF_DECODING_ATTN_SPLIT_K = 3, /* 0 never split the KV range, 1 heuristic (default), 2 always */
    B2LLM_CONF_ATTN_IMPL = 100,           /* 0 auto, 1 simple reference kernel, 2 tensor-core split-KV kernel */
    B2LLM_CONF_GEMM_IMPL = 101            /* 0 auto, 1 mma.sync baseline, 2 tcgen05; W4A16: 0/2 fused kernel, 1/3 dequant +
                                             fp16 GEMM cross-check path */
};
B2LLM_API int32_t b2llm_engine_configure(b2llm_engine* e, int32_t key, int64_t value);

/* weights.  load_weight takes the FULL (unsharded) fp16 tensor on the host and keeps this rank's
 * slice; with quant_method == online_i8i8 the projection weights are quantised per output channel on
 * load (the reference's "online" quantisation pass, resource_manager.cc:51-52); with w4a16 to int4 group-128.
 * random_init fills every weight with the seeded synthetic generator. */
B2LLM_API int32_t b2llm_engine_load_weight(b2llm_engine* e, int32_t kind, int32_t layer, const void* host_fp16,
                                           uint64_t num_elements);
/* same, but the host tensor already is THIS RANK's tensor-parallel shard, as a ppl.pmx export stores it in
 * model_slice_<rank>/model.onnx (resource_manager.cc:280-286; docs/llama_guide.md:14-36): QKV fp16
 * [(nq/tp + 2 nkv/tp) * head_dim, hidden] (local q heads, k heads, v heads), O [hidden, nq/tp * head_dim],
 * GATE / UP [intermediate/tp, hidden], DOWN [hidden, intermediate/tp]; embedding, lm_head and the norms are whole. */
B2LLM_API int32_t b2llm_engine_load_weight_shard(b2llm_engine* e, int32_t kind, int32_t layer, const void* host_fp16,
                                                 uint64_t num_elements);
B2LLM_API int32_t b2llm_engine_random_init(b2llm_engine* e, uint64_t seed);

/* KV memory is owned by the caller (resource_manager.cc:344-362 cudaMalloc's it and
 * llm_engine.h:142-146 SetBufferPtr()s it into inputs 9 / 10).  kv_scale_device is NULL for the fp16 cache
 * (cache_quant_bit 0: input 10 is not bound, llm_engine.h:134-136). */
B2LLM_API int32_t b2llm_engine_bind_kv(b2llm_engine* e, void* kv_cache_device, void* kv_scale_device,
                                       uint64_t kv_cache_max_tokens);
/* bytes per cached token of this rank's slice: cb and sb of resource_manager.cc:381-388 */
B2LLM_API int32_t b2llm_engine_kv_bytes_per_token(const b2llm_engine* e, uint64_t* cache_bytes, uint64_t* scale_bytes);

/* ---- the step ------------------------------------------------------------------------------
 * replaces: SetInputTask (llm_engine.cc:29-111): host vectors of ModelInput -> device inputs.
 * Pointers are host arrays; the page table (cache_mode 1) is re-uploaded only when
 * req_list_changed != 0, as in llm_engine.cc:66-72.  Copies are async on the engine stream. */
B2LLM_API int32_t b2llm_engine_set_inputs(b2llm_engine* e, const int64_t* token_ids, int64_t num_tokens,
                                          const int64_t* seq_starts, const int64_t* kv_starts,
                                          const int64_t* start_pos, int64_t batch,
                                          const int64_t* cache_indices_or_page_list, int64_t max_pages,
                                          int64_t decoding_batches, int64_t max_seq_len, int64_t max_kv_len,
                                          int32_t req_list_changed);

/* replaces: RunModelTask -> ppl::nn::Runtime::Run() (llm_engine.cc:113-116), the whole forward.
 * b2llm_engine_run uses the inputs staged by set_inputs; b2llm_engine_forward takes explicit
 * device pointers.  On return *logits_device is fp32 [batch, *logits_stride] (stride >= vocab),
 * valid until the next forward (output 0, llm_engine.cc:200,219-222). */
B2LLM_API int32_t b2llm_engine_run(b2llm_engine* e, int32_t cache_prefill, float** logits_device,
                                   int64_t* logits_stride);
B2LLM_API int32_t b2llm_engine_forward(b2llm_engine* e, const b2llm_step* step, float** logits_device,
                                       int64_t* logits_stride);
/* device pointers of the inputs staged by set_inputs (what llm_engine.cc:206-210 hands to the
 * penalty kernel): token_ids, seq_starts, start_pos */
B2LLM_API int32_t b2llm_engine_staged_inputs(b2llm_engine* e, const int64_t** token_ids, const int64_t** seq_starts,
                                             const int64_t** start_pos);
/* number of kernels the last forward launched (bench.py's gpu_launches) */
B2LLM_API int64_t b2llm_engine_last_launch_count(const b2llm_engine* e);
/* per-kernel-class device timing with CUDA events on the engine stream (bench.py's roofline leg).
 * classes: 0 attention (decode attention kernel [+ split merge, + prefill attention]), 1 the layer
 * GEMMs, 2 lm_head, 3 tensor-parallel exchanges (the fused residual joins or ncclAllReduce, + the logits
 * all-gather).  profile(e, 1) starts recording; profile_read sums the spans recorded since,
 * synchronises the stream, and resets. */
B2LLM_API int32_t b2llm_engine_profile(b2llm_engine* e, int32_t enable);
B2LLM_API int32_t b2llm_engine_profile_read(b2llm_engine* e, double* ms_by_class, int64_t* count_by_class,
                                            int32_t num_classes);
/* tensor parallelism: phase clocks of the fused residual-join kernel (csrc/tp_join.cu) accumulated since the last
 * call, read from the device and reset: out8 = { calls, ns waiting for the peers' partial sums (rank skew), ns
 * reducing / normalising / delivering this rank's rows, ns waiting for the peers' rows to land, (scratch), and for the
 * first row of CTA 0: ns until its peer loads arrived, ns of reductions + quantisation + issuing the peer stores, ns in
 * the system-scope fence }.  Zeros when the engine is not tensor parallel or uses the ncclAllReduce path
 * (B2LLM_TP_JOIN=nccl). */
B2LLM_API int32_t b2llm_engine_tp_join_stats(b2llm_engine* e, double* out8);
/* debugging / parity: copy an intermediate of the last forward to the host.
 * what: 0 residual stream fp16 [num_tokens, hidden] after the last layer; 1 qkv fp16 (last layer, after
 * rope); 2 attention output fp16 (last layer); 3 logits fp32 [batch, vocab] */
B2LLM_API int32_t b2llm_engine_debug_read(b2llm_engine* e, int32_t what, void* host_dst, uint64_t bytes);

/* ---- sampler / penalty ------------------------------------------------------------------------
 * replace: ppl::kernel::llm::cuda::pmx::sample_topk_topp(_get_workspace_size) and
 * pmx::apply_penalty, with the argument lists of their call sites
 * (src/backends/cuda/post_processor.cc:135, 190-193, 271-274). */
B2LLM_API int64_t b2llm_sample_topk_topp_get_workspace_size(int32_t batch, int32_t vocab_size, int32_t top_k);
B2LLM_API int32_t b2llm_sample_topk_topp(void* stream, const float* logits, const float* temperatures_optional,
                                         const float* top_p_optional, const float* rand_device, int32_t batch,
                                         int32_t vocab_size, int32_t batch_stride, int32_t top_k,
                                         float default_top_p, float default_rand, void* workspace,
                                         int32_t* output, float* logprobs);
B2LLM_API int32_t b2llm_apply_penalty(void* stream, const float* logits_in, const float* temperatures,
                                      const float* repetition_penalties, const float* presence_penalties_optional,
                                      const float* frequency_penalties_optional, const int64_t* batch_slots,
                                      const int64_t* token_inputs, const int64_t* seqstarts,
                                      const int64_t* start_pos, int32_t batch, int32_t vocab_size,
                                      uint16_t* penalty_count_map, float* logits_out);

/* ---- operator-level entry points (parity tests call the kernels one by one) ---------------------
 * All pointers are device pointers.  These are the K2..K11 device ops of SURVEY.md section 2.3. */

/* (skip-)RMSNorm + per-token int8 quantisation.  x fp16 [rows, hidden]; if skip != NULL the kernel
 * first forms x = fp16(x + skip) and writes it back to x (residual join).  q_out int8 [rows, hidden]
 * and scale_out fp32 [rows] when q_out != NULL, else y_out fp16 [rows, hidden] (unquantised). */
B2LLM_API int32_t b2llm_op_rmsnorm_quant(void* stream, void* x_fp16, const void* skip_fp16, const void* gamma_fp16,
                                         float eps, int64_t rows, int32_t hidden, int8_t* q_out, float* scale_out,
                                         void* y_out_fp16);
/* per-token int8 quantisation of fp16 rows */
B2LLM_API int32_t b2llm_op_quant_rows(void* stream, const void* x_fp16, int64_t rows, int32_t cols, int8_t* q_out,
                                      float* scale_out);
/* W8A8 GEMM: C[m, n] = epilogue((int32) sum_k A[m,k] * W[n,k], a_scale[m], w_scale[n]).
 * epilogue: 0 -> fp16 out [M, N]; 1 -> out = fp16(out + value) (residual add, in place);
 *           2 -> SwiGLU over interleaved (gate, up) column pairs -> fp16 out [M, N/2];
 * impl: 0 auto, 1 mma.sync baseline, 2 tcgen05 single-CTA kernel, 3 tcgen05 CTA-pair kernel (cta_group::2;
 *       falls back to the single-CTA kernel when M <= 128 or N % 256 != 0) */
B2LLM_API int32_t b2llm_op_gemm_w8a8(void* stream, const int8_t* a, const float* a_scale, const int8_t* w,
                                     const float* w_scale, int64_t M, int32_t N, int32_t K, int32_t epilogue,
                                     void* out_fp16, int32_t impl);
/* fp16 GEMM: A fp16 [M,K] x W fp16 [N,K]^T; epilogue 0/1/2 as above, 3 -> fp32 out [M, ldc] */
B2LLM_API int32_t b2llm_op_gemm_f16(void* stream, const void* a_fp16, const void* w_fp16, int64_t M, int32_t N,
                                    int32_t K, int32_t epilogue, void* out, int64_t ldc, int32_t impl);

/* KV geometry shared by the cache ops.  quant_group 8 = int8 cache + fp16 scales; quant_group 1 = fp16 cache (kv_cache
 * points at fp16 elements, kv_scale is NULL) */
typedef struct b2llm_kv_geom {
    int32_t num_layers, num_kv_heads, head_dim, quant_group;
    int32_t cache_layout, cache_mode, page_size, reserved;
    uint64_t max_tokens;
} b2llm_kv_geom;

/* RoPE on q,k (in place in qkv fp16 [num_tokens, (nq + 2 nkv) * head_dim]) + int8 group quantised
 * append of k,v into the cache of `layer`.  rope_cos / rope_sin fp32 [max_position, head_dim/2]. */
B2LLM_API int32_t b2llm_op_rope_kv_append(void* stream, void* qkv_fp16, const b2llm_step* step, int32_t num_heads,
                                          const b2llm_kv_geom* geom, int32_t layer, const float* rope_cos,
                                          const float* rope_sin, void* kv_cache, void* kv_scale);
/* attention for every token of the step: decode sequences read the int8 cache, prefill sequences
 * the fresh fp16 k/v in qkv (plus the cached prefix when step->cache_prefill).  out fp16
 * [num_tokens, nq * head_dim].  workspace: b2llm_attention_workspace_size bytes.
 * impl: 0 auto, 1 simple reference kernel for every token, 2 tensor-core kernels (split-KV flash-decoding for the
 * decode sequences, flash-attention forward for the prefill sequences; head_dim 128); 3 / 4 / 5 = 2 with the decode
 * kernel's TMA loader forced: 3 slim + merged K/V loads (the default where the layout allows), 4 dividing loader, 5 slim
 * loader; 6 / 8 = 2 with the prefill kernel forced: 6 tcgen05 / TMEM (the default; falls back to the mma.sync kernel when
 * the step has cached prefixes), 8 mma.sync */
B2LLM_API int64_t b2llm_attention_workspace_size(int64_t batch, int32_t num_heads, int32_t head_dim);
/* the decode attention kernel's launch plan (host logic only, needs no device): KV splits per sequence and warps per CTA
 * for `decoding_batches` sequences of at most `max_kv_len` cached tokens (one wave of one-warp CTAs where a cut into
 * 0.75 .. 1.0 of the 148 x 12 warp slots exists, else the cut with the best-filled last wave) -- the counterpart of the reference's
 * ENGINE_CONF_DECODING_ATTN_SPLIT_K = 1 "heuristic" and ENGINE_CONF_DECODING_ATTN_TPB knobs (resource_manager.cc:74-106) */
B2LLM_API int32_t b2llm_attention_decode_plan(int64_t decoding_batches, int32_t num_heads, int32_t num_kv_heads,
                                              int64_t max_kv_len, int32_t* nsplit, int32_t* warps);
/* Debug aid: per-CTA timeline of the decode attention kernel.  While a device buffer of 4 x capacity_ctas uint64 is armed,
 * every launch of the kernel with at most capacity_ctas CTAs writes {start ns, main loop entered ns, end ns, SM id}
 * (%globaltimer / %smid) at [4 * (blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z))]; later launches
 * overwrite earlier ones.  device_buf = NULL disarms.  Process-wide, not thread-safe against concurrent launches;
 * no reference counterpart. */
B2LLM_API int32_t b2llm_debug_attention_trace(void* device_buf, int64_t capacity_ctas);
B2LLM_API int32_t b2llm_op_attention(void* stream, const void* qkv_fp16, const b2llm_step* step, int32_t num_heads,
                                     const b2llm_kv_geom* geom, int32_t layer, const void* kv_cache,
                                     const void* kv_scale, void* workspace, void* out_fp16, int32_t impl);
/* rope table exactly as the engine builds it (host, fp32 [max_position, head_dim/2] each) */
B2LLM_API int32_t b2llm_rope_table(int32_t max_position, int32_t head_dim, float theta, float* cos_host,
                                   float* sin_host);
/* synthetic fp16 tensor (device), the generator random_init uses */
B2LLM_API int32_t b2llm_op_synth_fp16(void* stream, uint64_t seed, uint64_t tensor_id, uint64_t num_elements,
                                      float std, float mean, void* out_fp16);
/* per-output-channel int8 quantisation of an fp16 weight [N, K] (device) */
B2LLM_API int32_t b2llm_op_quant_weight(void* stream, const void* w_fp16, int32_t N, int32_t K, int8_t* q_out,
                                        float* scale_out);

/* W4A16 weight preparation (B2LLM_QUANT_W4A16): int4 group-128 quantisation of an fp16 weight [N, K] (device) into packed
 * nibbles [N, K/2] + fp16 scales [N, K/128], and its expansion to the fp16 GEMM operand fp16(q * scale) [N, K] */
B2LLM_API int32_t b2llm_op_quant_weight_w4(void* stream, const void* w_fp16, int32_t N, int32_t K, uint8_t* packed_out,
                                           void* scale_out_fp16);
B2LLM_API int32_t b2llm_op_dequant_w4(void* stream, const uint8_t* packed, const void* scale_fp16, int32_t N, int32_t K,
                                      void* w_out_fp16);

/* fused W4A16 GEMM: C[m, n] = sum_k A[m, k] (fp16) * fp16(q[n, k] * scale[n, k / 128]), fp32 accumulation on tcgen05;
 * epilogue 0 / 1 / 2 as b2llm_op_gemm_w8a8.  B2LLM_ERR_UNSUPPORTED when K % 128 != 0 or N % 128 != 0. */
B2LLM_API int32_t b2llm_op_gemm_w4a16(void* stream, const void* a_fp16, const uint8_t* packed, const void* scale_fp16,
                                      int64_t M, int32_t N, int32_t K, int32_t epilogue, void* out_fp16);

#ifdef __cplusplus
}
#endif
#endif /* B2LLM_H_ */
