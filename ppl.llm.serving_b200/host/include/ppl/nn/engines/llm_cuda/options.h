// ppl::nn::llm::cuda options (EXTERNAL): the names resource_manager.cc:43-67,74-112,239,248-265 and
// llm_engine.cc:114 use.  How b2llm honours each key is in host/src/pplnn_b200.cc (Engine::Configure).
#ifndef B2LLM_SHIM_PPL_NN_ENGINES_LLM_CUDA_OPTIONS_H_
#define B2LLM_SHIM_PPL_NN_ENGINES_LLM_CUDA_OPTIONS_H_

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string>
#include <vector>

namespace ppl { namespace nn { namespace llm { namespace cuda {

/** memory management policy */
enum {
    MM_PLAIN = 0,
    MM_COMPACT = 1,
};

enum {
    QUANT_METHOD_NONE = 0,
    QUANT_METHOD_ONLINE_I8I8 = 1,
    QUANT_METHOD_ONLINE_I4F16 = 2, // declared upstream, not selectable in the reference (resource_manager.cc:49-56)
};

enum {
    CUBLAS_LAYOUT_DEFAULT = 0,
    CUBLAS_LAYOUT_AMPERE = 1,
};

struct EngineOptions final {
    uint32_t device_id = 0;
    uint32_t mm_policy = MM_COMPACT;
    uint32_t quant_method = QUANT_METHOD_NONE;
    uint32_t cublas_layout_hint = CUBLAS_LAYOUT_DEFAULT;
    cudaStream_t runtime_stream = 0;
};

struct DeviceOptions final {
    uint32_t device_id = 0;
    uint32_t mm_policy = MM_COMPACT;
    cudaStream_t stream = 0;
};

struct HostDeviceOptions final {};

/** Engine::Configure keys */
enum {
    /** uint32_t: enable the shared-memory decode MHA algorithm */
    ENGINE_CONF_DECODING_SHM_MHA = 0,
    /** uint32_t: enable the online-softmax ("infinity") decode MHA algorithm */
    ENGINE_CONF_DECODING_INF_MHA,
    /** uint32_t: enable the online-softmax grouped-query decode algorithm */
    ENGINE_CONF_DECODING_INF_GQA,
    /** uint32_t: split-k decode attention: 0 off, 1 heuristic, 2 always */
    ENGINE_CONF_DECODING_ATTN_SPLIT_K,
    /** uint32_t: decode attention threads per block: 0 heuristic, 256, 512 */
    ENGINE_CONF_DECODING_ATTN_TPB,
    /** uint32_t: enable graph-level kernel fusion */
    ENGINE_CONF_GRAPH_FUSION,
    /** ncclComm_t: the tensor-parallel communicator of this rank */
    ENGINE_CONF_SET_TP_NCCL_COMM,
    /** uint32_t, per step: prefill sequences with start_pos > 0 read their cached prefix */
    ENGINE_CONF_CACHE_PREFILL,
    ENGINE_CONF_MAX,
};

/** DeviceContext::Configure keys */
enum {
    /** cudaStream_t*: the stream the context enqueues on */
    DEV_CONF_GET_STREAM = 0,
    DEV_CONF_MAX,
};

}}}} // namespace ppl::nn::llm::cuda

#endif
