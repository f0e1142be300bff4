// Shared device/host helpers for the b2llm kernels (sm_100a only).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>
#include <utility>
#include <vector>

#include "../../include/b2llm.h"

namespace b2llm {

// ------------------------------------------------------------------ host-side error plumbing
void set_last_error(const std::string& msg);
extern thread_local int64_t g_launch_count;  // kernels launched by this thread since last reset

#define B2_CHECK_CUDA(expr)                                                                  \
    do {                                                                                     \
        cudaError_t err__ = (expr);                                                          \
        if (err__ != cudaSuccess) {                                                          \
            ::b2llm::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(err__)); \
            return B2LLM_ERR_DEVICE;                                                         \
        }                                                                                    \
    } while (0)

#define B2_REQUIRE(cond, code, msg)             \
    do {                                        \
        if (!(cond)) {                          \
            ::b2llm::set_last_error(msg);       \
            return (code);                      \
        }                                       \
    } while (0)

#define B2_LAUNCH_CHECK()                                                                 \
    do {                                                                                  \
        ++::b2llm::g_launch_count;                                                        \
        cudaError_t err__ = cudaGetLastError();                                           \
        if (err__ != cudaSuccess) {                                                       \
            ::b2llm::set_last_error(std::string("kernel launch: ") + cudaGetErrorString(err__)); \
            return B2LLM_ERR_DEVICE;                                                      \
        }                                                                                 \
    } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE function attribute: in the reference's single-process
// tensor parallelism (one host thread per GPU, resource_manager.cc:410-418) every device must set it once.
// `mask` is a per-call-site bitmask of the devices already configured.
#define B2_ENSURE_DYN_SMEM(kern, bytes)                                                                        \
    do {                                                                                                       \
        static std::atomic<uint64_t> b2_smem_mask_{0};                                                         \
        int b2_dev_ = 0;                                                                                       \
        B2_CHECK_CUDA(cudaGetDevice(&b2_dev_));                                                                \
        const uint64_t b2_bit_ = 1ull << (b2_dev_ & 63);                                                       \
        if (!(b2_smem_mask_.load(std::memory_order_acquire) & b2_bit_)) {                                      \
            B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            b2_smem_mask_.fetch_or(b2_bit_, std::memory_order_release);                                        \
        }                                                                                                      \
    } while (0)


// ------------------------------------------------------------------ programmatic dependent launch (PDL)
// Every kernel of the step is launched with cudaLaunchAttributeProgrammaticStreamSerialization: it may be scheduled
// while its predecessor in the stream is still draining, runs its prologue (barrier init, TMEM allocation, tensor-map
// prefetch, index arithmetic) and blocks in griddepcontrol.wait until the predecessor has COMPLETED and its writes are
// visible -- the ~325 launch / drain gaps of a decode step overlap instead of adding up.  Rules every kernel follows:
//   * pdl_trigger() first (dependents may be scheduled once every CTA of this grid has started);
//   * no global-memory access that depends on an earlier kernel before pdl_wait().
// B2LLM_PDL=0 launches everything fully serialised (the instructions are no-ops then).
bool pdl_enabled();

template <class... KArgs, class... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// ------------------------------------------------------------------ KV addressing
// Element strides of the int8 cache for the reference's four layouts (llm_engine.cc:118-169).
// The fp16 scale tensor has the same strides divided by quant_group.
struct KvStrides {
    int64_t layer, kv, head, tok;
};

inline KvStrides kv_strides(const b2llm_kv_geom& g) {
    const int64_t D = g.head_dim, H = g.num_kv_heads, L = g.num_layers, T = (int64_t)g.max_tokens;
    KvStrides s{};
    switch (g.cache_layout) {
        case 0: s.tok = L * 2 * H * D; s.layer = 2 * H * D; s.kv = H * D; s.head = D; break;
        case 1: s.layer = T * 2 * H * D; s.tok = 2 * H * D; s.kv = H * D; s.head = D; break;
        case 2: s.layer = 2 * T * H * D; s.kv = T * H * D; s.tok = H * D; s.head = D; break;
        default: s.layer = 2 * H * T * D; s.kv = H * T * D; s.head = T * D; s.tok = D; break;
    }
    return s;
}

// ------------------------------------------------------------------ device helpers
#ifdef __CUDACC__

__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// block-wide reductions through shared memory; `scratch` holds >= 32 floats
__device__ __forceinline__ float block_sum(float v, float* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float r = (lane < nw) ? scratch[lane] : 0.f;
    return warp_sum(r);
}
__device__ __forceinline__ float block_max(float v, float* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float r = (lane < nw) ? scratch[lane] : -INFINITY;
    return warp_max(r);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 16-byte async copy global -> shared; src_bytes == 0 zero-fills (predicated-off rows)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}

__device__ __forceinline__ void mma_f16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                              uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_s8_16832(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                             uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// silu(g) * u with every operation individually rounded (matches oracle/llama_ref.py:silu_mul)
__device__ __forceinline__ float silu_mul_f32(float g, float u) {
    const float s = __fdiv_rn(g, __fadd_rn(1.0f, expf(-g)));
    return __fmul_rn(s, u);
}

// sequence index of a token: largest b with seq_starts[b] <= t   (seq_starts has batch + 1 entries)
__device__ __forceinline__ int find_seq(const int64_t* __restrict__ seq_starts, int batch, int64_t t) {
    int lo = 0, hi = batch - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (seq_starts[mid] <= t) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// cache slot of position p of sequence b (llm_engine.cc:60-72; oracle/llama_ref.py:Step.slots)
__device__ __forceinline__ int64_t kv_slot(const int64_t* __restrict__ cache_indices, int cache_mode, int page_size,
                                           int64_t max_pages, int b, int64_t p) {
    if (cache_mode == 0) return cache_indices[b] + p;
    const int64_t page = cache_indices[(int64_t)b * max_pages + p / page_size];
    return page + p % page_size;
}

#endif  // __CUDACC__

// ------------------------------------------------------------------ kernel launchers (one per .cu)
int32_t launch_rmsnorm_quant(cudaStream_t s, __half* x, const __half* skip, const __half* gamma, float eps,
                             int64_t rows, int hidden, int8_t* q, float* scale, __half* y);
int32_t launch_quant_rows(cudaStream_t s, const __half* x, int64_t rows, int cols, int8_t* q, float* scale);
int32_t launch_embedding(cudaStream_t s, const int64_t* ids, const __half* table, int64_t n, int hidden, int vocab,
                         __half* out);
int32_t launch_gather_rows(cudaStream_t s, const __half* x, const int64_t* seq_starts, int64_t batch, int hidden,
                           __half* out);
int32_t launch_interleave_blocks(cudaStream_t s, const float* src, int parts, int64_t rows, int cols, float* dst);

int32_t launch_gemm_mma(cudaStream_t s, bool is_i8, const void* a, const float* a_scale, const void* w,
                        const float* w_scale, int64_t M, int N, int K, int epilogue, void* out, int64_t ldc);
int32_t launch_gemm_tc(cudaStream_t s, bool is_i8, const void* a, const float* a_scale, const void* w,
                       const float* w_scale, int64_t M, int N, int K, int epilogue, void* out, int64_t ldc,
                       int pair_mode = -1);  // -1 default (B2LLM_GEMM_2CTA), 0 single-CTA kernel, 2 CTA-pair kernel
bool gemm_tc_available();
int32_t launch_gemm_w4a16(cudaStream_t s, const void* a_fp16, const uint8_t* packed, const void* scale_fp16, int64_t M, int N,
                          int K, int epilogue, void* out, int64_t ldc);

struct AttnArgs {
    const __half* qkv;     // [T, (nq + 2 nkv) * D]
    const b2llm_step* step;
    int num_heads;         // q heads (this rank)
    b2llm_kv_geom geom;
    int layer;
    const int8_t* kv_cache;
    const __half* kv_scale;
    void* workspace;
    __half* out;           // [T, nq * D]
    int split_k = 1;       // ENGINE_CONF_DECODING_ATTN_SPLIT_K: 0 off, 1 heuristic, 2 always
    int loader = -1;       // decode kernel's TMA loader: -1 auto (env / default), 0 dividing, 1 slim, 2 slim + merged K/V loads
};
int32_t launch_rope_kv_append(cudaStream_t s, __half* qkv, const b2llm_step* step, int num_heads,
                              const b2llm_kv_geom& geom, int layer, const float* cos_t, const float* sin_t,
                              int8_t* kv_cache, __half* kv_scale);
int32_t launch_attention_simple(cudaStream_t s, const AttnArgs& a, int64_t token_begin, int64_t token_end);
int32_t launch_attention_decode_mma(cudaStream_t s, const AttnArgs& a);
int32_t launch_attention_prefill_mma(cudaStream_t s, const AttnArgs& a);
// tcgen05 / TMEM prefill attention (attention_prefill_tc.cu): B2LLM_ERR_UNSUPPORTED when its preconditions do not hold
// (cached prefixes in the step, head_dim != 128: the caller then uses the mma.sync kernel)
int32_t launch_attention_prefill_tc(cudaStream_t s, const AttnArgs& a);
// prefill attention of the step: the tcgen05 kernel where applicable, else the mma.sync kernel
// which: -1 default (tcgen05 unless B2LLM_PREFILL_IMPL=mma), 0 the mma.sync kernel, 1 the tcgen05 kernel
int32_t launch_attention_prefill(cudaStream_t s, const AttnArgs& a, int which = -1);  // sequences [decoding_batches, batch)
int64_t attention_workspace_bytes(int64_t batch, int num_heads, int head_dim);
// the decode kernel's launch plan for `base_ctas` = sequences x kv heads x q-head chunks (host logic, no device needed)
void attention_decode_plan(int64_t base_ctas, int64_t batch, int64_t max_kv_len, int* nsplit, int* warps);

// ---- tensor-parallel fused residual join over NVLink peer memory (tp_join.cu)
constexpr int kTpMaxRanks = 8;
struct TpPeers {            // every rank's communication buffer as mapped into THIS process / device
    int tp, rank;
    uint8_t* base[kTpMaxRanks];
};
struct TpLayout {           // byte offsets inside a communication buffer (identical on all ranks)
    size_t flags;           // 64 flag words + CTA counter + fault word
    size_t partial;         // fp16 [tokens, hidden]: this rank's row-parallel GEMM output
    size_t x;               // fp16 [tokens, hidden]: residual stream
    size_t q, qscale;       // int8 [tokens, act_cols] + fp32 [tokens]: quantised GEMM input
    size_t y;               // fp16 [tokens, hidden]: unquantised GEMM input (fp16 / W4A16 weights)
    size_t total;
};
TpLayout tp_layout(int64_t max_tokens, int hidden, int act_cols);
// mode 0: residual join only, 1: + RMSNorm -> int8 + scale, 2: + RMSNorm -> fp16; bcast_x: every rank gets the new x rows
int32_t launch_tp_join(cudaStream_t s, const TpPeers& peers, const TpLayout& L, int mode, bool bcast_x, const __half* gamma,
                       float eps, int64_t rows, int hidden, uint32_t epoch);
// collective over the engine's NCCL communicator: publish `local`, map every peer's buffer (peer access inside one
// process, CUDA IPC across processes; mappings opened through IPC are appended to *ipc_opened for cudaIpcCloseMemHandle)
int32_t tp_comm_exchange(cudaStream_t s, void* nccl_comm, int (*allgather)(const void*, void*, size_t, int, void*, cudaStream_t),
                         int rank, int tp, void* local, TpPeers* out, std::vector<void*>* ipc_opened);

int32_t launch_synth_fp16(cudaStream_t s, uint64_t seed, uint64_t tid, uint64_t n, float std, float mean, __half* out);
int32_t launch_synth_fp16_2d(cudaStream_t s, uint64_t seed, uint64_t tid, int64_t rows, int64_t cols, int64_t row0,
                             int64_t col0, int64_t full_cols, float std, float mean, __half* out);
int32_t launch_quant_weight(cudaStream_t s, const __half* w, int N, int K, int8_t* q, float* scale);
int32_t launch_quant_weight_w4(cudaStream_t s, const __half* w, int N, int K, uint8_t* packed, __half* scale);
int32_t launch_dequant_w4(cudaStream_t s, const uint8_t* packed, const __half* scale, int N, int K, __half* out);
int32_t launch_interleave_rows(cudaStream_t s, const __half* a, const __half* b, int rows, int cols, __half* out);

}  // namespace b2llm
