// K12 / K13 of SURVEY.md section 2.3: apply_penalty and sample_topk_topp, the device operators behind
// CudaPostProcessor (src/backends/cuda/post_processor.cc:190-193, 271-274).
// Semantics: oracle/sampler_ref.py (the pmx kernels themselves are EXTERNAL).
//
// One CTA per row.  Greedy (top_k == 1) is one streaming pass: online (max, argmax, sum exp) per
// thread, merged across the CTA -- 128 KB of logits per row read once.  General top-k stages the
// row in shared memory when it fits (vocab * 4 B <= 200 KB, i.e. vocab <= 51 200) and runs k arg-max rounds over it,
// then one thread applies softmax / top-p / inverse-CDF over the k candidates.
// Guards: a temperature that is not > 0 samples that row greedily from the raw logits (no division by zero); a row
// without any comparable logit (all NaN) yields token 0 rather than an out-of-range id; top_k <= 0 is rejected on the host.
#include <map>
#include <mutex>

#include "common.cuh"

namespace b2llm {

namespace {

constexpr int kThreads = 512;
constexpr int64_t kPenaltySlots = 65536;  // batch slots tracked per count map (>= any max_running_batch in use)

struct Best {
    float v;
    int i;
};
__device__ __forceinline__ Best better(Best a, Best b) {  // larger value, then lower index
    return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}
__device__ __forceinline__ Best block_best(Best x, Best* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Best y{__shfl_xor_sync(0xffffffffu, x.v, o), __shfl_xor_sync(0xffffffffu, x.i, o)};
        x = better(x, y);
    }
    __syncthreads();
    if (lane == 0) scratch[warp] = x;
    __syncthreads();
    Best r = lane < nw ? scratch[lane] : Best{-INFINITY, 0x7fffffff};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Best y{__shfl_xor_sync(0xffffffffu, r.v, o), __shfl_xor_sync(0xffffffffu, r.i, o)};
        r = better(r, y);
    }
    return r;
}

// (m, s) pairs of an online log-sum-exp
__device__ __forceinline__ void lse_merge(float& m, float& s, float m2, float s2) {
    const float mn = fmaxf(m, m2);
    if (mn == -INFINITY) { m = mn; s = 0.f; return; }
    s = s * __expf(m - mn) + s2 * __expf(m2 - mn);
    m = mn;
}

__global__ void __launch_bounds__(kThreads) sample_kernel(const float* __restrict__ logits, const float* __restrict__ temps,
                                                         const float* __restrict__ top_ps, const float* __restrict__ rnd,
                                                         int vocab, int stride, int top_k, float default_top_p,
                                                         float default_rand, float* __restrict__ cand_ws,
                                                         int* __restrict__ out, float* __restrict__ logprobs,
                                                         int row_in_smem) {
    extern __shared__ float srow[];
    __shared__ Best sbest[32];
    __shared__ float sred[64];
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.x;
    const float* row = logits + (int64_t)b * stride;
    float T = temps ? temps[b] : 1.0f;
    if (!(T > 0.f)) {  // T <= 0 or NaN: greedy on the raw logits
        T = 1.0f;
        top_k = 1;
    }

    // pass 1: arg-max + log-sum-exp of l' = l / T
    Best best{-INFINITY, 0x7fffffff};
    float m = -INFINITY, s = 0.f;
    if ((vocab & 3) == 0 && (reinterpret_cast<uintptr_t>(row) & 15) == 0) {
        // 16-byte loads, one running-max update per four logits: the row (128 KB at vocab 32000) is read once and this
        // kernel is the step's host/device join (round 1: 80 us for 131 MB = 25 % of the HBM peak with scalar loads)
        const float4* row4 = reinterpret_cast<const float4*>(row);
        for (int i4 = threadIdx.x; i4 < (vocab >> 2); i4 += kThreads) {
            const float4 r4 = row4[i4];
            const float v[4] = {__fdiv_rn(r4.x, T), __fdiv_rn(r4.y, T), __fdiv_rn(r4.z, T), __fdiv_rn(r4.w, T)};
            const int i = i4 << 2;
            if (row_in_smem) *reinterpret_cast<float4*>(srow + i) = make_float4(v[0], v[1], v[2], v[3]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (v[j] > best.v) { best.v = v[j]; best.i = i + j; }
            const float vmax = fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3]));
            if (vmax > m) { s *= __expf(m - vmax); m = vmax; }
            s += (__expf(v[0] - m) + __expf(v[1] - m)) + (__expf(v[2] - m) + __expf(v[3] - m));
        }
    } else {
        for (int i = threadIdx.x; i < vocab; i += kThreads) {
            const float v = __fdiv_rn(row[i], T);
            if (row_in_smem) srow[i] = v;
            if (v > best.v) { best.v = v; best.i = i; }
            if (v > m) { s = s * __expf(m - v) + 1.f; m = v; } else { s += __expf(v - m); }
        }
    }
    Best top = block_best(best, sbest);
    {   // block log-sum-exp
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = kThreads >> 5;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
            lse_merge(m, s, m2, s2);
        }
        __syncthreads();
        if (lane == 0) { sred[warp] = m; sred[32 + warp] = s; }
        __syncthreads();
        m = lane < nw ? sred[lane] : -INFINITY;
        s = lane < nw ? sred[32 + lane] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
            lse_merge(m, s, m2, s2);
        }
    }
    const float lse = m + logf(s);

    if (top_k <= 1) {
        if (threadIdx.x == 0) {
            out[b] = (top.i >= 0 && top.i < vocab) ? top.i : 0;
            logprobs[b] = top.v - lse;
        }
        return;
    }

    // k selection rounds; candidates go to the workspace: values [k] then indices [k]
    const int k = min(top_k, vocab);
    float* cval = cand_ws + (int64_t)b * 2 * top_k;
    int* cidx = reinterpret_cast<int*>(cval + top_k);
    if (threadIdx.x == 0) { cval[0] = top.v; cidx[0] = top.i; }
    Best last = top;
    for (int r = 1; r < k; ++r) {
        // next candidate: the best element strictly after `last` in (value desc, index asc) order
        Best cur{-INFINITY, 0x7fffffff};
        for (int i = threadIdx.x; i < vocab; i += kThreads) {
            const float v = row_in_smem ? srow[i] : __fdiv_rn(row[i], T);
            const bool after = v < last.v || (v == last.v && i > last.i);
            if (after && (v > cur.v || (v == cur.v && i < cur.i))) { cur.v = v; cur.i = i; }
        }
        last = block_best(cur, sbest);
        if (threadIdx.x == 0) { cval[r] = last.v; cidx[r] = last.i; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const float top_p = top_ps ? top_ps[b] : default_top_p;
        const float r = rnd ? rnd[b] : default_rand;
        const float c0 = cval[0];
        float sum = 0.f;
        for (int i = 0; i < k; ++i) sum = __fadd_rn(sum, expf(cval[i] - c0));
        // cumulative probabilities (fp32 running sum of p_i = e_i / sum)
        int keep = k;
        float cum = 0.f, mass = 0.f;
        if (top_p <= 0.f) {
            keep = 1;
            mass = __fdiv_rn(expf(0.f), sum);
        } else {
            bool reached = false;
            for (int i = 0; i < k; ++i) {
                cum = __fadd_rn(cum, __fdiv_rn(expf(cval[i] - c0), sum));
                if (!reached && cum >= top_p) { keep = i + 1; mass = cum; reached = true; break; }
            }
            if (!reached) mass = cum;
        }
        const float thr = __fmul_rn(r, mass);
        int pick = keep - 1;
        cum = 0.f;
        for (int i = 0; i < keep; ++i) {
            cum = __fadd_rn(cum, __fdiv_rn(expf(cval[i] - c0), sum));
            if (cum > thr) { pick = i; break; }
        }
        out[b] = (cidx[pick] >= 0 && cidx[pick] < vocab) ? cidx[pick] : 0;
        logprobs[b] = cval[pick] - lse;
    }
}

// apply_penalty: one CTA per row
__global__ void __launch_bounds__(256) penalty_kernel(const float* __restrict__ in, const float* __restrict__ temps,
                                                     const float* __restrict__ rep, const float* __restrict__ presence,
                                                     const float* __restrict__ freq, const int64_t* __restrict__ slots,
                                                     const int64_t* __restrict__ tokens, const int64_t* __restrict__ seqstarts,
                                                     const int64_t* __restrict__ start_pos, int vocab,
                                                     uint16_t* __restrict__ count_map, int64_t* __restrict__ next_pos,
                                                     float* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.x;
    const int64_t slot = slots[b];
    uint16_t* cnt = count_map + slot * (int64_t)vocab;
    // A slot's counts are cleared on the FIRST step of a request.  start_pos == 0 identifies it only without the prefix
    // cache: a request admitted on a cache hit enters at start_pos = cache_hit_count (or hit - 1 with a single token,
    // llm_generator.cc:229-242) and looks like a decode step.  What a continuing request always satisfies is
    // start_pos == previous start_pos + previous token count (llm_generator.cc:706-717), so the library keeps the
    // expected next position per slot and clears the row whenever a step does not continue the previous one.
    __shared__ int s_reset;
    if (threadIdx.x == 0) {
        const int64_t sp = start_pos[b];
        bool reset = sp == 0;
        if (next_pos != nullptr && slot >= 0 && slot < kPenaltySlots) {
            reset |= next_pos[slot] != sp;
            next_pos[slot] = sp + (seqstarts[b + 1] - seqstarts[b]);
        }
        s_reset = reset ? 1 : 0;
    }
    __syncthreads();
    if (s_reset) {
        for (int i = threadIdx.x; i < vocab; i += blockDim.x) cnt[i] = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // few tokens per step; serial keeps duplicate tokens exact
        for (int64_t t = seqstarts[b]; t < seqstarts[b + 1]; ++t) {
            const int64_t id = tokens[t];
            if (id >= 0 && id < vocab && cnt[id] < 65535) cnt[id] += 1;
        }
    }
    __syncthreads();
    const float r = rep[b], T = temps[b];
    const float pp = presence ? presence[b] : 0.f, fp = freq ? freq[b] : 0.f;
    const float* irow = in + (int64_t)b * vocab;
    float* orow = out + (int64_t)b * vocab;
    for (int i = threadIdx.x; i < vocab; i += blockDim.x) {
        float l = irow[i];
        const int c = cnt[i];
        if (c > 0) {
            l = l > 0.f ? __fdiv_rn(l, r) : __fmul_rn(l, r);
            if (presence) l = __fsub_rn(l, pp);
        }
        if (freq) l = __fsub_rn(l, __fmul_rn(fp, (float)c));
        orow[i] = __fdiv_rn(l, T);
    }
}

}  // namespace
}  // namespace b2llm

using namespace b2llm;

extern "C" int64_t b2llm_sample_topk_topp_get_workspace_size(int32_t batch, int32_t vocab_size, int32_t top_k) {
    (void)vocab_size;
    if (top_k <= 1) return 0;
    return (int64_t)batch * 2 * top_k * (int64_t)sizeof(float);
}

extern "C" int32_t b2llm_sample_topk_topp(void* stream, const float* logits, const float* temperatures_optional,
                                          const float* top_p_optional, const float* rand_device, int32_t batch,
                                          int32_t vocab_size, int32_t batch_stride, int32_t top_k, float default_top_p,
                                          float default_rand, void* workspace, int32_t* output, float* logprobs) {
    B2_REQUIRE(logits && output && logprobs, B2LLM_ERR_INVALID_VALUE, "sample_topk_topp: null pointer");
    B2_REQUIRE(batch >= 0 && vocab_size > 0 && batch_stride >= vocab_size, B2LLM_ERR_INVALID_VALUE,
               "sample_topk_topp: bad shape");
    // the reference forwards its top_k unchanged (post_processor.cc:133-136, 190-193) and sizes no workspace for
    // top_k <= 0; "no top-k limit" would need a full-vocabulary sort this kernel does not implement -- refuse loudly
    // instead of silently sampling greedily
    B2_REQUIRE(top_k >= 1, B2LLM_ERR_INVALID_VALUE, "sample_topk_topp: top_k must be >= 1 (top_k <= 0 is not supported)");
    B2_REQUIRE(top_k <= 1 || workspace != nullptr, B2LLM_ERR_INVALID_VALUE, "sample_topk_topp: workspace required");
    if (batch == 0) return B2LLM_OK;
    cudaStream_t s = (cudaStream_t)stream;
    size_t smem = 0;
    int in_smem = 0;
    if (top_k > 1 && (size_t)vocab_size * sizeof(float) <= 200 * 1024) {
        smem = (size_t)vocab_size * sizeof(float);
        in_smem = 1;
        B2_ENSURE_DYN_SMEM(sample_kernel, 200 * 1024);  // a per-device attribute (single-process TP: one thread per GPU)
    }
    launch_kernel(sample_kernel, dim3(batch), dim3(kThreads), smem, s, logits, temperatures_optional, top_p_optional, rand_device,
                  vocab_size, batch_stride, top_k, default_top_p, default_rand, (float*)workspace, output, logprobs, in_smem);
    B2_LAUNCH_CHECK();
    return B2LLM_OK;
}

extern "C" int32_t b2llm_apply_penalty(void* stream, const float* logits_in, const float* temperatures,
                                       const float* repetition_penalties, const float* presence_penalties_optional,
                                       const float* frequency_penalties_optional, const int64_t* batch_slots,
                                       const int64_t* token_inputs, const int64_t* seqstarts, const int64_t* start_pos,
                                       int32_t batch, int32_t vocab_size, uint16_t* penalty_count_map,
                                       float* logits_out) {
    B2_REQUIRE(logits_in && logits_out && temperatures && repetition_penalties && batch_slots && token_inputs &&
                   seqstarts && start_pos && penalty_count_map,
               B2LLM_ERR_INVALID_VALUE, "apply_penalty: null pointer");
    if (batch == 0) return B2LLM_OK;
    // per (device, count map): expected next position of every batch slot, -1 = slot not seen yet
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, int64_t*> tracks;
    int64_t* next_pos = nullptr;
    {
        int dev = 0;
        B2_CHECK_CUDA(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lk(mu);
        auto key = std::make_pair(dev, (const void*)penalty_count_map);
        auto it = tracks.find(key);
        if (it == tracks.end()) {
            B2_CHECK_CUDA(cudaMalloc(&next_pos, kPenaltySlots * sizeof(int64_t)));
            B2_CHECK_CUDA(cudaMemsetAsync(next_pos, 0xFF, kPenaltySlots * sizeof(int64_t), (cudaStream_t)stream));
            tracks[key] = next_pos;
        } else {
            next_pos = it->second;
        }
    }
    launch_kernel(penalty_kernel, dim3(batch), dim3(256), 0, (cudaStream_t)stream, logits_in, temperatures, repetition_penalties,
                  presence_penalties_optional, frequency_penalties_optional, batch_slots, token_inputs, seqstarts, start_pos,
                  vocab_size, penalty_count_map, next_pos, logits_out);
    B2_LAUNCH_CHECK();
    return B2LLM_OK;
}
