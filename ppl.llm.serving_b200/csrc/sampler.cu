// K12 / K13 of SURVEY.md section 2.3: apply_penalty and sample_topk_topp, the device operators behind
// CudaPostProcessor (src/backends/cuda/post_processor.cc:190-193, 271-274).
// Semantics: oracle/sampler_ref.py (the pmx kernels themselves are EXTERNAL).
//
// One CTA per row.  Greedy (top_k == 1) is one streaming pass: online (max, argmax, sum exp) per
// thread, merged across the CTA -- 128 KB of logits per row read once.  General top-k stages the
// row in shared memory when it fits (vocab <= 56 K) and runs k arg-max rounds over it, then one
// thread applies softmax / top-p / inverse-CDF over the k candidates.
#include "common.cuh"

namespace b2llm {

namespace {

constexpr int kThreads = 512;

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
    const int b = blockIdx.x;
    const float* row = logits + (int64_t)b * stride;
    const float T = temps ? temps[b] : 1.0f;

    // pass 1: arg-max + log-sum-exp of l' = l / T
    Best best{-INFINITY, 0x7fffffff};
    float m = -INFINITY, s = 0.f;
    for (int i = threadIdx.x; i < vocab; i += kThreads) {
        const float v = __fdiv_rn(row[i], T);
        if (row_in_smem) srow[i] = v;
        if (v > best.v) { best.v = v; best.i = i; }
        if (v > m) { s = s * __expf(m - v) + 1.f; m = v; } else { s += __expf(v - m); }
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
            out[b] = top.i;
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
        out[b] = cidx[pick];
        logprobs[b] = cval[pick] - lse;
    }
}

// apply_penalty: one CTA per row
__global__ void __launch_bounds__(256) penalty_kernel(const float* __restrict__ in, const float* __restrict__ temps,
                                                     const float* __restrict__ rep, const float* __restrict__ presence,
                                                     const float* __restrict__ freq, const int64_t* __restrict__ slots,
                                                     const int64_t* __restrict__ tokens, const int64_t* __restrict__ seqstarts,
                                                     const int64_t* __restrict__ start_pos, int vocab,
                                                     uint16_t* __restrict__ count_map, float* __restrict__ out) {
    const int b = blockIdx.x;
    uint16_t* cnt = count_map + slots[b] * (int64_t)vocab;
    if (start_pos[b] == 0) {
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
    B2_REQUIRE(top_k <= 1 || workspace != nullptr, B2LLM_ERR_INVALID_VALUE, "sample_topk_topp: workspace required");
    if (batch == 0) return B2LLM_OK;
    cudaStream_t s = (cudaStream_t)stream;
    size_t smem = 0;
    int in_smem = 0;
    if (top_k > 1 && (size_t)vocab_size * sizeof(float) <= 200 * 1024) {
        smem = (size_t)vocab_size * sizeof(float);
        in_smem = 1;
        static size_t configured = 0;
        if (smem > configured) {
            B2_CHECK_CUDA(cudaFuncSetAttribute(sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            configured = 200 * 1024;
        }
    }
    sample_kernel<<<batch, kThreads, smem, s>>>(logits, temperatures_optional, top_p_optional, rand_device, vocab_size,
                                                batch_stride, top_k, default_top_p, default_rand, (float*)workspace,
                                                output, logprobs, in_smem);
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
    penalty_kernel<<<batch, 256, 0, (cudaStream_t)stream>>>(logits_in, temperatures, repetition_penalties,
                                                           presence_penalties_optional, frequency_penalties_optional,
                                                           batch_slots, token_inputs, seqstarts, start_pos, vocab_size,
                                                           penalty_count_map, logits_out);
    B2_LAUNCH_CHECK();
    return B2LLM_OK;
}
