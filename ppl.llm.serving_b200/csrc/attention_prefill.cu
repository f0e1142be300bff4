// K7 of SURVEY.md section 2.3: causal attention for the PREFILL sequences of a step (sequences
// [decoding_batches, batch), llm_generator.cc:229-242), ragged, with an optional cached prefix
// (start_pos > 0: prefix-cache hit, ENGINE_CONF_CACHE_PREFILL, llm_engine.cc:113-116).
//
// Flash-attention forward, one CTA (4 warps) per (64-query tile, q head, sequence); each warp owns 16 query rows.
// Keys are walked in blocks of 64: fresh K/V come straight out of this step's fp16 qkv activation with
// cp.async, cached-prefix K/V are dequantised (fp16(int8 * scale), the oracle's definition) on their way into the
// same shared-memory tile, so the math below is one code path.  S = Q K^T and O += P V run on mma.sync m16n8k16
// (fp32 accumulate); P goes from accumulator to A-operand layout in registers and is fed as hi + lo fp16 halves so
// that the probabilities keep ~22 bits (the end-to-end parity bar is 1e-3 on logits, tests/test_engine_gpu.py).
// Shared memory: 2 x (K 16 KB + V 16 KB), 16-byte chunks XOR-swizzled by (row & 7) for conflict-free ldmatrix.
//
// This is the tensor-core replacement of attn_simple_kernel for head_dim 128; a tcgen05/TMEM formulation (S and O
// in tensor memory) is the next step for the prefill microbench (BASELINE config 5).
// Numeric contract: oracle/llama_ref.py _attend.
#include "common.cuh"

namespace b2llm {

namespace {

constexpr int D = 128;
constexpr int BQ = 64;   // queries per CTA
constexpr int BK = 64;   // keys per block
constexpr int TILE_BYTES = BK * D * 2;  // 16 KB

struct PrefillParams {
    const __half* qkv;
    const int64_t* seq_starts;
    const int64_t* start_pos;
    const int64_t* cache_indices;
    int decoding_batches;
    int64_t max_pages;
    int nq, nkv;
    int cache_mode, page_size;
    const int8_t* cache;   // layer offset applied
    const __half* scale;   // layer offset applied
    int kv16;              // fp16 cache (cache_quant_bit 0): cached-prefix rows are copied, not dequantised
    KvStrides cs;
    float sl2;             // log2(e) / sqrt(D)
    __half* out;           // [T, nq * D]
};

__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int row, int chunk) {
    return base + (uint32_t)(row * (D * 2) + ((chunk ^ (row & 7)) << 4));
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}

// stage keys [k0, k0 + 64) of sequence b (absolute positions) into the K and V tiles
__device__ __forceinline__ void load_kv_block(const PrefillParams& p, int b, int hk, int64_t sp, int64_t seq_tok0, int64_t kv_end,
                                              int64_t k0, uint32_t sK, uint32_t sV) {
    const int heads = p.nq + 2 * p.nkv;
    // 64 keys x 16 chunks = 1024 chunks per tensor; 128 threads -> 8 per thread; a thread keeps its chunk column
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int idx = i * 128 + threadIdx.x;
        const int row = idx >> 4, chunk = idx & 15;
        const int64_t pos = k0 + row;
        const uint32_t dk = tile_addr(sK, row, chunk), dv = tile_addr(sV, row, chunk);
        if (pos >= kv_end) {
            // beyond the last visible key: zeros (masked anyway, but keep NaNs out of the MMAs)
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dk), "r"(0u));
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dv), "r"(0u));
        } else if (pos >= sp) {
            const __half* krow = p.qkv + (seq_tok0 + (pos - sp)) * (int64_t)heads * D + (int64_t)(p.nq + hk) * D + chunk * 8;
            cp_async16(dk, krow, 16);
            cp_async16(dv, krow + (int64_t)p.nkv * D, 16);
        } else {
            const int64_t slot = kv_slot(p.cache_indices, p.cache_mode, p.page_size, p.max_pages, b, pos);
            const int64_t off = hk * p.cs.head + slot * p.cs.tok + chunk * 8;
            if (p.kv16) {
                const __half* c16 = reinterpret_cast<const __half*>(p.cache) + off;
                cp_async16(dk, c16, 16);
                cp_async16(dv, c16 + p.cs.kv, 16);
                continue;
            }
#pragma unroll
            for (int kv = 0; kv < 2; ++kv) {
                const int64_t o = off + kv * p.cs.kv;
                const uint2 raw = *reinterpret_cast<const uint2*>(p.cache + o);
                const float sc = __half2float(p.scale[o / 8]);
                const int8_t* q8 = reinterpret_cast<const int8_t*>(&raw);
                uint32_t w[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const __half2 h = __floats2half2_rn(__fmul_rn((float)q8[2 * j], sc), __fmul_rn((float)q8[2 * j + 1], sc));
                    w[j] = *reinterpret_cast<const uint32_t*>(&h);
                }
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(kv ? dv : dk), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]));
            }
        }
    }
}

__global__ void __launch_bounds__(128) attn_prefill_kernel(PrefillParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int b = p.decoding_batches + blockIdx.z;
    const int hq = blockIdx.y;
    const int hk = hq / (p.nq / p.nkv);
    const int64_t seq_tok0 = p.seq_starts[b];
    const int n = (int)(p.seq_starts[b + 1] - seq_tok0);
    const int q0 = blockIdx.x * BQ;  // first query (index inside the sequence) of this CTA
    if (q0 >= n) return;
    const int64_t sp = p.start_pos[b];
    const int heads = p.nq + 2 * p.nkv;
    const int q_last = min(n, q0 + BQ) - 1;
    const int64_t kv_end = sp + q_last + 1;  // keys [0, kv_end) are visible to some query of the tile

    // ---- Q fragments of this warp's 16 rows, straight from global memory (A operand layout of m16n8k16)
    const int r_lo = q0 + warp * 16 + g, r_hi = r_lo + 8;
    uint32_t qf[8][4];
    {
        const __half* qlo = p.qkv + (seq_tok0 + min(r_lo, n - 1)) * (int64_t)heads * D + (int64_t)hq * D;
        const __half* qhi = p.qkv + (seq_tok0 + min(r_hi, n - 1)) * (int64_t)heads * D + (int64_t)hq * D;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            qf[ks][0] = *reinterpret_cast<const uint32_t*>(qlo + 16 * ks + 2 * t);
            qf[ks][1] = *reinterpret_cast<const uint32_t*>(qhi + 16 * ks + 2 * t);
            qf[ks][2] = *reinterpret_cast<const uint32_t*>(qlo + 16 * ks + 8 + 2 * t);
            qf[ks][3] = *reinterpret_cast<const uint32_t*>(qhi + 16 * ks + 8 + 2 * t);
        }
    }
    const int64_t qpos_lo = sp + r_lo, qpos_hi = sp + r_hi;  // absolute positions (causal limit) of the two rows

    float o[16][4];
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

    const int nblk = (int)((kv_end + BK - 1) / BK);
    load_kv_block(p, b, hk, sp, seq_tok0, kv_end, 0, sbase, sbase + TILE_BYTES);
    cp_async_commit();

    for (int blk = 0; blk < nblk; ++blk) {
        const uint32_t sK = sbase + (blk & 1) * 2 * TILE_BYTES, sV = sK + TILE_BYTES;
        if (blk + 1 < nblk) {
            const uint32_t nK = sbase + ((blk + 1) & 1) * 2 * TILE_BYTES;
            load_kv_block(p, b, hk, sp, seq_tok0, kv_end, (int64_t)(blk + 1) * BK, nK, nK + TILE_BYTES);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();

        // ---- S = Q K^T for 64 keys: 8 n-tiles x 8 k-steps
        float s[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {  // pairs of n-tiles: one ldmatrix.x4
                // matrix m: n-tile 2 np + (m >> 1), dims chunk 2 ks + (m & 1); lane supplies row (lane & 7) of matrix lane >> 3
                const int m = lane >> 3, rr = lane & 7;
                const int row = 8 * (2 * np + (m >> 1)) + rr;
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4(b0, b1, b2, b3, tile_addr(sK, row, 2 * ks + (m & 1)));
                mma_f16_16816(s[2 * np], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b0, b1);
                mma_f16_16816(s[2 * np + 1], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b2, b3);
            }
        }

        // ---- causal mask + online softmax (exp2 domain); thread holds rows g (regs 0,1) and g + 8 (regs 2,3)
        const int64_t kbase = (int64_t)blk * BK;
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int64_t kp = kbase + 8 * nt + 2 * t + c;
                s[nt][c] = kp <= qpos_lo ? s[nt][c] * p.sl2 : -INFINITY;
                s[nt][2 + c] = kp <= qpos_hi ? s[nt][2 + c] * p.sl2 : -INFINITY;
                mx[0] = fmaxf(mx[0], s[nt][c]);
                mx[1] = fmaxf(mx[1], s[nt][2 + c]);
            }
        }
        float corr[2], m_safe[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_run[r], mx[r]);
            m_safe[r] = m_new == -INFINITY ? 0.f : m_new;
            corr[r] = exp2f(m_run[r] - m_safe[r]);  // m_run == -inf -> 0
            m_run[r] = m_new;
            l_run[r] *= corr[r];
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1];
        }
        // P (fp32) -> hi + lo fp16 A fragments, 4 k-steps of 16 keys
        uint32_t ph[4][4], pl[4][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float p0 = exp2f(s[nt][0] - m_safe[0]), p1 = exp2f(s[nt][1] - m_safe[0]);
            const float p2 = exp2f(s[nt][2] - m_safe[1]), p3 = exp2f(s[nt][3] - m_safe[1]);
            l_run[0] += p0 + p1;
            l_run[1] += p2 + p3;
            const __half2 h01 = __floats2half2_rn(p0, p1), h23 = __floats2half2_rn(p2, p3);
            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
            const __half2 l01 = __floats2half2_rn(p0 - f01.x, p1 - f01.y), l23 = __floats2half2_rn(p2 - f23.x, p3 - f23.y);
            const int kk = nt >> 1, hi = nt & 1;  // A regs: {row g k0-7, row g+8 k0-7, row g k8-15, row g+8 k8-15}
            ph[kk][2 * hi] = *reinterpret_cast<const uint32_t*>(&h01);
            ph[kk][2 * hi + 1] = *reinterpret_cast<const uint32_t*>(&h23);
            pl[kk][2 * hi] = *reinterpret_cast<const uint32_t*>(&l01);
            pl[kk][2 * hi + 1] = *reinterpret_cast<const uint32_t*>(&l23);
        }

        // ---- O += P V: 4 k-steps (16 keys) x 16 n-tiles (8 dims)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int dp = 0; dp < 8; ++dp) {  // pairs of dim tiles: one ldmatrix.x4.trans
                // matrix m: keys 16 kk + 8 (m & 1) .. + 7, dims chunk 2 dp + (m >> 1)
                const int m = lane >> 3, rr = lane & 7;
                const int row = 16 * kk + 8 * (m & 1) + rr;
                uint32_t v0, v1, v2, v3;
                ldmatrix_x4_trans(v0, v1, v2, v3, tile_addr(sV, row, 2 * dp + (m >> 1)));
                mma_f16_16816(o[2 * dp], ph[kk][0], ph[kk][1], ph[kk][2], ph[kk][3], v0, v1);
                mma_f16_16816(o[2 * dp], pl[kk][0], pl[kk][1], pl[kk][2], pl[kk][3], v0, v1);
                mma_f16_16816(o[2 * dp + 1], ph[kk][0], ph[kk][1], ph[kk][2], ph[kk][3], v2, v3);
                mma_f16_16816(o[2 * dp + 1], pl[kk][0], pl[kk][1], pl[kk][2], pl[kk][3], v2, v3);
            }
        }
        __syncthreads();  // everyone is done with this buffer before it is refilled two iterations later
    }

    // ---- finalise: row sums across the quad, normalise, store fp16
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
    if (r_lo < n) {
        __half* orow = p.out + (seq_tok0 + r_lo) * (int64_t)p.nq * D + (int64_t)hq * D;
#pragma unroll
        for (int i = 0; i < 16; ++i)
            *reinterpret_cast<__half2*>(orow + 8 * i + 2 * t) = __floats2half2_rn(o[i][0] * inv0, o[i][1] * inv0);
    }
    if (r_hi < n) {
        __half* orow = p.out + (seq_tok0 + r_hi) * (int64_t)p.nq * D + (int64_t)hq * D;
#pragma unroll
        for (int i = 0; i < 16; ++i)
            *reinterpret_cast<__half2*>(orow + 8 * i + 2 * t) = __floats2half2_rn(o[i][2] * inv1, o[i][3] * inv1);
    }
}

}  // namespace

// attention for the prefill sequences [decoding_batches, batch) of the step (head_dim 128, int8 group-8 cache)
int32_t launch_attention_prefill_mma(cudaStream_t s, const AttnArgs& a) {
    B2_REQUIRE(a.geom.head_dim == 128 && (a.geom.quant_group == 8 || a.geom.quant_group == 1), B2LLM_ERR_UNSUPPORTED,
               "prefill attention (tensor-core path): head_dim 128; int8 group-8 or fp16 cache");
    const bool kv16 = a.geom.quant_group == 1;
    const int64_t prefill_seqs = a.step->batch - a.step->decoding_batches;
    if (prefill_seqs <= 0 || a.step->max_seq_len <= 0) return B2LLM_OK;
    B2_REQUIRE(prefill_seqs <= 65535 && a.num_heads <= 65535, B2LLM_ERR_INVALID_VALUE, "too many prefill sequences / heads");
    PrefillParams p{};
    p.qkv = a.qkv;
    p.seq_starts = a.step->seq_starts;
    p.start_pos = a.step->start_pos;
    p.cache_indices = a.step->cache_indices;
    p.decoding_batches = (int)a.step->decoding_batches;
    p.max_pages = a.step->max_pages;
    p.nq = a.num_heads;
    p.nkv = a.geom.num_kv_heads;
    p.cache_mode = a.geom.cache_mode;
    p.page_size = a.geom.page_size;
    p.cs = kv_strides(a.geom);
    p.cache = a.kv_cache + (int64_t)a.layer * p.cs.layer * (kv16 ? 2 : 1);
    p.scale = kv16 ? nullptr : a.kv_scale + (int64_t)a.layer * p.cs.layer / a.geom.quant_group;
    p.kv16 = kv16 ? 1 : 0;
    p.sl2 = 1.4426950408889634f / sqrtf((float)D);
    p.out = a.out;
    constexpr int smem_bytes = 4 * TILE_BYTES;
    B2_ENSURE_DYN_SMEM(attn_prefill_kernel, smem_bytes);
    dim3 grid((unsigned)((a.step->max_seq_len + BQ - 1) / BQ), (unsigned)a.num_heads, (unsigned)prefill_seqs);
    attn_prefill_kernel<<<grid, 128, smem_bytes, s>>>(p);
    B2_LAUNCH_CHECK();
    return B2LLM_OK;
}

int32_t launch_attention_prefill(cudaStream_t s, const AttnArgs& a, int which) {
    // the tcgen05 / TMEM kernel (P as hi + lo fp16 halves) is the default since round 2 run 10: 519.2 vs 201.9 TFLOP/s causal
    // at 8 x 4096 tokens (profiles/r2_prefill_attention.txt; 602.7 with a single fp16 P, rejected on end-to-end parity);
    // steps with cached prefixes fall back to the mma.sync kernel inside it (B2LLM_ERR_UNSUPPORTED).
    // B2LLM_PREFILL_IMPL=mma keeps the mma.sync kernel for everything.
    static const bool env_tc = [] {
        const char* e = getenv("B2LLM_PREFILL_IMPL");
        return !(e != nullptr && e[0] == 'm');
    }();
    if (which == 1 || (which < 0 && env_tc)) {
        const int32_t rc = launch_attention_prefill_tc(s, a);
        if (rc != B2LLM_ERR_UNSUPPORTED) return rc;
    }
    return launch_attention_prefill_mma(s, a);
}

}  // namespace b2llm
