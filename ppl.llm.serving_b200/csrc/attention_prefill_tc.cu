// K7 on the 5th-generation tensor cores: flash-attention forward for FRESH prompts (start_pos == 0) with tcgen05.mma,
// TMEM accumulators and TMA-staged Q / K / V tiles.
//
//   STATUS: parity-green on the device since round 2 run 1 (tests/test_ops_gpu.py::test_attention_prefill_tcgen05); selected
//   by B2LLM_PREFILL_IMPL=tc or b2llm_op_attention impl 6.  The first cut (one softmax thread per query row) ran at
//   219 TFLOP/s causal, barely above the mma.sync kernel (202): with one warp per scheduler and 128-long max / sum
//   dependency chains the softmax warps were latency-bound and the tensor pipe idled.  This version splits every row
//   over TWO threads (see below).
//
// Why: one whole prefill step of BASELINE config 5 spends 54 % of its time in the mma.sync prefill attention at
// 136-170 TFLOP/s (profiles/r1_prefill_step_run14.json) -- < 10 % of the fp16 tensor peak.
//
// One CTA = 128 consecutive queries of one (sequence, q head); key blocks of 128 run 0 .. q_tile (causal).  192 threads:
//   warp 0   TMA producer: Q once (2 boxes of 64 head-dims x 128 rows, 128 B swizzle), then per key block K and V
//            (2 + 2 boxes) into a 2-stage ring -- all straight from the step's qkv activation [T, (nq + 2 nkv) * 128];
//   warp 1   one thread issues  S_j = Q K_j^T   (A = Q K-major, B = K_j K-major; M = N = 128, 8 k-steps of 16)
//            and                PV_j = P_j V_j  (A = P_j K-major from smem, B = V_j MN-major: V is stored [key][d], i.e.
//            the N = d dimension is contiguous -- descriptor with LBO = stride between the two 64-d slabs, SBO = 1024,
//            b_major = 1 in the instruction descriptor; cute/atom/mma_traits_sm100.hpp make_umma_desc<Major::MN>);
//            S and PV each double-buffered in TMEM (4 x 128 columns = all 512);
//   warps 2-9  two threads per query row (warps 2-5: keys / head-dims [0, 64) of the row's TMEM lane, warps 6-9:
//            [64, 128) -- a warp may touch the TMEM lanes of quarter warp % 4, so both warps of a pair qualify): the
//            64 scores of the thread's half stay in registers between the row-max pass and the exp2 / row-sum / fp16 P_j
//            pass (P written K-major + swizzled into smem, fence.proxy.async); the two half-maxima meet through shared
//            memory and a 64-thread named barrier, the row sums stay per-thread partials until the end; one block later
//            O += PV_j from TMEM with the running-max correction, O kept in 64 fp32 registers per thread.  Two warps
//            per scheduler and four independent max / sum chains per thread instead of one 128-long chain.
// Pipeline: S_{j+1} is issued before PV_j, so the tensor core computes the next scores while the softmax warps work on
// S_j; the O update of block j is deferred until after P_{j+1} so that PV_j's latency is hidden as well.
//
// Numeric contract: oracle/llama_ref.py _attend (fp16 Q/K/V, fp32 scores and sums, P carried as hi + lo fp16 halves
// for the PV product like the mma.sync kernel: the tensor core sees ~22 bits of P).
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"
#include "tma_utils.cuh"

namespace b2llm {

namespace {

constexpr int D = 128;
constexpr int BQ = 128, BN = 128;
constexpr int SLAB = BQ * 128;             // 16 KB: 128 rows x 64 fp16 (one 128-byte swizzle atom per row)
constexpr int Q_BYTES = 2 * SLAB;          // 32 KB
constexpr int KV_STAGE = 4 * SLAB;         // K (2 slabs) + V (2 slabs) = 64 KB
constexpr int KV_STAGES = 2;
constexpr int P_BYTES = 2 * SLAB;          // 32 KB per P buffer; two of them: P as hi + lo fp16 halves (see below)
constexpr int SMEM_BYTES = Q_BYTES + KV_STAGES * KV_STAGE + 2 * P_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 1024 /*half-row exchange*/;
static_assert(SMEM_BYTES <= 227 * 1024, "attn_prefill_tc_kernel: shared memory budget");
constexpr int TMEM_COLS = 512;             // S0 | S1 | PV0 | PV1, 128 fp32 columns each
constexpr int NUM_THREADS = 320;           // TMA warp, MMA warp, 8 softmax warps

struct PrefillTcParams {
    const int64_t* seq_starts;
    int decoding_batches;
    int nq, nkv;
    float sl2;       // log2(e) / sqrt(D)
    __half* out;     // [T, nq * D]
};

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
    return y;
}

// K-major operand, 128-byte swizzle (Q, K, P): 8-row groups 1024 B apart; +32 B per k-step inside the atom
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)(1024 >> 4) << 32;   // stride byte offset
    d |= (uint64_t)1 << 46;             // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;             // SWIZZLE_128B
    return d;
}
// MN-major operand, 128-byte swizzle (V as B of the PV product): rows are keys (K dimension) of 64 contiguous head
// dims (128 B); 8-key groups 1024 B apart (SBO), the second 64-d slab `slab_bytes` further (LBO)
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t smem_addr, uint32_t slab_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((slab_bytes >> 4) & 0x3FFF) << 16;  // leading byte offset
    d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16, fp16 x fp16 -> fp32, M = 128, N = 128; b_mn: B operand is MN-major (bit 16)
__host__ __device__ constexpr uint32_t make_idesc_f16(bool b_mn) {
    uint32_t d = 0;
    d |= 1u << 4;                        // accumulator F32
    d |= (b_mn ? 1u : 0u) << 16;         // b_major
    d |= (uint32_t)(128 >> 3) << 17;     // N
    d |= (uint32_t)(128 >> 4) << 24;     // M
    return d;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
    attn_prefill_tc_kernel(PrefillTcParams p, const __grid_constant__ CUtensorMap map_qkv) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t sQ = smem_base, sKV = sQ + Q_BYTES, sP = sKV + KV_STAGES * KV_STAGE;
    const uint32_t bars = sP + 2 * P_BYTES;
    // barriers: 0 q_full | 1,2 kv_full | 3,4 kv_empty | 5,6 s_full | 7,8 s_empty | 9 p_full | 10 p_empty | 11,12 o_full | 13,14 o_empty
    auto bar = [&](int i) { return bars + 8u * i; };
    const uint32_t tmem_slot = bars + 8u * 16;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem + (tmem_slot - smem_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = p.decoding_batches + blockIdx.z;
    const int hq = blockIdx.y;
    const int hk = hq / (p.nq / p.nkv);
    const int q_tile = (int)gridDim.x - 1 - (int)blockIdx.x;  // heaviest (last) query tiles first
    const int64_t seq_tok0 = p.seq_starts[b];
    const int n = (int)(p.seq_starts[b + 1] - seq_tok0);
    const int q0 = q_tile * BQ;
    if (q0 >= n) return;  // whole CTA: nothing allocated yet
    const int nblk = q_tile + 1;  // causal, start_pos == 0: key blocks 0 .. q_tile

    if (threadIdx.x == 0) {
        mbar_init(bar(0), 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar(1 + s), 1);
            mbar_init(bar(3 + s), 1);
            mbar_init(bar(5 + s), 1);
            mbar_init(bar(7 + s), 256);
            mbar_init(bar(11 + s), 1);
            mbar_init(bar(13 + s), 256);
        }
        mbar_init(bar(9), 256);
        mbar_init(bar(10), 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    // BYTE columns in a qkv row (the tensor map is a byte tensor: coordinate 0 counts bytes)
    const int qcol = hq * D * 2, kcol = (p.nq + hk) * D * 2, vcol = (p.nq + p.nkv + hk) * D * 2;

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&map_qkv);
            mbar_expect_tx(bar(0), Q_BYTES);
            tma_load_2d(sQ, &map_qkv, bar(0), qcol, (int)(seq_tok0 + q0));
            tma_load_2d(sQ + SLAB, &map_qkv, bar(0), qcol + 128, (int)(seq_tok0 + q0));
            for (int j = 0; j < nblk; ++j) {
                const int s = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                mbar_wait(bar(3 + s), ph ^ 1);
                const uint32_t dst = sKV + s * KV_STAGE;
                const int row = (int)(seq_tok0 + j * BN);
                mbar_expect_tx(bar(1 + s), KV_STAGE);
                tma_load_2d(dst, &map_qkv, bar(1 + s), kcol, row);
                tma_load_2d(dst + SLAB, &map_qkv, bar(1 + s), kcol + 128, row);
                tma_load_2d(dst + 2 * SLAB, &map_qkv, bar(1 + s), vcol, row);
                tma_load_2d(dst + 3 * SLAB, &map_qkv, bar(1 + s), vcol + 128, row);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc_f16(false), idesc_pv = make_idesc_f16(true);
            auto issue_s = [&](int j) {
                const int s = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                mbar_wait(bar(1 + s), ph);        // K_j, V_j landed
                mbar_wait(bar(7 + s), ph ^ 1);    // S buffer s drained by the softmax warps
                tc_fence_after();
                const uint32_t sK = sKV + s * KV_STAGE;
#pragma unroll
                for (int k = 0; k < D / 16; ++k) {
                    const uint64_t da = desc_kmajor(sQ + (k >> 2) * SLAB) + 2 * (k & 3);
                    const uint64_t db = desc_kmajor(sK + (k >> 2) * SLAB) + 2 * (k & 3);
                    tc_mma_f16(tmem_base + s * 128, da, db, idesc_s, k != 0 ? 1u : 0u);
                }
                tc_commit(bar(5 + s));            // S_j complete
            };
            auto issue_pv = [&](int j) {
                const int s = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                mbar_wait(bar(9), (uint32_t)(j & 1));  // P_j written
                mbar_wait(bar(13 + s), ph ^ 1);        // PV buffer s consumed
                tc_fence_after();
                const uint32_t sV = sKV + s * KV_STAGE + 2 * SLAB;
                // P = hi + lo (two fp16 halves, ~22 bits): PV_j = P_hi V_j + P_lo V_j into the same accumulator.  With a
                // single fp16 P the op-level tolerance holds, but end to end the 2^-11 error of P moves the prefill rows by
                // ~1e-3 and W8A8 re-quantisation carries that into every later step (round 2 run 9: median logits error
                // 1.7e-3 instead of 1e-6) -- the mma.sync kernel splits P for the same reason.
#pragma unroll
                for (int half = 0; half < 2; ++half) {
#pragma unroll
                    for (int k = 0; k < BN / 16; ++k) {
                        const uint64_t da = desc_kmajor(sP + half * P_BYTES + (k >> 2) * SLAB) + 2 * (k & 3);
                        const uint64_t db = desc_mnmajor(sV + k * 16 * 128, SLAB);  // 16 keys = 16 rows of 128 B further
                        tc_mma_f16(tmem_base + 256 + s * 128, da, db, idesc_pv, (half | k) != 0 ? 1u : 0u);
                    }
                }
                tc_commit(bar(11 + s));           // PV_j complete
                tc_commit(bar(10));               // P buffer free
                tc_commit(bar(3 + s));            // K_j / V_j stage free
            };
            mbar_wait(bar(0), 0);                 // Q landed
            issue_s(0);
            for (int j = 0; j < nblk; ++j) {
                if (j + 1 < nblk) issue_s(j + 1);
                issue_pv(j);
            }
        }
    } else {
        const int quarter = warp & 3;             // TMEM lane quarter of this warp
        const int hf = (warp - 2) >> 2;           // which half of the row's keys / head dims this thread owns
        const int r = quarter * 32 + lane;        // query row inside the tile
        const int qrow = q0 + r;                  // position in the sequence (start_pos == 0)
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + hf * 64;
        float* smax = reinterpret_cast<float*>(smem + (bars + 8u * 20 - smem_base));  // [2 halves][128 rows], after the barriers
        constexpr int H = D / 2;                  // 64 columns per thread
        float o[H];
#pragma unroll
        for (int i = 0; i < H; ++i) o[i] = 0.f;
        float m_run = -INFINITY, l_run = 0.f, m_o = -INFINITY;  // l_run: this thread's partial row sum; m_o: the max o is relative to
        float m_blk_prev = -INFINITY;             // running max after the previous block (the reference of PV_{j-1})

        auto accumulate = [&](int j, float m_ref) {   // o += PV_j (this thread's 64 head dims), PV_j being relative to m_ref
            const int s = j & 1;
            const uint32_t ph = (j >> 1) & 1;
            mbar_wait(bar(11 + s), ph);
            tc_fence_after();
            const float corr = m_o == -INFINITY ? 0.f : ex2_approx(m_o - m_ref);
#pragma unroll
            for (int c = 0; c < H / 32; ++c) {
                uint32_t v[32];
                tmem_ld32(lane_addr + 256 + s * 128 + c * 32, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], corr, __uint_as_float(v[i]));
            }
            m_o = m_ref;
            tc_fence_before();
            mbar_arrive(bar(13 + s));
        };

        for (int j = 0; j < nblk; ++j) {
            const int s = j & 1;
            const uint32_t ph = (j >> 1) & 1;
            const bool last = j == nblk - 1;      // the only block with masked keys (diagonal / past the sequence end)
            mbar_wait(bar(5 + s), ph);
            tc_fence_after();
            // the thread's 64 scores: loaded once, scaled and masked in place, kept for both passes
            float sc[H];
            {
                uint32_t v0[32], v1[32];
                tmem_ld32(lane_addr + s * 128, v0);
                tmem_ld32(lane_addr + s * 128 + 32, v1);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    sc[i] = __uint_as_float(v0[i]) * p.sl2;
                    sc[32 + i] = __uint_as_float(v1[i]) * p.sl2;
                }
            }
            tc_fence_before();
            mbar_arrive(bar(7 + s));              // S buffer s may be overwritten: the scores live in registers now
            if (last) {
                const int key0 = j * BN + hf * H;
#pragma unroll
                for (int i = 0; i < H; ++i) {
                    const int key = key0 + i;
                    if (!(key <= qrow && key < n)) sc[i] = -INFINITY;
                }
            }
            float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int i = 0; i < H; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], sc[i]);
            const float mx_half = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            // the other half of the row lives in the partner warp (same quarter): exchange through smem
            smax[hf * 128 + r] = mx_half;
            asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
            const float mx = fmaxf(mx_half, smax[(hf ^ 1) * 128 + r]);
            asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");  // both read before the next block overwrites
            const float m_new = fmaxf(m_run, mx);
            const float m_safe = m_new == -INFINITY ? 0.f : m_new;
            l_run *= m_run == -INFINITY ? 0.f : ex2_approx(m_run - m_safe);
            m_run = m_new;
            // P = exp2(s - m), partial row sum, fp16 P into smem (K-major, 128 B swizzle: chunk c16 of row r at c16 ^ (r & 7));
            // this thread's 64 keys are exactly slab `hf` of the P operand
            mbar_wait(bar(10), (uint32_t)((j & 1) ^ 1));  // PV_{j-1} has read the P buffer
            float sum4[4] = {0.f, 0.f, 0.f, 0.f};
            uint8_t* prow = smem + (sP - smem_base) + hf * SLAB + r * 128;
#pragma unroll
            for (int c16 = 0; c16 < 8; ++c16) {
                uint32_t pk[4], pl[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int i = c16 * 8 + 2 * q;
                    const float p0 = ex2_approx(sc[i] - m_safe), p1 = ex2_approx(sc[i + 1] - m_safe);  // exp2(-inf) = 0 for masked keys
                    sum4[q] += p0 + p1;
                    const __half2 h = __floats2half2_rn(p0, p1);
                    const float2 hf2 = __half22float2(h);
                    const __half2 lo = __floats2half2_rn(p0 - hf2.x, p1 - hf2.y);
                    pk[q] = *reinterpret_cast<const uint32_t*>(&h);
                    pl[q] = *reinterpret_cast<const uint32_t*>(&lo);
                }
                *reinterpret_cast<uint4*>(prow + ((c16 ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                *reinterpret_cast<uint4*>(prow + P_BYTES + ((c16 ^ (r & 7)) << 4)) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
            }
            l_run += (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
            fence_proxy_async_smem();             // P: generic-proxy stores -> tensor core's async-proxy reads
            mbar_arrive(bar(9));                  // P_j ready (256 arrivals: both halves of every row)
            if (j > 0) accumulate(j - 1, m_blk_prev);  // deferred by one block: PV_{j-1} finished long ago
            m_blk_prev = m_safe;
        }
        accumulate(nblk - 1, m_blk_prev);

        // full row sum = the two partials
        smax[hf * 128 + r] = l_run;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
        const float l_row = l_run + smax[(hf ^ 1) * 128 + r];
        if (qrow < n) {
            const float inv = 1.f / l_row;
            __half* orow = p.out + (seq_tok0 + qrow) * (int64_t)p.nq * D + (int64_t)hq * D + hf * H;
#pragma unroll
            for (int c = 0; c < H / 8; ++c) {
                uint32_t w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const __half2 h = __floats2half2_rn(o[c * 8 + 2 * i] * inv, o[c * 8 + 2 * i + 1] * inv);
                    w[i] = *reinterpret_cast<const uint32_t*>(&h);
                }
                *reinterpret_cast<uint4*>(orow + c * 8) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    }
}

}  // namespace

// tcgen05 prefill attention for the prefill sequences [decoding_batches, batch) of the step.  Preconditions (else
// B2LLM_ERR_UNSUPPORTED and the caller uses the mma.sync kernel): head_dim 128, no cached prefix in the step
// (cache_prefill == 0, i.e. every prefill sequence starts at position 0), TMA available, qkv rows 16-byte aligned.
int32_t launch_attention_prefill_tc(cudaStream_t s, const AttnArgs& a) {
    const int64_t prefill_seqs = a.step->batch - a.step->decoding_batches;
    if (prefill_seqs <= 0 || a.step->max_seq_len <= 0) return B2LLM_OK;
    if (a.geom.head_dim != D || a.step->cache_prefill != 0 || !tma_available() || prefill_seqs > 65535 || a.num_heads > 65535)
        return B2LLM_ERR_UNSUPPORTED;
    const int heads = a.num_heads + 2 * a.geom.num_kv_heads;
    const uint64_t T = (uint64_t)a.step->num_tokens;
    CUtensorMap map;
    {
        const uint64_t dims[2] = {(uint64_t)heads * D * 2, T};       // bytes per row, rows
        const uint64_t strides[1] = {(uint64_t)heads * D * 2};
        const uint32_t box[2] = {128, (uint32_t)BQ};                  // 64 fp16 = one swizzle atom wide, 128 rows
        if (!tma_encode_bytes(&map, a.qkv, 2, dims, strides, box, true)) return B2LLM_ERR_UNSUPPORTED;
    }
    PrefillTcParams p{};
    p.seq_starts = a.step->seq_starts;
    p.decoding_batches = (int)a.step->decoding_batches;
    p.nq = a.num_heads;
    p.nkv = a.geom.num_kv_heads;
    p.sl2 = 1.4426950408889634f / sqrtf((float)D);
    p.out = a.out;
    B2_ENSURE_DYN_SMEM(attn_prefill_tc_kernel, SMEM_BYTES);
    dim3 grid((unsigned)((a.step->max_seq_len + BQ - 1) / BQ), (unsigned)a.num_heads, (unsigned)prefill_seqs);
    attn_prefill_tc_kernel<<<grid, NUM_THREADS, SMEM_BYTES, s>>>(p, map);
    B2_LAUNCH_CHECK();
    return B2LLM_OK;
}

}  // namespace b2llm
