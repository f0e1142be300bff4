// Blackwell-native GEMM for the W8A8 projections (K3/K8/K9/K10) and the fp16 lm_head (K11):
//   C[M,N] = A[M,K] * W[N,K]^T,  int8 x int8 -> int32 (tcgen05.mma kind::i8, exact) or
//   fp16 x fp16 -> fp32 (kind::f16), accumulators in TMEM, operands staged by TMA.
//
// Structure (one persistent CTA per SM, 192 threads):
//   warp 0     TMA producer: cp.async.bulk.tensor 2D loads of a 128 x 128 B A tile and a BN x 128 B W
//              tile per stage (128-byte swizzle), mbarrier expect_tx / complete_tx
//   warp 1     TMEM allocator + MMA issuer: one lane issues 4 x tcgen05.mma (UMMA 128 x BN x 32 B) per
//              stage, tcgen05.commit releases the smem stage / publishes the accumulator
//   warps 2-5  epilogue: tcgen05.ld of the accumulator (one TMEM lane = one output row per thread),
//              fused dequant (a_scale[m] * w_scale[n]) + {fp16 store | residual add | SwiGLU | fp32 store}
// The accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps the
// main loop of tile i + 1.  Tiles are walked m-fastest so CTAs running side by side share the weight
// tile through L2.  Everything is expressed in BYTES of K: int8 k32 and fp16 k16 UMMA steps are both
// 32 bytes, so one kernel serves both element types.
//
// Results are bit-identical to gemm_mma.cu for int8 (integer accumulation, same fp32 epilogue).
#include <cuda.h>

#include <cstring>
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "tma_utils.cuh"

namespace b2llm {

namespace {

constexpr int BM = 128;
constexpr int BKB = 128;          // bytes of K per stage = one 128 B swizzle atom = 4 UMMA k-steps
// warps 2..9 are the epilogue: two warps per TMEM lane quarter (warp % 4), each draining half of the tile's columns.  With
// four warps -- one per scheduler -- the epilogue was latency-bound (~8 us per 256 x 256 tile, ~14 us with SwiGLU) and at
// M = 1024, where a CTA pair computes only 1..5 tiles, it was a fixed cost on top of every GEMM: measured times fit
// "main loop at the M = 8192 rate + 11 us" for qkv / o / down and "5 x epilogue" for gate_up (round 2 run 11).
constexpr int NUM_THREADS = 320;
constexpr int EPI_WARP0 = 2;
constexpr int EPI_THREADS = 256;

template <int BN>
struct Cfg {
    static constexpr int STAGES = BN == 256 ? 4 : 6;
    static constexpr int A_BYTES = BM * BKB;            // 16 KB
    static constexpr int B_BYTES = BN * BKB;            // 32 / 16 KB
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = 2 * BN;            // double-buffered accumulator
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

// ---------------------------------------------------------------- tcgen05 PTX wrappers (mbarrier / TMA ones: tma_utils.cuh)
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <bool I8>
__device__ __forceinline__ void tc_mma(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
    if constexpr (I8) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
            ::"r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum), "r"(0u) : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
            ::"r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum), "r"(0u) : "memory");
    }
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

// K-major, 128-byte swizzle operand descriptor (sm_100 version 1): 8-row groups are 1024 B apart
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);        // start address
    d |= (uint64_t)0 << 16;                            // leading byte offset (unused: one atom along K)
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset
    d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
    return d;
}

template <bool I8>
__host__ __device__ constexpr uint32_t make_idesc(int bn) {
    uint32_t d = 0;
    d |= (I8 ? 2u : 1u) << 4;              // accumulator: S32 / F32
    d |= (I8 ? 1u : 0u) << 7;              // A: signed int8 / fp16
    d |= (I8 ? 1u : 0u) << 10;             // B
    // a_major = b_major = 0 (K-major)
    d |= (uint32_t)(bn >> 3) << 17;        // N
    d |= (uint32_t)(BM >> 4) << 24;        // M
    return d;
}

// one epilogue chunk: 32 consecutive output columns [n0, n0 + 32) of row m, straight from the accumulator
template <bool I8, int EPI>
__device__ __forceinline__ void epilogue_store32(const uint32_t (&r)[32], int m, int n0, float sa_m,
                                                 const float* __restrict__ w_scale, void* __restrict__ out, int64_t ldc) {
    float v[32];
    if constexpr (I8) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {  // the 32 channel scales as eight 16-byte broadcast loads
            const float4 ws4 = __ldg(reinterpret_cast<const float4*>(w_scale + n0) + q);
            v[4 * q] = dequant((int)r[4 * q], sa_m, ws4.x);
            v[4 * q + 1] = dequant((int)r[4 * q + 1], sa_m, ws4.y);
            v[4 * q + 2] = dequant((int)r[4 * q + 2], sa_m, ws4.z);
            v[4 * q + 3] = dequant((int)r[4 * q + 3], sa_m, ws4.w);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
    }
    if constexpr (EPI == EPI_F16 || EPI == EPI_RESIDUAL) {
        __half* orow = reinterpret_cast<__half*>(out) + (int64_t)m * ldc + n0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint4 pk;
            uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
            if constexpr (EPI == EPI_RESIDUAL) {
                const uint4 old = *reinterpret_cast<const uint4*>(orow + q * 8);
                const uint32_t* ow = reinterpret_cast<const uint32_t*>(&old);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 o2 = __half22float2(*reinterpret_cast<const __half2*>(&ow[j]));
                    __half2 h = __floats2half2_rn(__fadd_rn(o2.x, v[q * 8 + 2 * j]), __fadd_rn(o2.y, v[q * 8 + 2 * j + 1]));
                    pw[j] = *reinterpret_cast<uint32_t*>(&h);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    __half2 h = __floats2half2_rn(v[q * 8 + 2 * j], v[q * 8 + 2 * j + 1]);
                    pw[j] = *reinterpret_cast<uint32_t*>(&h);
                }
            }
            *reinterpret_cast<uint4*>(orow + q * 8) = pk;
        }
    } else if constexpr (EPI == EPI_SWIGLU) {
        __half* orow = reinterpret_cast<__half*>(out) + (int64_t)m * ldc + (n0 >> 1);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            uint4 pk;
            uint32_t* pw = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = q * 16 + 4 * j;
                __half2 h = __floats2half2_rn(silu_mul_f32(v[i], v[i + 1]), silu_mul_f32(v[i + 2], v[i + 3]));
                pw[j] = *reinterpret_cast<uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(orow + q * 8) = pk;
        }
    } else {
        float* orow = reinterpret_cast<float*>(out) + (int64_t)m * ldc + n0;
#pragma unroll
        for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(orow + q * 4) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
}

template <bool I8, int EPI, int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                   const float* __restrict__ a_scale, const float* __restrict__ w_scale, int M, int N, int Kb,
                   void* __restrict__ out, int64_t ldc) {
    using C = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // 1024 B alignment for the swizzle atoms
    const uint32_t bars = smem_base + C::STAGES * C::STAGE_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * C::STAGES + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * C::STAGES + 2 + a); };
    const uint32_t tmem_slot = bars + 8u * (2 * C::STAGES + 4);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_m = (M + BM - 1) / BM, num_n = (N + BN - 1) / BN;
    const int num_tiles = num_m * num_n;
    const int nk = (Kb + BKB - 1) / BKB;  // a partial last k-block reads zeros beyond K (TMA out-of-bounds fill) in A and W
    pdl_trigger();

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), EPI_THREADS);
        }
        mbar_fence_init();
    }
    if (warp == 1) {  // whole warp allocates TMEM
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(C::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&map_a);
            tma_prefetch_desc(&map_w);
            // programmatic dependent launch: the WEIGHT tiles of the first ring fill depend on nothing an earlier kernel
            // of the step wrote -- they are requested before griddepcontrol.wait, so their HBM latency hides behind the
            // predecessor's tail.  The activation tiles (and everything else) come after the wait.
            const int pre = blockIdx.x < num_tiles ? (nk < C::STAGES ? nk : C::STAGES) : 0;
            for (int kb = 0; kb < pre; ++kb) {
                mbar_expect_tx(full_bar(kb), C::STAGE_BYTES);
                tma_load_2d(smem_base + kb * C::STAGE_BYTES + C::A_BYTES, &map_w, full_bar(kb), kb * BKB, (blockIdx.x / num_m) * BN);
            }
            pdl_wait();
            int stage = 0;
            uint32_t phase = 0;
            int done = 0;  // k-blocks issued so far: the first `pre` already have their barrier armed and weights in flight
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile % num_m, n_blk = tile / num_m;
                for (int kb = 0; kb < nk; ++kb, ++done) {
                    const uint32_t sa = smem_base + stage * C::STAGE_BYTES, sb = sa + C::A_BYTES;
                    if (done >= pre) {
                        mbar_wait(empty_bar(stage), phase ^ 1);
                        mbar_expect_tx(full_bar(stage), C::STAGE_BYTES);
                        tma_load_2d(sb, &map_w, full_bar(stage), kb * BKB, n_blk * BN);
                    }
                    tma_load_2d(sa, &map_a, full_bar(stage), kb * BKB, m_blk * BM);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc<I8>(BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_c = tmem_base + acc * BN;
                for (int kb = 0; kb < nk; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * C::STAGE_BYTES, sb = sa + C::A_BYTES;
                    const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sb);
#pragma unroll
                    for (int k = 0; k < BKB / 32; ++k)  // +32 B along K inside the swizzle atom = +2 in the address field
                        tc_mma<I8>(tmem_c, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    tc_commit(empty_bar(stage));  // frees the smem stage once these MMAs have read it
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(tfull_bar(acc));  // accumulator complete
            }
        }
    } else {
        // epilogue: TMEM lane quarter is fixed by warp id % 4
        const int quarter = warp & 3;
        int it = 0;
        pdl_wait();  // a_scale, the residual rows and the output buffer belong to earlier kernels
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int m_blk = tile % num_m, n_blk = tile / num_m;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const int m = m_blk * BM + quarter * 32 + lane;
            const float sa_m = (I8 && m < M) ? a_scale[m] : 1.f;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN;
            const int chalf = (warp - EPI_WARP0) >> 2;  // which half of the tile's 32-column chunks this warp drains
#pragma unroll 1
            for (int c = chalf * (BN / 64); c < (chalf + 1) * (BN / 64); ++c) {
                const int n0 = n_blk * BN + c * 32;
                if (n0 >= N) break;  // warp-uniform
                uint32_t r[32];
                tmem_ld32(taddr + c * 32, r);
                if (m < M) epilogue_store32<I8, EPI>(r, m, n0, sa_m, w_scale, out, ldc);
            }
            tc_fence_before();
            mbar_arrive(tempty_bar(acc));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS));
    }
}

// ---------------------------------------------------------------- CTA-pair kernel (tcgen05 cta_group::2)
// The 1-CTA kernel above is bounded by L2 -> shared-memory traffic, not by the tensor pipe: a 128 x 256 tile
// needs 48 KB per 128-byte k-block, ~12 TB/s at the rate the pipe could consume it (profiles/r1_ncu_gemm.txt:
// tensor pipe 47-49 % active).  A CTA pair computes a 256 x 256 tile with ONE UMMA of M = 256: each CTA stages
// its own 128 rows of A and only HALF of the weight tile (128 of the 256 output channels), the pair's tensor
// cores read both halves -- 32 KB per CTA per k-block for the same math, i.e. 2/3 of the traffic, and a six-
// instead of four-stage ring in the same shared memory.
//   * cluster (2,1,1); rank 0 = leader.  Both CTAs run the TMA producer (their own A rows / W half) with
//     cp.async.bulk.tensor ... .cta_group::2 completing on the LEADER's full barrier (mapa address);
//   * only the leader issues tcgen05.mma.cta_group::2 (M = 256, N = 256); tcgen05.commit ... multicast 0b11
//     releases the smem stage in both CTAs / publishes the accumulator to both epilogues;
//   * each CTA's epilogue drains its own TMEM (its 128 rows); both arrive on the leader's tempty barrier
//     (the peer through a remote mbarrier.arrive).
constexpr int BN2 = 256;
struct Cfg2 {
    static constexpr int STAGES = 6;
    static constexpr int A_BYTES = BM * BKB;                 // 16 KB: this CTA's 128 rows of A
    static constexpr int B_BYTES = (BN2 / 2) * BKB;          // 16 KB: this CTA's half of the weight tile
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;    // 32 KB
    static constexpr int TMEM_COLS = 2 * BN2;                // double-buffered accumulator
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// dst: this CTA's shared memory; bar: an mbarrier of either CTA of the pair (shared::cluster address)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// L2 prefetch of a tensor-map box (no shared-memory destination, no barrier): warms the weight tiles a ring fill ahead
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {  // same barrier offset in both CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
template <bool I8>
__device__ __forceinline__ void tc_mma_pair(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
    if constexpr (I8) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum) : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum) : "memory");
    }
}

template <bool I8>
__host__ __device__ constexpr uint32_t make_idesc_pair() {
    uint32_t d = 0;
    d |= (I8 ? 2u : 1u) << 4;
    d |= (I8 ? 1u : 0u) << 7;
    d |= (I8 ? 1u : 0u) << 10;
    d |= (uint32_t)(BN2 >> 3) << 17;       // N = 256
    d |= (uint32_t)((2 * BM) >> 4) << 24;  // M = 256 across the pair
    return d;
}

template <bool I8, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
    gemm_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                    const float* __restrict__ a_scale, const float* __restrict__ w_scale, int M, int N, int Kb,
                    void* __restrict__ out, int64_t ldc) {
    using C = Cfg2;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = smem_base + C::STAGES * C::STAGE_BYTES;
    auto full_bar = [&](int s) { return bars + 8u * s; };                          // leader's is used
    auto empty_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };           // per CTA
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * C::STAGES + a); };       // per CTA
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * C::STAGES + 2 + a); };  // leader's is used
    const uint32_t tmem_slot = bars + 8u * (2 * C::STAGES + 4);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int num_m = (M + 2 * BM - 1) / (2 * BM), num_n = N / BN2;
    const int num_tiles = num_m * num_n;
    const int nk = (Kb + BKB - 1) / BKB;  // a partial last k-block reads zeros beyond K (TMA out-of-bounds fill) in A and W
    pdl_trigger();

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 2 * EPI_THREADS);  // the epilogue threads of both CTAs
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(C::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer's barriers are initialised and its TMEM allocated before anything crosses
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&map_a);
            tma_prefetch_desc(&map_w);
            // weight halves of the first ring fill go out before griddepcontrol.wait (see gemm_tc_kernel)
            const int pre = pair < num_tiles ? (nk < C::STAGES ? nk : C::STAGES) : 0;
            for (int kb = 0; kb < pre; ++kb) {
                if (rank == 0) mbar_expect_tx(full_bar(kb), 2 * C::STAGE_BYTES);  // both CTAs' bytes
                tma_load_2d_pair(smem_base + kb * C::STAGE_BYTES + C::A_BYTES, &map_w, mapa_u32(full_bar(kb), 0), kb * BKB,
                                 (pair / num_m) * BN2 + (int)rank * (BN2 / 2));
            }
            // L2 prefetch distance.  At M = 1024 only four CTA pairs share a weight tile and they run in lock step, so every
            // weight load is an HBM miss; six 32 KB stages buy ~1.8 us of look-ahead, less than the loaded HBM latency, and the
            // tensor pipe sat at 50-53 % (profiles/r2_ncu_step_run9.txt; 87 % at M = 8192 where the tiles come from L2).  The
            // producer therefore asks L2 for the weight tile kPF k-blocks ahead of the one it stages.
            constexpr int kPF = 16;
            auto prefetch_w = [&](int g) {   // g-th k-block of this pair's sequence of tiles
                const int it2 = g / nk, kb2 = g - it2 * nk;
                const int tile2 = pair + it2 * num_pairs;
                if (tile2 < num_tiles) tma_prefetch_l2_2d(&map_w, kb2 * BKB, (tile2 / num_m) * BN2 + (int)rank * (BN2 / 2));
            };
            for (int g = pre; g < kPF; ++g) prefetch_w(g);
            pdl_wait();
            int stage = 0;
            uint32_t phase = 0;
            int done = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs) {
                const int m_blk = tile % num_m, n_blk = tile / num_m;
                const int row_a = m_blk * 2 * BM + (int)rank * BM;
                const int row_w = n_blk * BN2 + (int)rank * (BN2 / 2);
                for (int kb = 0; kb < nk; ++kb, ++done) {
                    prefetch_w(done + kPF);
                    const uint32_t sa = smem_base + stage * C::STAGE_BYTES, sb = sa + C::A_BYTES;
                    const uint32_t fb = mapa_u32(full_bar(stage), 0);
                    if (done >= pre) {
                        mbar_wait(empty_bar(stage), phase ^ 1);
                        if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * C::STAGE_BYTES);  // both CTAs' bytes
                        tma_load_2d_pair(sb, &map_w, fb, kb * BKB, row_w);
                    }
                    tma_load_2d_pair(sa, &map_a, fb, kb * BKB, row_a);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = make_idesc_pair<I8>();
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_c = tmem_base + acc * BN2;
                for (int kb = 0; kb < nk; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * C::STAGE_BYTES, sb = sa + C::A_BYTES;
                    const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sb);
#pragma unroll
                    for (int k = 0; k < BKB / 32; ++k)
                        tc_mma_pair<I8>(tmem_c, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    tc_commit_pair(empty_bar(stage));
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit_pair(tfull_bar(acc));
            }
        }
    } else {
        const int quarter = warp & 3;
        const uint32_t tempty_leader0 = mapa_u32(tempty_bar(0), 0), tempty_leader1 = mapa_u32(tempty_bar(1), 0);
        int it = 0;
        pdl_wait();  // a_scale, the residual rows and the output buffer belong to earlier kernels
        for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
            const int m_blk = tile % num_m, n_blk = tile / num_m;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const int m = m_blk * 2 * BM + (int)rank * BM + quarter * 32 + lane;
            const float sa_m = (I8 && m < M) ? a_scale[m] : 1.f;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN2;
            const int chalf = (warp - EPI_WARP0) >> 2;
#pragma unroll 1
            for (int c = chalf * (BN2 / 64); c < (chalf + 1) * (BN2 / 64); ++c) {
                const int n0 = n_blk * BN2 + c * 32;
                uint32_t r[32];
                tmem_ld32(taddr + c * 32, r);
                if (m < M) epilogue_store32<I8, EPI>(r, m, n0, sa_m, w_scale, out, ldc);
            }
            tc_fence_before();
            if (rank == 0) mbar_arrive(tempty_bar(acc));
            else mbar_arrive_remote(acc ? tempty_leader1 : tempty_leader0);
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // nobody leaves (or frees TMEM) while the pair may still signal / read it
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS));
    }
}

// ---------------------------------------------------------------- W4A16: int4 group-128 weights, dequant fused into operand staging
// C[M,N] = A[M,K] (fp16) x fp16(q[N,K] * scale[N, K/128])^T, fp32 accumulation in TMEM (tcgen05.mma kind::f16).
// The packed nibbles travel HBM -> smem as they are (0.5 B per weight, TMA, BN x 32 B per 64-element k-block); four
// converter warps expand them to the fp16 operand value fp16(q * scale) -- the oracle's definition, bit for bit --
// straight into the 128-byte-swizzled K-major tile the UMMA descriptor expects, make the writes visible to the async
// proxy and arrive on the stage's b_full barrier.  One weight tile feeds up to two 128-row A tiles (M <= 256 per
// CTA, the decode batch of BASELINE config 4 per rank), so the conversion cost is paid once per 256 rows.
//   warp 0     TMA producer (A tiles + packed W)          warps 2-9   converters (two threads per output channel: the
//   warp 1     TMEM alloc + MMA issuer                                k-block's elements [0, 32) and [32, 64))
//                                                         warps 10-13 epilogue (TMEM lane quarter = warp % 4)
// Eight converter warps, not four: with one warp per scheduler the expansion (a dependent lop / sub / mul / st chain per
// chunk) was latency-bound and the MMAs waited for it -- 365 TFLOP/s at M = 256 (round 1 run 10).
constexpr int W4_BN = 128;
constexpr int W4_THREADS = 448;  // TMA warp, MMA warp, 8 converter warps, 4 epilogue warps
template <int MT>
struct CfgW4 {
    static constexpr int STAGES = MT == 2 ? 4 : 5;
    static constexpr int A_BYTES = MT * BM * BKB;          // 16 / 32 KB
    static constexpr int RAW_BYTES = W4_BN * (BKB / 4);    // 4 KB: 64 nibbles = 32 B per row
    static constexpr int B_BYTES = W4_BN * BKB;            // 16 KB
    static constexpr int STAGE_BYTES = A_BYTES + RAW_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = MT == 2 ? 512 : 256;  // 2 buffers x MT x 128 columns
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 8 nibbles (one uint32, element i in bits [4i, 4i+4)) -> 8 fp16 values (q - 8) * scale as 4 packed half2.
// (w >> 4j) & 0x000F000F | 0x64006400 is the half2 (1024 + e_j, 1024 + e_{j+4}) in ONE lop3; the subtraction is exact
// and the product rounds once: fp16(q * s), what the oracle's dequantisation defines.  Four PRMTs put the pairs back in
// k order.  19 instructions per 8 elements (the byte-wise version took ~32 and made the converters issue-bound).
__device__ __forceinline__ uint32_t lop3_and_or(uint32_t x, uint32_t m, uint32_t k) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xea;\n" : "=r"(r) : "r"(x), "r"(m), "r"(k));  // (x & m) | k
    return r;
}
__device__ __forceinline__ uint32_t prmt_b32(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;\n" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}
__device__ __forceinline__ uint4 w4_expand8(uint32_t w, __half2 sc) {
    const __half2 bias = __halves2half2(__ushort_as_half(0x6408), __ushort_as_half(0x6408));  // 1032 = 1024 + 8
    uint32_t p[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t pk = lop3_and_or(w >> (4 * j), 0x000F000Fu, 0x64006400u);           // halves 1024 + e_j, 1024 + e_{j+4}
        __half2 h = *reinterpret_cast<__half2*>(&pk);
        h = __hmul2(__hsub2(h, bias), sc);
        p[j] = *reinterpret_cast<uint32_t*>(&h);
    }
    return make_uint4(prmt_b32(p[0], p[1], 0x5410u), prmt_b32(p[2], p[3], 0x5410u),    // (e0, e1) (e2, e3)
                      prmt_b32(p[0], p[1], 0x7632u), prmt_b32(p[2], p[3], 0x7632u));   // (e4, e5) (e6, e7)
}

template <int EPI, int MT>
__global__ void __launch_bounds__(W4_THREADS, 1)
    gemm_w4_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                   const __half* __restrict__ w_scale, int M, int N, int K, void* __restrict__ out, int64_t ldc, int splitk,
                   float* __restrict__ ws) {
    // splitk > 1 (few output tiles, e.g. qkv of 70B at TP = 8: N = 1280 -> 10 tiles on 148 SMs): a work item is
    // (tile, k-slice); every item stores its fp32 partial tile to ws[slice][M][N] and w4_splitk_reduce_kernel sums the
    // slices in slice order -- deterministic -- and applies the epilogue.  (A first version let the last item of a tile
    // do that inside this kernel: 10 CTAs reducing 18 MB with row-strided loads took 209 us, round 2 run 10.)
    using C = CfgW4<MT>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = smem_base + C::STAGES * C::STAGE_BYTES;
    auto afull_bar = [&](int s) { return bars + 8u * s; };
    auto bfull_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };
    auto empty_bar = [&](int s) { return bars + 8u * (2 * C::STAGES + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (3 * C::STAGES + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (3 * C::STAGES + 2 + a); };
    const uint32_t tmem_slot = bars + 8u * (3 * C::STAGES + 4);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_m = (M + MT * BM - 1) / (MT * BM), num_n = N / W4_BN;
    const int num_tiles = num_m * num_n;
    const int nk = (2 * K) / BKB;   // k-blocks of 64 elements (128 bytes of fp16)
    const int gpr = K / 128;        // scale groups per output channel
    const int num_items = num_tiles * splitk;
    auto kb_begin = [&](int ks) { return (int)((int64_t)nk * ks / splitk); };
    pdl_trigger();

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(afull_bar(s), 1);
            mbar_init(bfull_bar(s), 8);   // one elected lane per converter warp
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 128);
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(C::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();  // barriers, TMEM and the tensor maps are set up; the activations belong to the predecessor

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&map_a);
            tma_prefetch_desc(&map_w);
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                const int tile = item / splitk, ks = item - tile * splitk;
                const int m_blk = tile % num_m, n_blk = tile / num_m;
                for (int kb = kb_begin(ks); kb < kb_begin(ks + 1); ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t sa = smem_base + stage * C::STAGE_BYTES, sraw = sa + C::A_BYTES;
                    mbar_expect_tx(afull_bar(stage), C::A_BYTES + C::RAW_BYTES);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
                        tma_load_2d(sa + mt * BM * BKB, &map_a, afull_bar(stage), kb * BKB, (m_blk * MT + mt) * BM);
                    tma_load_2d(sraw, &map_w, afull_bar(stage), kb * (BKB / 4), n_blk * W4_BN);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc<false>(W4_BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
                const int ks = item % splitk;
                const int kb0 = kb_begin(ks), kb1 = kb_begin(ks + 1);
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_c = tmem_base + acc * (MT * W4_BN);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(afull_bar(stage), phase);
                    mbar_wait(bfull_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * C::STAGE_BYTES, sb = sa + C::A_BYTES + C::RAW_BYTES;
                    const uint64_t db = make_smem_desc(sb);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        const uint64_t da = make_smem_desc(sa + mt * BM * BKB);
#pragma unroll
                        for (int k = 0; k < BKB / 32; ++k)
                            tc_mma<false>(tmem_c + mt * W4_BN, da + 2 * k, db + 2 * k, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
                    }
                    tc_commit(empty_bar(stage));
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(tfull_bar(acc));
            }
        }
    } else if (warp < 10) {
        // converters: threads (r, hf) own output channel (row) r of the weight tile, hf = which 32 of the k-block's 64 elements
        const int r = (threadIdx.x - 64) & 127, hf = (threadIdx.x - 64) >> 7;
        int stage = 0;
        uint32_t phase = 0;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
            const int tile = item / splitk, ks = item - tile * splitk;
            const int n_blk = tile / num_m;
            const int kb0 = kb_begin(ks), kb1 = kb_begin(ks + 1);
            const __half* srow = w_scale + (int64_t)(n_blk * W4_BN + r) * gpr;
            // scales: one fp16 per 128 elements = per 2 k-blocks, and every thread walks its OWN row (rows are gpr halves
            // apart), so a per-k-block load is a 32-line-diverged LDG in each of the 8 warps -- ~256 LSU passes per k-block,
            // most of the 0.24 us an EMPTY k-block cost (round 2 run 28).  Four scales per 8-byte load, the next four in flight.
            const bool vec = (gpr & 3) == 0 && (reinterpret_cast<uintptr_t>(w_scale) & 7) == 0;
            auto load4 = [&](int grp) {     // scales [4 grp, 4 grp + 4) of this row as two packed half2
                if (vec) return __ldg(reinterpret_cast<const uint2*>(srow + 4 * grp));
                uint32_t h[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) h[i] = 4 * grp + i < gpr ? (uint32_t)__half_as_ushort(srow[4 * grp + i]) : 0u;
                return make_uint2(h[0] | (h[1] << 16), h[2] | (h[3] << 16));
            };
            int grp = kb0 >> 3;
            uint2 cur = load4(grp), nxt = cur;
            if (((grp + 1) << 3) < kb1) nxt = load4(grp + 1);
            for (int kb = kb0; kb < kb1; ++kb) {
                if ((kb >> 3) != grp) {
                    grp = kb >> 3;
                    cur = nxt;
                    if (((grp + 1) << 3) < kb1) nxt = load4(grp + 1);
                }
                const int j = (kb >> 1) & 3;
                const uint32_t word = (j & 2) ? cur.y : cur.x;
                const __half sc1 = __ushort_as_half((unsigned short)((j & 1) ? (word >> 16) : (word & 0xFFFFu)));
                mbar_wait(afull_bar(stage), phase);
                const uint32_t sraw = smem_base + stage * C::STAGE_BYTES + C::A_BYTES, sb = sraw + C::RAW_BYTES;
                const __half2 sc = __halves2half2(sc1, sc1);
                uint4 raw;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w)
                             : "r"(sraw + r * 32 + hf * 16));
                const uint32_t wds[4] = {raw.x, raw.y, raw.z, raw.w};
                uint4 v[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) v[c] = w4_expand8(wds[c], sc);  // 16-byte chunk 4 hf + c = elements [8 (4 hf + c), + 8)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t dst = sb + (uint32_t)(r * 128 + (((4 * hf + c) ^ (r & 7)) << 4));
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v[c].x), "r"(v[c].y), "r"(v[c].z), "r"(v[c].w) : "memory");
                }
                fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
                __syncwarp();              // every lane's writes + fence precede the warp's single arrival (256 arrivals
                if (lane == 0) mbar_arrive(bfull_bar(stage));  // on one mbarrier serialised: 0.24 us per k-block, run 28)
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        const int quarter = warp & 3;
        int it = 0;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
            const int tile = item / splitk, ks = item - tile * splitk;
            const int m_blk = tile % num_m, n_blk = tile / num_m;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int mt = 0; mt < MT; ++mt) {
                const int m = (m_blk * MT + mt) * BM + quarter * 32 + lane;
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * (MT * W4_BN) + mt * W4_BN;
#pragma unroll 1
                for (int c = 0; c < W4_BN / 32; ++c) {
                    const int n0 = n_blk * W4_BN + c * 32;
                    uint32_t rr[32];
                    tmem_ld32(taddr + c * 32, rr);
                    if (splitk == 1) {
                        if (m < M) epilogue_store32<false, EPI>(rr, m, n0, 1.f, nullptr, out, ldc);
                    } else if (m < M) {  // this slice's partial sums
                        float4* dst = reinterpret_cast<float4*>(ws + ((int64_t)ks * M + m) * N + n0);
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            dst[q] = make_float4(__uint_as_float(rr[4 * q]), __uint_as_float(rr[4 * q + 1]), __uint_as_float(rr[4 * q + 2]),
                                                 __uint_as_float(rr[4 * q + 3]));
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(tempty_bar(acc));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS));
    }
}

// ---------------------------------------------------------------- W4A16, transposed: the weight tile goes through TMEM
// The kernel above keeps BOTH MMA operands in shared memory: per 64-element k-block the tensor core reads 64 KB of it, TMA
// writes 36 KB and the converters move another 20 KB -- measured (round 2 runs 28-31: temporary skip-a-part switches and
// clock64 counters per role, profiles/r2_gemm_w4_transposed.txt) the MMAs
// alone then take 670 cycles per k-block against a floor of 512, the conversion alone 580, together 920: shared-memory
// bandwidth, not the tensor pipe.  This kernel computes the transposed product instead,
//     D^T[channel, row] = W[channel, k] . X[row, k]^T        M (MMA) = 128 output channels, N (MMA) = NA activation rows,
// with the dequantised weight tile as the MMA's A operand IN TENSOR MEMORY (tcgen05.mma [d], [a], b-desc): the
// converter threads write their 32 fp16 values per k-block straight into their own TMEM lane with tcgen05.st -- no
// shared-memory round trip, no proxy fence -- and the activations are the (single, N = 256) B operand.  Per k-block the
// shared memory now sees 32 KB of operand reads, 36 KB of TMA writes and 4 KB of nibble reads.
// The accumulator holds output channels in lanes: a warp's lanes are 32 consecutive channels of one activation row, so
// stores (fp16 outputs and the fp32 split-K partials alike) are contiguous per row.
//   warp 0  TMA producer (activations + packed W)      warps 2-9  converters, then the item's epilogue
//   warp 1  TMEM alloc + MMA issuer                                 (TMEM lane quarter = warp % 4, column half = (warp - 2) / 4)
constexpr int W4T_THREADS = 320;
template <int NA>
struct CfgW4T {
    static constexpr int ACT_BYTES = NA * BKB;             // 16 / 32 KB: NA activation rows x 64 fp16
    static constexpr int RAW_BYTES = W4_BN * (BKB / 4);    // 4 KB
    static constexpr int STAGE_BYTES = ACT_BYTES + RAW_BYTES;
    static constexpr int STAGES = NA == 256 ? 6 : 8;
    static constexpr int A_COLS = BKB / 4;                 // 32 columns: 64 fp16 per lane and k-block
    static constexpr int TMEM_COLS = 512;                  // NA accumulator columns + STAGES x 32 operand columns
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
    static_assert(NA + STAGES * A_COLS <= TMEM_COLS, "tensor memory budget");
};

__device__ __forceinline__ void tc_mma_ts_f16(uint32_t tmem_c, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_c), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accum), "r"(0u) : "memory");
}
// 16 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// 32 accumulator columns (= activation rows m0 .. m0 + 31) of this lane's output channel n -> memory.  part != nullptr: fp32
// split-K partials [M][N] of this slice; else the epilogue.  A warp's lanes are 32 consecutive channels of one row.
template <int EPI>
__device__ __forceinline__ void w4t_store32(const uint32_t (&rr)[32], int m0, int n, int lane, int M, int N, void* __restrict__ out,
                                            int64_t ldc, float* __restrict__ part) {
    if (part != nullptr) {   // 128 B per row and warp
        float* dst = part + (int64_t)m0 * N + n;
#pragma unroll
        for (int jj = 0; jj < 32; ++jj)
            if (m0 + jj < M) dst[(int64_t)jj * N] = __uint_as_float(rr[jj]);
    } else {
        // Lane pairs exchange so that every lane owns ONE row of a channel pair: even lanes row jj of channels (n, n + 1),
        // odd lanes row jj + 1 of (n - 1, n).  Three passes (exchange, loads, math + stores): a load placed after a store to
        // the same array cannot be hoisted by the compiler, and 64 dependent read-modify-writes cost 45K cycles per tile
        // (round 2 run 35); the SwiGLU chain (shuffle -> expf -> store) likewise ran one column at a time.
        const int odd = lane & 1;
        float a[16], b[16];   // values of channels (n & ~1) and (n & ~1) + 1 in this lane's row of each row pair
#pragma unroll
        for (int p = 0; p < 16; ++p) {
            const float v0 = __uint_as_float(rr[2 * p]), v1 = __uint_as_float(rr[2 * p + 1]);
            const float got = __shfl_xor_sync(0xffffffffu, odd ? v0 : v1, 1);
            a[p] = odd ? got : v0;
            b[p] = odd ? v1 : got;
        }
        if constexpr (EPI == EPI_SWIGLU) {   // (gate, up) = channels (2i, 2i + 1) -> one output column i
            __half* dst = reinterpret_cast<__half*>(out) + (int64_t)(m0 + odd) * ldc + (n >> 1);
#pragma unroll
            for (int p = 0; p < 16; ++p) a[p] = silu_mul_f32(a[p], b[p]);
#pragma unroll
            for (int p = 0; p < 16; ++p)
                if (m0 + 2 * p + odd < M) dst[(int64_t)(2 * p) * ldc] = __float2half_rn(a[p]);
        } else {
            __half2* dst = reinterpret_cast<__half2*>(reinterpret_cast<__half*>(out) + (int64_t)(m0 + odd) * ldc + (n & ~1));
            const int64_t step = ldc;   // two rows, in half2 units
            if constexpr (EPI == EPI_RESIDUAL) {
                __half2 old[16];
#pragma unroll
                for (int p = 0; p < 16; ++p) old[p] = m0 + 2 * p + odd < M ? dst[(int64_t)p * step] : __half2();
#pragma unroll
                for (int p = 0; p < 16; ++p) {
                    const float2 o2 = __half22float2(old[p]);
                    a[p] = __fadd_rn(o2.x, a[p]);
                    b[p] = __fadd_rn(o2.y, b[p]);
                }
            }
#pragma unroll
            for (int p = 0; p < 16; ++p)
                if (m0 + 2 * p + odd < M) dst[(int64_t)p * step] = __floats2half2_rn(a[p], b[p]);
        }
    }
}

template <int EPI, int NA>
__global__ void __launch_bounds__(W4T_THREADS, 1)
    gemm_w4t_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                    const __half* __restrict__ w_scale, int M, int N, int K, void* __restrict__ out, int64_t ldc, int splitk,
                    float* __restrict__ ws) {
    // Two k-slices are summed by w4_splitk_reduce_kernel through an fp32 scratch in L2.  Summing them inside the kernel --
    // clusters of two CTAs, rank 1 handing its 128 KB partial tile to rank 0 through distributed shared memory -- was built
    // and measured (round 2 run 33): gate_up 34.1 -> 47.4 us, down 22.8 -> 42.6 us.  DSMEM moves ~21 B/clk per SM; L2 is faster.
    using C = CfgW4T<NA>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = smem_base + C::STAGES * C::STAGE_BYTES;
    auto xfull_bar = [&](int s) { return bars + 8u * s; };                    // TMA: activations + nibbles landed
    auto wfull_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };      // converters: weight k-block is in TMEM
    auto empty_bar = [&](int s) { return bars + 8u * (2 * C::STAGES + s); };  // MMAs of the stage retired
    const uint32_t tfull_bar = bars + 8u * (3 * C::STAGES), tempty_bar = tfull_bar + 8u;
    const uint32_t tmem_slot = tfull_bar + 16u;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_m = (M + NA - 1) / NA, num_n = N / W4_BN;
    const int nk = (2 * K) / BKB;   // k-blocks of 64 elements
    const int gpr = K / 128;        // scale groups per output channel
    const int num_items = num_m * num_n * splitk;
    auto kb_begin = [&](int ks) { return (int)((int64_t)nk * ks / splitk); };
    pdl_trigger();

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(xfull_bar(s), 1);
            mbar_init(wfull_bar(s), 8);   // one elected lane per converter warp
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tfull_bar, 1);
        mbar_init(tempty_bar, 8);
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(C::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const uint32_t tmem_w = tmem_base + NA;   // operand ring: stage s at columns [NA + 32 s, NA + 32 s + 32)
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&map_x);
            tma_prefetch_desc(&map_w);
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                const int tile = item / splitk, ks = item - tile * splitk;
                const int m_blk = tile % num_m, n_blk = tile / num_m;
                for (int kb = kb_begin(ks); kb < kb_begin(ks + 1); ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t sx = smem_base + stage * C::STAGE_BYTES, sraw = sx + C::ACT_BYTES;
                    mbar_expect_tx(xfull_bar(stage), C::ACT_BYTES + C::RAW_BYTES);
#pragma unroll
                    for (int h = 0; h < NA / BM; ++h)   // rows past M read as zero
                        tma_load_2d(sx + h * BM * BKB, &map_x, xfull_bar(stage), kb * BKB, m_blk * NA + h * BM);
                    tma_load_2d(sraw, &map_w, xfull_bar(stage), kb * (BKB / 4), n_blk * W4_BN);
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc<false>(NA);   // M = 128 channels, N = NA rows
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
                const int ks = item % splitk;
                const int kb0 = kb_begin(ks), kb1 = kb_begin(ks + 1);
                mbar_wait(tempty_bar, (uint32_t)(it & 1) ^ 1);   // the previous item's epilogue has drained the accumulator
                tc_fence_after();
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(xfull_bar(stage), phase);
                    mbar_wait(wfull_bar(stage), phase);
                    tc_fence_after();
                    const uint64_t dx = make_smem_desc(smem_base + stage * C::STAGE_BYTES);
                    const uint32_t ta = tmem_w + stage * C::A_COLS;
#pragma unroll
                    for (int k = 0; k < BKB / 32; ++k)   // 16 fp16 of K per MMA = 8 operand columns, 32 B of the smem row
                        tc_mma_ts_f16(tmem_base, ta + 8 * k, dx + 2 * k, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
                    tc_commit(empty_bar(stage));
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(tfull_bar);
            }
        }
    } else {
        // thread = output channel r of the tile (its TMEM lane); hf = which 32 of the k-block's 64 elements it converts
        // and, in the epilogue, which half of the activation rows it stores
        const int quarter = warp & 3, r = quarter * 32 + lane, hf = (warp - 2) >> 2;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
            const int tile = item / splitk, ks = item - tile * splitk;
            const int m_blk = tile % num_m, n_blk = tile / num_m;
            const int kb0 = kb_begin(ks), kb1 = kb_begin(ks + 1);
            const __half* srow = w_scale + (int64_t)(n_blk * W4_BN + r) * gpr;
            // one fp16 scale per 128 elements = per 2 k-blocks; four per 8-byte load, the next four in flight
            const bool vec = (gpr & 3) == 0 && (reinterpret_cast<uintptr_t>(w_scale) & 7) == 0;
            auto load4 = [&](int grp) {
                if (vec) return __ldg(reinterpret_cast<const uint2*>(srow + 4 * grp));
                uint32_t h[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) h[i] = 4 * grp + i < gpr ? (uint32_t)__half_as_ushort(srow[4 * grp + i]) : 0u;
                return make_uint2(h[0] | (h[1] << 16), h[2] | (h[3] << 16));
            };
            int grp = kb0 >> 3;
            uint2 cur = load4(grp), nxt = cur;
            if (((grp + 1) << 3) < kb1) nxt = load4(grp + 1);
            for (int kb = kb0; kb < kb1; ++kb) {
                if ((kb >> 3) != grp) {
                    grp = kb >> 3;
                    cur = nxt;
                    if (((grp + 1) << 3) < kb1) nxt = load4(grp + 1);
                }
                const int j = (kb >> 1) & 3;
                const uint32_t word = (j & 2) ? cur.y : cur.x;
                const __half sc1 = __ushort_as_half((unsigned short)((j & 1) ? (word >> 16) : (word & 0xFFFFu)));
                const __half2 sc = __halves2half2(sc1, sc1);
                mbar_wait(xfull_bar(stage), phase);
                const uint32_t sraw = smem_base + stage * C::STAGE_BYTES + C::ACT_BYTES;
                uint4 raw;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w)
                             : "r"(sraw + r * 32 + hf * 16));
                const uint32_t wds[4] = {raw.x, raw.y, raw.z, raw.w};
                uint32_t v[16];   // elements [32 hf, 32 hf + 32) of channel r as 16 half2 (k, k + 1) = 16 operand columns
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint4 e = w4_expand8(wds[c], sc);
                    v[4 * c] = e.x; v[4 * c + 1] = e.y; v[4 * c + 2] = e.z; v[4 * c + 3] = e.w;
                }
                tmem_st16(tmem_w + lane_addr + stage * C::A_COLS + hf * 16, v);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(wfull_bar(stage));
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }

            // ---- epilogue of the item: lanes = channels n, accumulator columns = activation rows
            mbar_wait(tfull_bar, (uint32_t)(it & 1));
            tc_fence_after();
            const int n = n_blk * W4_BN + r;
            constexpr int HALF = NA / 2;
#pragma unroll 1
            for (int c = 0; c < HALF / 32; ++c) {
                const int col0 = hf * HALF + c * 32;
                const int m0 = m_blk * NA + col0;
                if (m0 >= M) break;                         // warp-uniform
                uint32_t rr[32];
                tmem_ld32(tmem_base + lane_addr + col0, rr);
                w4t_store32<EPI>(rr, m0, n, lane, M, N, out, ldc, splitk > 1 ? ws + (int64_t)ks * M * N : nullptr);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS));
    }
}

// out = epilogue(sum over slices of ws[slice][m][n]), 8 consecutive columns per thread (coalesced float4 loads)
template <int EPI>
__global__ void __launch_bounds__(256) w4_splitk_reduce_kernel(const float* __restrict__ ws, int splitk, int M, int N,
                                                               void* __restrict__ out, int64_t ldc) {
    pdl_trigger();
    pdl_wait();
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int n8 = N >> 3;
    if (idx >= (int64_t)M * n8) return;
    const int m = (int)(idx / n8), n0 = (int)(idx - (int64_t)m * n8) * 8;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
    for (int k = 0; k < splitk; ++k) {
        const float4* src = reinterpret_cast<const float4*>(ws + ((int64_t)k * M + m) * N + n0);
        const float4 a = __ldcg(src), b = __ldcg(src + 1);
        v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
        v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    if constexpr (EPI == EPI_SWIGLU) {
        __half* orow = reinterpret_cast<__half*>(out) + (int64_t)m * ldc + (n0 >> 1);
        const __half2 h0 = __floats2half2_rn(silu_mul_f32(v[0], v[1]), silu_mul_f32(v[2], v[3]));
        const __half2 h1 = __floats2half2_rn(silu_mul_f32(v[4], v[5]), silu_mul_f32(v[6], v[7]));
        *reinterpret_cast<uint2*>(orow) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
    } else {
        __half* orow = reinterpret_cast<__half*>(out) + (int64_t)m * ldc + n0;
        uint32_t w[4];
        uint4 old = make_uint4(0, 0, 0, 0);
        if constexpr (EPI == EPI_RESIDUAL) old = *reinterpret_cast<const uint4*>(orow);
        const uint32_t* ow = reinterpret_cast<const uint32_t*>(&old);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float a = v[2 * j], b = v[2 * j + 1];
            if constexpr (EPI == EPI_RESIDUAL) {
                const float2 o2 = __half22float2(*reinterpret_cast<const __half2*>(&ow[j]));
                a = __fadd_rn(o2.x, a);
                b = __fadd_rn(o2.y, b);
            }
            const __half2 h = __floats2half2_rn(a, b);
            w[j] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(orow) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// ---------------------------------------------------------------- host side
// 2D byte tensor [rows, Kb] with a (128 B x box_rows) box and 128 B swizzle; out-of-range rows read as zero
bool encode_map(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t Kb, uint32_t box_rows) {
    const uint64_t dims[2] = {Kb, rows};
    const uint64_t strides[1] = {Kb};
    const uint32_t box[2] = {(uint32_t)BKB, box_rows};
    return tma_encode_bytes(map, ptr, 2, dims, strides, box, true);
}

std::mutex g_map_mutex;
std::map<std::tuple<const void*, uint64_t, uint64_t, uint32_t>, CUtensorMap> g_map_cache;  // weights are long-lived

bool cached_map(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t Kb, uint32_t box_rows) {
    std::lock_guard<std::mutex> lk(g_map_mutex);
    auto key = std::make_tuple(ptr, rows, Kb, box_rows);
    auto it = g_map_cache.find(key);
    if (it != g_map_cache.end()) {
        *map = it->second;
        return true;
    }
    if (!encode_map(map, ptr, rows, Kb, box_rows)) return false;
    if (g_map_cache.size() > 4096) g_map_cache.clear();
    g_map_cache[key] = *map;
    return true;
}

// number of SMs the persistent GEMM grids may occupy (B2LLM_GEMM_SMS; default all).  A smaller budget leaves SMs to a
// kernel running concurrently on another stream (scripts/overlap_probe.py).
int gemm_sm_budget() {
    static int budget = [] {
        const char* e = getenv("B2LLM_GEMM_SMS");
        const int n = e ? atoi(e) : 0;
        const int sms = device_num_sms();
        return (n >= 2 && n <= sms) ? n : sms;
    }();
    return budget;
}

template <bool I8, int EPI, int BN>
int32_t launch(cudaStream_t s, const void* a, const float* a_scale, const void* w, const float* w_scale, int64_t M, int N,
               int Kb, void* out, int64_t ldc) {
    using C = Cfg<BN>;
    auto kern = gemm_tc_kernel<I8, EPI, BN>;
    B2_ENSURE_DYN_SMEM(kern, C::SMEM_BYTES);
    CUtensorMap ma, mw;
    B2_REQUIRE(cached_map(&ma, a, (uint64_t)M, (uint64_t)Kb, BM), B2LLM_ERR_DEVICE, "cuTensorMapEncodeTiled(A) failed");
    B2_REQUIRE(cached_map(&mw, w, (uint64_t)N, (uint64_t)Kb, BN), B2LLM_ERR_DEVICE, "cuTensorMapEncodeTiled(W) failed");
    const int tiles = (int)((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    const int sms = gemm_sm_budget();
    const int grid = tiles < sms ? tiles : sms;
    launch_kernel(kern, dim3(grid), dim3(NUM_THREADS), C::SMEM_BYTES, s, ma, mw, a_scale, w_scale, (int)M, N, Kb, out, ldc);
    B2_LAUNCH_CHECK();
    return B2LLM_OK;
}

template <bool I8, int EPI>
int32_t launch_pair(cudaStream_t s, const void* a, const float* a_scale, const void* w, const float* w_scale, int64_t M, int N,
                    int Kb, void* out, int64_t ldc) {
    using C = Cfg2;
    auto kern = gemm_tc2_kernel<I8, EPI>;
    B2_ENSURE_DYN_SMEM(kern, C::SMEM_BYTES);
    CUtensorMap ma, mw;
    B2_REQUIRE(cached_map(&ma, a, (uint64_t)M, (uint64_t)Kb, BM), B2LLM_ERR_DEVICE, "cuTensorMapEncodeTiled(A) failed");
    B2_REQUIRE(cached_map(&mw, w, (uint64_t)N, (uint64_t)Kb, BN2 / 2), B2LLM_ERR_DEVICE, "cuTensorMapEncodeTiled(W) failed");
    const int tiles = (int)((M + 2 * BM - 1) / (2 * BM)) * (N / BN2);
    const int max_pairs = gemm_sm_budget() / 2;
    const int pairs = tiles < max_pairs ? tiles : max_pairs;
    launch_kernel(kern, dim3(2 * pairs), dim3(NUM_THREADS), C::SMEM_BYTES, s, ma, mw, a_scale, w_scale, (int)M, N, Kb, out, ldc);
    B2_LAUNCH_CHECK();
    return B2LLM_OK;
}

int g_gemm_pair_mode = -1;  // B2LLM_GEMM_2CTA: 0 never, 1 auto (default), 2 whenever the shape allows
std::once_flag g_gemm_env_once;

template <bool I8, int EPI>
int32_t pick_bn(cudaStream_t s, const void* a, const float* a_scale, const void* w, const float* w_scale, int64_t M, int N,
                int Kb, void* out, int64_t ldc, int pair_mode) {
    // wave efficiency of the persistent schedule for both tile widths
    auto eff = [&](int bn) {
        const int64_t tiles = ((M + BM - 1) / BM) * ((N + bn - 1) / bn);
        const int g_num_sms = device_num_sms();
        const int64_t rounds = (tiles + g_num_sms - 1) / g_num_sms;
        const double useful = (double)M * N / ((double)((M + BM - 1) / BM * BM) * ((N + bn - 1) / bn * bn));
        return useful * (double)tiles / (double)(rounds * g_num_sms);
    };
    std::call_once(g_gemm_env_once, [] {
        const char* e = getenv("B2LLM_GEMM_2CTA");
        g_gemm_pair_mode = e ? atoi(e) : 1;
    });
    // CTA pairs: 256 x 256 tiles.  Worth it once there are at least two 128-row blocks of A to pair up.
    const bool pair_ok = N % BN2 == 0 && M > BM;
    const int mode = pair_mode >= 0 ? pair_mode : g_gemm_pair_mode;
    if (pair_ok && (mode == 2 || (mode == 1 && M >= 2 * BM)))
        return launch_pair<I8, EPI>(s, a, a_scale, w, w_scale, M, N, Kb, out, ldc);
    if (N >= 256 && eff(256) >= eff(128) * 0.97)
        return launch<I8, EPI, 256>(s, a, a_scale, w, w_scale, M, N, Kb, out, ldc);
    return launch<I8, EPI, 128>(s, a, a_scale, w, w_scale, M, N, Kb, out, ldc);
}

// split-K scratch of the W4A16 kernel, one per (device, stream): fp32 partial tiles + one counter per output tile
struct W4Scratch {
    float* ws = nullptr;
    size_t ws_floats = 0;
};
std::map<std::pair<int, cudaStream_t>, W4Scratch> g_w4_scratch;

int32_t w4_scratch(cudaStream_t s, size_t floats, W4Scratch* out) {
    int dev = 0;
    B2_CHECK_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_map_mutex);
    W4Scratch& sc = g_w4_scratch[{dev, s}];
    if (floats > sc.ws_floats) {
        if (sc.ws) {
            B2_CHECK_CUDA(cudaStreamSynchronize(s));
            cudaFree(sc.ws);
        }
        sc.ws = nullptr;
        sc.ws_floats = 0;
        B2_CHECK_CUDA(cudaMalloc(&sc.ws, floats * sizeof(float)));
        sc.ws_floats = floats;
    }
    *out = sc;
    return B2LLM_OK;
}

// packed-weight tensor map: [N, K/2] bytes, box {32 B, 128 rows}, no swizzle
int32_t w4_weight_map(CUtensorMap* mw, const uint8_t* packed, int N, int K) {
    std::lock_guard<std::mutex> lk(g_map_mutex);
    auto key = std::make_tuple((const void*)packed, (uint64_t)N, (uint64_t)(K / 2), (uint32_t)0xFFFF0004u);
    auto it = g_map_cache.find(key);
    if (it != g_map_cache.end()) {
        *mw = it->second;
        return B2LLM_OK;
    }
    const uint64_t dims[2] = {(uint64_t)(K / 2), (uint64_t)N};
    const uint64_t strides[1] = {(uint64_t)(K / 2)};
    const uint32_t box[2] = {(uint32_t)(BKB / 4), (uint32_t)W4_BN};
    B2_REQUIRE(tma_encode_bytes(mw, packed, 2, dims, strides, box, false), B2LLM_ERR_DEVICE, "cuTensorMapEncodeTiled(W4) failed");
    g_map_cache[key] = *mw;
    return B2LLM_OK;
}

template <int EPI, int MT>
int32_t launch_w4(cudaStream_t s, const void* a, const uint8_t* packed, const __half* scale, int64_t M, int N, int K, void* out,
                  int64_t ldc) {
    using C = CfgW4<MT>;
    auto kern = gemm_w4_kernel<EPI, MT>;
    B2_ENSURE_DYN_SMEM(kern, C::SMEM_BYTES);
    CUtensorMap ma, mw;
    B2_REQUIRE(cached_map(&ma, a, (uint64_t)M, (uint64_t)(2 * K), BM), B2LLM_ERR_DEVICE, "cuTensorMapEncodeTiled(A) failed");
    if (const int32_t rc = w4_weight_map(&mw, packed, N, K)) return rc;
    const int tiles = (int)((M + MT * BM - 1) / (MT * BM)) * (N / W4_BN);
    const int sms = gemm_sm_budget();
    // split-K when the tiles alone leave most of the machine idle: slices of >= 8 k-blocks, about one work item per SM
    const int nk = (2 * K) / BKB;
    int splitk = 1;
    if (tiles * 2 <= sms) splitk = std::max(1, std::min(sms / tiles, nk / 8));
    W4Scratch sc{};
    if (splitk > 1) {
        const int32_t rc = w4_scratch(s, (size_t)splitk * (size_t)M * (size_t)N, &sc);
        if (rc) return rc;
    }
    const int items = tiles * splitk;
    const int grid = items < sms ? items : sms;
    launch_kernel(kern, dim3(grid), dim3(W4_THREADS), C::SMEM_BYTES, s, ma, mw, scale, (int)M, N, K, out, ldc, splitk, sc.ws);
    B2_LAUNCH_CHECK();
    if (splitk > 1) {
        const int64_t threads = M * (N / 8);
        launch_kernel(w4_splitk_reduce_kernel<EPI>, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, s, (const float*)sc.ws, splitk,
                      (int)M, N, out, ldc);
        B2_LAUNCH_CHECK();
    }
    return B2LLM_OK;
}

// the transposed kernel (weights through TMEM): NA activation rows per tile
template <int EPI, int NA>
int32_t launch_w4t(cudaStream_t s, const void* a, const uint8_t* packed, const __half* scale, int64_t M, int N, int K, void* out,
                   int64_t ldc) {
    using C = CfgW4T<NA>;
    auto kern = gemm_w4t_kernel<EPI, NA>;
    B2_ENSURE_DYN_SMEM(kern, C::SMEM_BYTES);
    CUtensorMap mx, mw;
    B2_REQUIRE(cached_map(&mx, a, (uint64_t)M, (uint64_t)(2 * K), BM), B2LLM_ERR_DEVICE, "cuTensorMapEncodeTiled(X) failed");
    if (const int32_t rc = w4_weight_map(&mw, packed, N, K)) return rc;
    const int tiles = (int)((M + NA - 1) / NA) * (N / W4_BN);
    const int sms = gemm_sm_budget();
    // split-K when the tiles alone leave most of the machine idle: about one work item per SM, slices of >= 16 k-blocks (a
    // slice costs the fp32 scratch round trip and the reduce kernel: o of 70B at TP = 8, 16 k-blocks, is 14.0 us whole and
    // 14.6 us in two slices; down, 56 k-blocks, 27.4 vs 22.0 -- round 2 run 36)
    const int nk = (2 * K) / BKB;
    int splitk = 1;
    if (tiles * 2 <= sms) splitk = std::max(1, std::min(sms / tiles, nk / 16));
    static const int force_split = [] { const char* e = getenv("B2LLM_W4_SPLITK"); return e ? atoi(e) : 0; }();   // experiments
    if (force_split > 0) splitk = std::max(1, std::min(std::min(force_split, nk), std::max(1, sms / tiles)));
    const int items = tiles * splitk;
    W4Scratch sc{};
    if (splitk > 1) {
        const int32_t rc = w4_scratch(s, (size_t)splitk * (size_t)M * (size_t)N, &sc);
        if (rc) return rc;
    }
    launch_kernel(kern, dim3(items < sms ? items : sms), dim3(W4T_THREADS), C::SMEM_BYTES, s, mx, mw, scale, (int)M, N, K, out, ldc,
                  splitk, sc.ws);
    B2_LAUNCH_CHECK();
    if (splitk > 1) {
        const int64_t threads = M * (N / 8);
        launch_kernel(w4_splitk_reduce_kernel<EPI>, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, s, (const float*)sc.ws, splitk,
                      (int)M, N, out, ldc);
        B2_LAUNCH_CHECK();
    }
    return B2LLM_OK;
}

template <int EPI>
int32_t launch_w4t_pick(cudaStream_t s, const void* a, const uint8_t* packed, const __half* scale, int64_t M, int N, int K, void* out,
                        int64_t ldc) {
    return M > BM ? launch_w4t<EPI, 256>(s, a, packed, scale, M, N, K, out, ldc) : launch_w4t<EPI, 128>(s, a, packed, scale, M, N, K, out, ldc);
}

}  // namespace

bool gemm_tc_available() { return tma_available(); }

// fused W4A16 GEMM; B2LLM_ERR_UNSUPPORTED for shapes outside its envelope (caller falls back to dequant + fp16 GEMM)
int32_t launch_gemm_w4a16(cudaStream_t s, const void* a_fp16, const uint8_t* packed, const void* scale_fp16, int64_t M, int N, int K,
                          int epilogue, void* out, int64_t ldc) {
    if (!gemm_tc_available() || K % 128 != 0 || N % W4_BN != 0 || M >= (1ll << 31) || ((uintptr_t)a_fp16 & 15) ||
        ((uintptr_t)packed & 15) || ((uintptr_t)out & 15) || (ldc % 8) != 0 || epilogue < EPI_F16 || epilogue > EPI_SWIGLU) {
        set_last_error("tcgen05 w4a16 gemm: shape / alignment outside the kernel envelope");
        return B2LLM_ERR_UNSUPPORTED;
    }
    if (M == 0) return B2LLM_OK;
    const __half* sc = (const __half*)scale_fp16;
    // default: the transposed kernel (weight tile through TMEM).  B2LLM_W4_IMPL=ss keeps the both-operands-in-smem kernel
    // reachable for A/B measurements.
    static const bool use_ss = [] { const char* e = getenv("B2LLM_W4_IMPL"); return e && !strcmp(e, "ss"); }();
    if (!use_ss) {
        switch (epilogue) {
            case EPI_F16: return launch_w4t_pick<EPI_F16>(s, a_fp16, packed, sc, M, N, K, out, ldc);
            case EPI_RESIDUAL: return launch_w4t_pick<EPI_RESIDUAL>(s, a_fp16, packed, sc, M, N, K, out, ldc);
            default: return launch_w4t_pick<EPI_SWIGLU>(s, a_fp16, packed, sc, M, N, K, out, ldc);
        }
    }
    // two A tiles per weight tile halve the conversion work, but only pay once the grid still fills half the machine
    // two A tiles per weight tile (MT = 2) halve the conversion work per flop; split-K refills the machine when that leaves
    // few tiles.  It pays while a work item still runs a long K loop -- measured at M = 256 (round 2 run 17): gate_up
    // (64 k-blocks per item) 69.4 -> 52.7 us, down (28) 32.8 -> 31.1, but o (8) 14.7 -> 19.6 and qkv (9 vs 18) 20.5 -> 23.1.
    const int sms_w4 = gemm_sm_budget();
    const int nk_w4 = (2 * K) / BKB;
    const int tiles2 = (int)((M + 2 * BM - 1) / (2 * BM)) * (N / W4_BN);
    const int split2 = tiles2 * 2 <= sms_w4 ? std::max(1, std::min(sms_w4 / tiles2, nk_w4 / 8)) : 1;
    static const int force_mt = [] { const char* e = getenv("B2LLM_W4_MT"); return e ? atoi(e) : 0; }();  // 1 / 2: force (experiments)
    const bool two = M > BM && (force_mt == 2 || (force_mt != 1 && nk_w4 / split2 >= 24));
    switch (epilogue) {
        case EPI_F16: return two ? launch_w4<EPI_F16, 2>(s, a_fp16, packed, sc, M, N, K, out, ldc) : launch_w4<EPI_F16, 1>(s, a_fp16, packed, sc, M, N, K, out, ldc);
        case EPI_RESIDUAL: return two ? launch_w4<EPI_RESIDUAL, 2>(s, a_fp16, packed, sc, M, N, K, out, ldc) : launch_w4<EPI_RESIDUAL, 1>(s, a_fp16, packed, sc, M, N, K, out, ldc);
        default: return two ? launch_w4<EPI_SWIGLU, 2>(s, a_fp16, packed, sc, M, N, K, out, ldc) : launch_w4<EPI_SWIGLU, 1>(s, a_fp16, packed, sc, M, N, K, out, ldc);
    }
}

int32_t launch_gemm_tc(cudaStream_t s, bool is_i8, const void* a, const float* a_scale, const void* w, const float* w_scale,
                       int64_t M, int N, int K, int epilogue, void* out, int64_t ldc, int pair_mode) {
    if (!gemm_tc_available()) {
        set_last_error("tcgen05 gemm: needs an sm_100 device and cuTensorMapEncodeTiled");
        return B2LLM_ERR_UNSUPPORTED;
    }
    const int Kb = is_i8 ? K : 2 * K;
    // shapes outside the kernel's envelope go to the mma.sync baseline
    if (Kb % 16 != 0 || N % 32 != 0 || M >= (1ll << 31) || ((uintptr_t)a & 15) || ((uintptr_t)w & 15) ||
        ((uintptr_t)out & 15) || (ldc % 8) != 0) {
        set_last_error("tcgen05 gemm: shape / alignment outside the kernel envelope");
        return B2LLM_ERR_UNSUPPORTED;
    }
    if (M == 0) return B2LLM_OK;
    if (is_i8) {
        switch (epilogue) {
            case EPI_F16: return pick_bn<true, EPI_F16>(s, a, a_scale, w, w_scale, M, N, Kb, out, ldc, pair_mode);
            case EPI_RESIDUAL: return pick_bn<true, EPI_RESIDUAL>(s, a, a_scale, w, w_scale, M, N, Kb, out, ldc, pair_mode);
            case EPI_SWIGLU: return pick_bn<true, EPI_SWIGLU>(s, a, a_scale, w, w_scale, M, N, Kb, out, ldc, pair_mode);
            default: break;
        }
    } else {
        switch (epilogue) {
            case EPI_F16: return pick_bn<false, EPI_F16>(s, a, nullptr, w, nullptr, M, N, Kb, out, ldc, pair_mode);
            case EPI_RESIDUAL: return pick_bn<false, EPI_RESIDUAL>(s, a, nullptr, w, nullptr, M, N, Kb, out, ldc, pair_mode);
            case EPI_SWIGLU: return pick_bn<false, EPI_SWIGLU>(s, a, nullptr, w, nullptr, M, N, Kb, out, ldc, pair_mode);
            case EPI_F32: return pick_bn<false, EPI_F32>(s, a, nullptr, w, nullptr, M, N, Kb, out, ldc, pair_mode);
            default: break;
        }
    }
    set_last_error("tcgen05 gemm: unsupported epilogue");
    return B2LLM_ERR_UNSUPPORTED;
}

}  // namespace b2llm
