// K6 (+K7 reference form) of SURVEY.md section 2.3: attention over the paged / indexed int8 KV cache.
//
//  * attn_decode_kernel -- THE dominant kernel of the decode step (HBM-bound: it streams
//    kv_len * 320 B per (sequence, kv head)).  One CTA per (sequence, kv head [, q-head chunk], kv split);
//    every warp owns a private 3-stage ring of 16-token units (K 2 KB + V 2 KB + scales 1 KB) and
//    never synchronises with the other warps until the final merge.  Two loaders fill the ring:
//      - TMA (default): one lane issues four cp.async.bulk.tensor 3D loads per unit -- a 16-token x
//        128 B box of the int8 cache (128-byte swizzle) for K and V, and 16 x 32 B of fp16 scales for
//        each -- against tensor maps that describe any of the reference's four cache layouts;
//        completion through a per-stage mbarrier (complete_tx).  Needs the 16 tokens of a unit to be
//        contiguous slots: cache_mode 0, or paging with page_size % 16 == 0.
//      - cp.async (fallback for other page sizes): 10 x 16 B LDGSTS per lane per unit, same smem image.
//    int8 -> fp16 dequant happens in registers (magic-number trick: (b ^ 0x80) | 0x6400 is the half
//    1024 + (b + 128); minus 1152, times the group scale) directly into mma.sync m16n8k16 fragments:
//        S[q-head, token]  = Q[q-head, d]   . K^T[d, token]     (A = Q, B = K as stored: token-major)
//        O[q-head, d]      = P[q-head, tok] . V[tok, d]         (A = P from S's accumulator layout,
//                                                                as hi + lo fp16 halves)
//    the head-dim and token orders inside a fragment are permuted so that every lane reads whole
//    16-byte chunks, conflict-free under the 128 B swizzle; fp32 accumulation, online softmax in the
//    exp2 domain, split-KV partials merged by attn_merge_kernel.  GQA packs up to 8 q-heads of a kv
//    head into the MMA M dimension.
//  * attn_simple_kernel -- one warp per (token, q-head), straightforward fp32 math.  Reference form
//    used for prefill tokens (fresh fp16 K/V + optional cached prefix) and as the checker of the
//    MMA kernel (impl = 1).
//
// Numeric contract: oracle/llama_ref.py attention_decode / _attend.
#include <atomic>
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"
#include "tma_utils.cuh"

namespace b2llm {

namespace {

struct AttnParams {
    const __half* qkv;
    const int64_t* seq_starts;
    const int64_t* start_pos;
    const int64_t* cache_indices;
    int batch;
    int decoding_batches;
    int64_t max_pages;
    int nq, nkv, D;
    int cache_mode, page_size, group, cache_prefill;
    int kv16;              // fp16 cache without scales (cache_quant_bit 0 / group 1); strides stay in ELEMENTS
    const int8_t* cache;   // layer offset applied
    const __half* scale;   // layer offset applied
    KvStrides cs;
    float sm_scale;        // 1 / sqrt(D)
    __half* out;           // [T, nq * D]
    float* ws;             // split partials
    int nsplit;
    int64_t token_begin, token_end;
    unsigned long long* trace;  // debug (b2llm_debug_attention_trace): 4 words per CTA of the decode kernel, else nullptr
};

// ------------------------------------------------------------------------------------------
// simple kernel: warp per (token, q head)
template <int D>
__global__ void __launch_bounds__(128) attn_simple_kernel(AttnParams p) {
    constexpr int PER = D / 32;
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    const int64_t ntok = p.token_end - p.token_begin;
    if (w >= ntok * p.nq) return;
    const int64_t t = p.token_begin + w / p.nq;
    const int hq = (int)(w % p.nq);
    const int hk = hq / (p.nq / p.nkv);
    const int b = find_seq(p.seq_starts, p.batch, t);
    const int64_t sp = p.start_pos[b];
    const int64_t pos = sp + (t - p.seq_starts[b]);
    const bool decode = b < p.decoding_batches;
    const int64_t fresh_from = decode ? pos + 1 : sp;  // keys >= fresh_from come from this step's qkv
    const int heads = p.nq + 2 * p.nkv;

    float q[PER];
    const __half* qrow = p.qkv + t * (int64_t)heads * D + (int64_t)hq * D;
#pragma unroll
    for (int i = 0; i < PER; ++i) q[i] = __half2float(qrow[lane * PER + i]);

    float m = -INFINITY, l = 0.f, o[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) o[i] = 0.f;
    const int g = (lane * PER) / 8;  // scale group of this lane's dims (PER <= 8 and divides 8)

    for (int64_t j = 0; j <= pos; ++j) {
        float kx[PER], vx[PER];
        if (j < fresh_from) {
            const int64_t slot = kv_slot(p.cache_indices, p.cache_mode, p.page_size, p.max_pages, b, j);
            const int64_t off = hk * p.cs.head + slot * p.cs.tok;
            if (p.kv16) {
                const __half* kr = reinterpret_cast<const __half*>(p.cache) + off;
                const __half* vr = kr + p.cs.kv;
#pragma unroll
                for (int i = 0; i < PER; ++i) {
                    kx[i] = __half2float(kr[lane * PER + i]);
                    vx[i] = __half2float(vr[lane * PER + i]);
                }
            } else {
                const int8_t* kr = p.cache + off;
                const int8_t* vr = p.cache + p.cs.kv + off;
                const float ks = __half2float(p.scale[off / p.group + g]);
                const float vs = __half2float(p.scale[(p.cs.kv + off) / p.group + g]);
#pragma unroll
                for (int i = 0; i < PER; ++i) {
                    kx[i] = __half2float(__float2half_rn(__fmul_rn((float)kr[lane * PER + i], ks)));
                    vx[i] = __half2float(__float2half_rn(__fmul_rn((float)vr[lane * PER + i], vs)));
                }
            }
        } else {
            const int64_t tj = p.seq_starts[b] + (j - sp);
            const __half* kr = p.qkv + tj * (int64_t)heads * D + (int64_t)(p.nq + hk) * D;
            const __half* vr = kr + (int64_t)p.nkv * D;
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                kx[i] = __half2float(kr[lane * PER + i]);
                vx[i] = __half2float(vr[lane * PER + i]);
            }
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) s += q[i] * kx[i];
        s = warp_sum(s) * p.sm_scale;
        const float mn = fmaxf(m, s);
        const float corr = __expf(m - mn), pj = __expf(s - mn);
        l = l * corr + pj;
#pragma unroll
        for (int i = 0; i < PER; ++i) o[i] = o[i] * corr + pj * vx[i];
        m = mn;
    }
    __half* orow = p.out + t * (int64_t)p.nq * D + (int64_t)hq * D;
    const float inv = 1.f / l;
#pragma unroll
    for (int i = 0; i < PER; ++i) orow[lane * PER + i] = __float2half_rn(o[i] * inv);
}

// ------------------------------------------------------------------------------------------
// tensor-core split-KV decode kernel (D = 128, int8 group-8 cache)
constexpr int UNIT = 16;                 // tokens per warp iteration
#ifndef B2_ATTN_NSTAGE
#define B2_ATTN_NSTAGE 3
#endif
// 4 stages (3 units in flight, but only 10 one-warp CTAs per SM) were measured against 3 x 12 (round 2 run 43): the headline
// shape slows from 0.7745 to 0.7990 ms per launch, 13B / TP 4 from 0.650 to 0.719; resident warps beat deeper rings.
constexpr int NSTAGE = B2_ATTN_NSTAGE;            // ring stages per warp: NSTAGE - 1 units in flight while one is consumed
constexpr int kCtasPerSm = NSTAGE >= 4 ? 10 : 12; // one-warp CTAs per SM the ring leaves room for (int8 cache: 5 KB per stage)
constexpr int K_BYTES = UNIT * 128;      // 2048
constexpr int S_BYTES = UNIT * 32;       // 512 (16 fp16 scales per token)
constexpr int STAGE = 2 * K_BYTES + 2 * S_BYTES;  // 5120: K | V | K scales | V scales
constexpr int STAGE16 = 4 * K_BYTES;     // fp16 cache: K d[0,64) | V d[0,64) | K d[64,128) | V d[64,128), 16 tokens x 128 B each
constexpr int MAX_WARPS = 4;

struct DecodeTma {          // coordinates of the TMA loader (see make_kv_maps)
    int tok_dim;            // 1: tokens run along tensor dim 1 (layout 3), 2: along dim 2 (layouts 0..2)
    int k_fixed, v_fixed;   // the non-token coordinate for K / V of head 0 (+ head index in the kernel)
    int k_tok0, v_tok0;     // token coordinate base for K / V
};

__device__ __forceinline__ uint32_t lop3_and_xor(uint32_t x, uint32_t m, uint32_t k) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x6a;\n" : "=r"(r) : "r"(x), "r"(m), "r"(k));  // (x & m) ^ k
    return r;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;\n" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}
// two biased bytes (already in half2 lanes as 0x64xx) -> (x - 1152) * scale, packed half2
__device__ __forceinline__ uint32_t deq2(uint32_t e, __half2 sc) {
    const __half2 bias = __halves2half2(__ushort_as_half(0x6480), __ushort_as_half(0x6480));  // 1152.0
    __half2 h = *reinterpret_cast<__half2*>(&e);
    h = __hmul2(__hsub2(h, bias), sc);
    return *reinterpret_cast<uint32_t*>(&h);
}

// smem image of a 16-token x 128 B tile: 16-byte chunk c of row r lives at chunk c ^ (r & 7)
// (exactly what TMA's 128-byte swizzle produces; the cp.async loader writes the same image)
__device__ __forceinline__ uint32_t tile_off(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

struct WarpState {
    float o[8][4];          // O^T fragments: m-tile i holds d = 16 g + i (regs 0,1) and 16 g + 8 + i (regs 2,3) x heads 2t, 2t+1
    float m_run[2], l_run[2];  // running max / per-lane partial sum for heads 2t and 2t + 1
};

__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t a) {
    uint32_t d;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;\n" : "=r"(d) : "r"(a));
    return d;
}
__device__ __forceinline__ void mma_f16_full(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                             uint32_t b1) {
    mma_f16_16816(d, a0, a1, a2, a3, b0, b1);
}

// one 16-token unit in the "transposed" formulation (tokens / head-dims fill the MMA M dimension, the
// q heads of the kv head the N dimension):
//     S^T[token, head] = K[token, d] . Q^T[d, head]          A = dequantised K rows (all 4 fragment regs live)
//     O^T[d, head]    += V^T[d, token] . P^T[token, head]    A = dequantised V (token pairs gathered by PRMT),
//                                                            B = P moved from accumulator to operand layout
//                                                                by movmatrix.trans, as hi + lo fp16 halves
// ex2.approx.ftz: one MUFU instead of exp2f's range check + two scalings (results below 2^-126 flush to zero)
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
    return y;
}
template <bool SLIM>
__device__ __forceinline__ float exp2_sel(float x) {
    if constexpr (SLIM) return ex2_approx(x);
    else return exp2f(x);
}

// online softmax of one unit's scores + P in the B-operand layout (shared by the int8 and the fp16 cache paths).
// s_acc: S^T accumulator (rows = tokens g, g + 8; cols = heads 2t, 2t + 1).  Out: P^T fragments as hi + lo fp16 halves.
template <bool SLIM>
__device__ __forceinline__ void softmax_unit(const float (&s_acc)[4], WarpState& st, int tbase, int kv_len, float sl2, int g,
                                             uint32_t& bh0, uint32_t& bh1, uint32_t& bl0, uint32_t& bl1) {
    const bool ok0 = tbase + g < kv_len, ok1 = tbase + g + 8 < kv_len;
    float sv[2][2], pv[2][2];
    bool moved = false;
    float m_safe[2], m_new[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        sv[c][0] = ok0 ? s_acc[c] * sl2 : -INFINITY;
        sv[c][1] = ok1 ? s_acc[2 + c] * sl2 : -INFINITY;
        float mx = fmaxf(sv[c][0], sv[c][1]);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
        m_new[c] = fmaxf(st.m_run[c], mx);
        moved |= m_new[c] != st.m_run[c];
        m_safe[c] = m_new[c] == -INFINITY ? 0.f : m_new[c];
        pv[c][0] = exp2_sel<SLIM>(sv[c][0] - m_safe[c]);
        pv[c][1] = exp2_sel<SLIM>(sv[c][1] - m_safe[c]);
    }
    if (__any_sync(0xffffffffu, moved)) {  // some running max moved: rescale (rare after the first units)
        const float c0 = exp2_sel<SLIM>(st.m_run[0] - m_safe[0]), c1 = exp2_sel<SLIM>(st.m_run[1] - m_safe[1]);
        st.l_run[0] *= c0;
        st.l_run[1] *= c1;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            st.o[i][0] *= c0; st.o[i][1] *= c1; st.o[i][2] *= c0; st.o[i][3] *= c1;
        }
        st.m_run[0] = m_new[0];
        st.m_run[1] = m_new[1];
    }
    st.l_run[0] += pv[0][0] + pv[0][1];
    st.l_run[1] += pv[1][0] + pv[1][1];

    // P: accumulator layout (row = token, cols = heads 2t, 2t+1) -> B operand layout (row = head g, cols =
    // tokens 2t, 2t+1) with movmatrix.trans; hi + lo fp16 halves keep ~22 bits of P
    __half2 ph0 = __floats2half2_rn(pv[0][0], pv[1][0]), ph1 = __floats2half2_rn(pv[0][1], pv[1][1]);
    const float2 f0 = __half22float2(ph0), f1 = __half22float2(ph1);
    __half2 pl0 = __floats2half2_rn(pv[0][0] - f0.x, pv[1][0] - f0.y), pl1 = __floats2half2_rn(pv[0][1] - f1.x, pv[1][1] - f1.y);
    bh0 = movmatrix_trans(*reinterpret_cast<uint32_t*>(&ph0));
    bh1 = movmatrix_trans(*reinterpret_cast<uint32_t*>(&ph1));
    bl0 = movmatrix_trans(*reinterpret_cast<uint32_t*>(&pl0));
    bl1 = movmatrix_trans(*reinterpret_cast<uint32_t*>(&pl1));
}

template <bool SLIM>
__device__ __forceinline__ void process_unit(const uint8_t* __restrict__ stage, const uint32_t (&qb)[8][2], WarpState& st,
                                             int tbase, int kv_len, float sl2, int g, int t) {
    const uint8_t* sK = stage;
    const uint8_t* sV = stage + K_BYTES;
    const uint8_t* sKS = stage + 2 * K_BYTES;
    const uint8_t* sVS = sKS + S_BYTES;

    // ---- S^T: lane (g, t) dequantises bytes [32 t, 32 t + 32) of token rows g and g + 8
    float s_acc[4] = {0.f, 0.f, 0.f, 0.f};
    {
        const uint4 ca = *reinterpret_cast<const uint4*>(sK + tile_off(g, 2 * t));
        const uint4 cb = *reinterpret_cast<const uint4*>(sK + tile_off(g, 2 * t + 1));
        const uint4 cc = *reinterpret_cast<const uint4*>(sK + tile_off(g + 8, 2 * t));
        const uint4 cd = *reinterpret_cast<const uint4*>(sK + tile_off(g + 8, 2 * t + 1));
        const uint2 sc_lo = *reinterpret_cast<const uint2*>(sKS + g * 32 + 8 * t);        // groups 4t .. 4t+3 of token g
        const uint2 sc_hi = *reinterpret_cast<const uint2*>(sKS + (g + 8) * 32 + 8 * t);  // ... of token g + 8
        const uint32_t w_lo[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
        const uint32_t w_hi[8] = {cc.x, cc.y, cc.z, cc.w, cd.x, cd.y, cd.z, cd.w};
        const __half2 l01 = *reinterpret_cast<const __half2*>(&sc_lo.x), l23 = *reinterpret_cast<const __half2*>(&sc_lo.y);
        const __half2 h01 = *reinterpret_cast<const __half2*>(&sc_hi.x), h23 = *reinterpret_cast<const __half2*>(&sc_hi.y);
        const __half2 s_lo[4] = {__low2half2(l01), __high2half2(l01), __low2half2(l23), __high2half2(l23)};
        const __half2 s_hi[4] = {__low2half2(h01), __high2half2(h01), __low2half2(h23), __high2half2(h23)};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t a0 = deq2(lop3_and_xor(w_lo[j], 0x00FF00FFu, 0x64806480u), s_lo[j >> 1]);       // token g, bytes 0,2
            const uint32_t a2 = deq2(lop3_and_xor(w_lo[j] >> 8, 0x00FF00FFu, 0x64806480u), s_lo[j >> 1]);  // token g, bytes 1,3
            const uint32_t a1 = deq2(lop3_and_xor(w_hi[j], 0x00FF00FFu, 0x64806480u), s_hi[j >> 1]);       // token g + 8
            const uint32_t a3 = deq2(lop3_and_xor(w_hi[j] >> 8, 0x00FF00FFu, 0x64806480u), s_hi[j >> 1]);
            mma_f16_full(s_acc, a0, a1, a2, a3, qb[j][0], qb[j][1]);
        }
    }

    // ---- online softmax per head column c (heads 2t + c) over the 16 tokens (rows g and g + 8 of all lanes)
    uint32_t bh0, bh1, bl0, bl1;
    softmax_unit<SLIM>(s_acc, st, tbase, kv_len, sl2, g, bh0, bh1, bl0, bl1);

    // ---- O^T += V^T P^T: lane (g, t) dequantises bytes [16 g, 16 g + 16) of token rows {2t, 2t+1, 8+2t, 9+2t}
    const int r0 = 2 * t, r1 = 2 * t + 1, r2 = 8 + 2 * t, r3 = 9 + 2 * t;
    const uint4 va = *reinterpret_cast<const uint4*>(sV + tile_off(r0, g));
    const uint4 vb = *reinterpret_cast<const uint4*>(sV + tile_off(r1, g));
    const uint4 vc = *reinterpret_cast<const uint4*>(sV + tile_off(r2, g));
    const uint4 vd = *reinterpret_cast<const uint4*>(sV + tile_off(r3, g));
    const uint32_t sA = *reinterpret_cast<const uint32_t*>(sVS + r0 * 32 + 4 * g);  // groups 2g, 2g+1
    const uint32_t sB = *reinterpret_cast<const uint32_t*>(sVS + r1 * 32 + 4 * g);
    const uint32_t sC = *reinterpret_cast<const uint32_t*>(sVS + r2 * 32 + 4 * g);
    const uint32_t sD = *reinterpret_cast<const uint32_t*>(sVS + r3 * 32 + 4 * g);
    uint32_t u_ab[2] = {prmt(sA, sB, 0x5410), prmt(sA, sB, 0x7632)};  // (A, B) scales of group 2g / 2g+1
    uint32_t u_cd[2] = {prmt(sC, sD, 0x5410), prmt(sC, sD, 0x7632)};
    const __half2 s_ab0 = *reinterpret_cast<__half2*>(&u_ab[0]), s_ab1 = *reinterpret_cast<__half2*>(&u_ab[1]);
    const __half2 s_cd0 = *reinterpret_cast<__half2*>(&u_cd[0]), s_cd1 = *reinterpret_cast<__half2*>(&u_cd[1]);
    const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
    const uint32_t wc[4] = {vc.x, vc.y, vc.z, vc.w}, wdd[4] = {vd.x, vd.y, vd.z, vd.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int w = i >> 2, bsel = i & 3;
        const uint32_t sel = (uint32_t)(bsel | (bsel << 4) | ((4 + bsel) << 8) | ((4 + bsel) << 12));
        const uint32_t a0 = deq2(lop3_and_xor(prmt(wa[w], wb[w], sel), 0x00FF00FFu, 0x64806480u), s_ab0);            // d = 16g + i
        const uint32_t a1 = deq2(lop3_and_xor(prmt(wa[w + 2], wb[w + 2], sel), 0x00FF00FFu, 0x64806480u), s_ab1);    // d = 16g + 8 + i
        const uint32_t a2 = deq2(lop3_and_xor(prmt(wc[w], wdd[w], sel), 0x00FF00FFu, 0x64806480u), s_cd0);
        const uint32_t a3 = deq2(lop3_and_xor(prmt(wc[w + 2], wdd[w + 2], sel), 0x00FF00FFu, 0x64806480u), s_cd1);
        mma_f16_full(st.o[i], a0, a1, a2, a3, bh0, bh1);
        mma_f16_full(st.o[i], a0, a1, a2, a3, bl0, bl1);
    }
}

// fp16 cache (cache_quant_bit 0): the stage holds four 16-token x 128 B tiles (128 B swizzle): K d[0,64), V d[0,64),
// K d[64,128), V d[64,128).  No dequantisation: ldmatrix delivers the MMA operands as they lie --
//     S^T[token, head] = K[token, d] . Q^T[d, head]          A = K rows, natural d order (ldmatrix.x4)
//     O^T[d, head]    += V^T[d, token] . P^T[token, head]    A = V^T (ldmatrix.x4.trans), m-tile i = d [16 i, 16 i + 16)
// so accumulator o[i] holds d = 16 i + g (regs 0, 1) and 16 i + 8 + g (regs 2, 3) -- see acc_dim().
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
template <bool SLIM>
__device__ __forceinline__ void process_unit16(uint32_t stage, const uint32_t (&qb)[8][2], WarpState& st, int tbase, int kv_len,
                                               float sl2, int g, int lane) {
    const int mi = lane >> 3, lr = lane & 7;
    float s_acc[4] = {0.f, 0.f, 0.f, 0.f};
    {   // matrices of ldmatrix.x4: (tokens 0-7, chunk c), (tokens 8-15, c), (tokens 0-7, c + 1), (tokens 8-15, c + 1) = a0..a3
        const int r = lr + (mi & 1) * 8, cofs = mi >> 1;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t tile = stage + (j < 4 ? 0 : 2 * K_BYTES);
            uint32_t a0, a1, a2, a3;
            ldmatrix_x4(a0, a1, a2, a3, tile + tile_off(r, (j & 3) * 2 + cofs));
            mma_f16_full(s_acc, a0, a1, a2, a3, qb[j][0], qb[j][1]);
        }
    }
    uint32_t bh0, bh1, bl0, bl1;
    softmax_unit<SLIM>(s_acc, st, tbase, kv_len, sl2, g, bh0, bh1, bl0, bl1);
    {   // .trans matrices: (tokens 0-7, chunk c) -> a0, (tokens 0-7, c + 1) -> a1, (tokens 8-15, c) -> a2, (tokens 8-15, c + 1) -> a3
        const int r = lr + (mi >> 1) * 8, cofs = mi & 1;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t tile = stage + K_BYTES + (i < 4 ? 0 : 2 * K_BYTES);
            uint32_t a0, a1, a2, a3;
            ldmatrix_x4_trans(a0, a1, a2, a3, tile + tile_off(r, (i & 3) * 2 + cofs));
            mma_f16_full(st.o[i], a0, a1, a2, a3, bh0, bh1);
            mma_f16_full(st.o[i], a0, a1, a2, a3, bl0, bl1);
        }
    }
}

// head-dim index held by accumulator register pair `hi` (0: regs 0,1; 1: regs 2,3) of m-tile i in lane group g
template <bool KV16>
__device__ __forceinline__ int acc_dim(int i, int g, int hi) {
    return KV16 ? 16 * i + 8 * hi + g : 16 * g + 8 * hi + i;
}

// G: q heads per CTA (rows of the MMA M dimension in use), 1..8.  LOADER: 0 cp.async, 1 TMA, 2 TMA "slim" -- same
// data path and arithmetic as 1 with fewer instructions per unit around it: the page table is walked incrementally
// (no integer division per unit in the issuing lane) and exp2 is a bare ex2.approx.  The kernel's time follows the
// SM clock (in-step 1837 MHz: 0.896 ms, alone 1965 MHz: 0.838 ms), i.e. it is issue-bound before it is HBM-bound.
// LOADER 3 (default for cache layouts 2 / 3 since round 2 run 1): slim + K and V of a unit in ONE 4-D TMA box
// {row bytes, 16 tokens, 1 head, 2 (k, v)} and their scales in another -- 2 bulk loads per unit instead of 4
// (cache layouts 2 and 3, where k / v is an outer dimension; the smem image is unchanged: K | V | K scales | V scales).
// KV16: fp16 cache without scales (cache_quant_bit 0): 8 KB stages of four 128 B-swizzled tiles, ldmatrix operands
// (process_unit16); same ring, walk, softmax, split and merge code.
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// per-CTA timeline of the decode kernel (debug aid, off unless b2llm_debug_attention_trace armed it):
// {start ns, main loop entered ns, end ns, SM id} at trace[4 * linear CTA id]
__device__ __forceinline__ void trace_cta(unsigned long long* trace, unsigned long long t0, unsigned long long t1) {
    if (trace != nullptr && threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        unsigned long long* w = trace + 4ull * (blockIdx.x + (unsigned long long)gridDim.x * (blockIdx.y + (unsigned long long)gridDim.y * blockIdx.z));
        w[0] = t0;
        w[1] = t1;
        w[2] = global_timer_ns();
        w[3] = smid;
    }
}

template <int G, int WARPS, int LOADER, bool KV16>
__global__ void __launch_bounds__(WARPS * 32, kCtasPerSm / WARPS)
    attn_decode_kernel(AttnParams p, const __grid_constant__ CUtensorMap map_kv, const __grid_constant__ CUtensorMap map_sc,
                       DecodeTma tc) {
    constexpr bool TMA = LOADER != 0, SLIM = LOADER >= 2, MERGED = LOADER == 3;
    constexpr int STG = KV16 ? STAGE16 : STAGE;  // bytes per ring stage
    extern __shared__ uint8_t smem_raw[];
    const unsigned long long tr0 = p.trace ? global_timer_ns() : 0ull;
    pdl_trigger();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int split = blockIdx.x, b = blockIdx.z;
    const int gq = p.nq / p.nkv;                 // q heads per kv head
    const int chunks = (gq + G - 1) / G;          // CTAs per kv head along y
    const int hk = blockIdx.y / chunks;
    const int hq0 = hk * gq + (blockIdx.y % chunks) * G;
    const int nrow = min(G, hk * gq + gq - hq0);  // valid q-head rows

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // swizzle atoms need 1024 B alignment
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bars = smem_base + WARPS * NSTAGE * STG;

    if constexpr (TMA) {
        if (threadIdx.x == 0) {
            for (int i = 0; i < WARPS * NSTAGE; ++i) mbar_init(bars + 8u * i, 1);
            mbar_fence_init();
            tma_prefetch_desc(&map_kv);
            tma_prefetch_desc(&map_sc);
        }
        __syncthreads();
    }
    pdl_wait();  // everything above is local to the CTA; q, start_pos and this step's K / V rows come from earlier kernels

    const int kv_len = (int)(p.start_pos[b] + 1);
    const int units_total = (kv_len + UNIT - 1) / UNIT;
    const int units_per_split = (units_total + p.nsplit - 1) / p.nsplit;
    const int u0 = split * units_per_split;
    const int u1 = min(units_total, u0 + units_per_split);

    const int heads = p.nq + 2 * p.nkv;
    const int64_t tok = b;  // decode sequences come first, one token each (seq_starts[b] == b)

    // ---- Q^T fragments (B operand, column n = g is q head hq0 + g; columns >= nrow are zero).  k-slot
    // order follows the K byte order: step j, lane t covers d = 32 t + 4 j + {0,2} (b0) and {1,3} (b1)
    uint32_t qa[8][2];
    {
        const __half* qrow = p.qkv + tok * (int64_t)heads * 128 + (int64_t)(hq0 + g) * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int base = 32 * t + 4 * j;
            if (g < nrow && KV16) {  // natural k order: step j covers d = [16 j, 16 j + 16)
                qa[j][0] = *reinterpret_cast<const uint32_t*>(qrow + 16 * j + 2 * t);
                qa[j][1] = *reinterpret_cast<const uint32_t*>(qrow + 16 * j + 8 + 2 * t);
            } else if (g < nrow) {
                const uint2 v = *reinterpret_cast<const uint2*>(qrow + base);  // halves d..d+3
                qa[j][0] = prmt(v.x, v.y, 0x5410);  // (d+0, d+2)
                qa[j][1] = prmt(v.x, v.y, 0x7632);  // (d+1, d+3)
            } else {
                qa[j][0] = 0u;
                qa[j][1] = 0u;
            }
        }
    }

    uint8_t* wsm = smem + warp * (NSTAGE * STG);
    const uint32_t wsm_u32 = smem_base + warp * (NSTAGE * STG);
    const uint32_t wbar = bars + 8u * (warp * NSTAGE);

    // page-table walk: first cache slot of unit u (its 16 tokens are contiguous slots on the TMA path)
    auto unit_slot0 = [&](int u) -> int64_t {
        const int pos = u * UNIT;
        if (p.cache_mode == 0) return p.cache_indices[b] + pos;
        return p.cache_indices[(int64_t)b * p.max_pages + pos / p.page_size] + pos % p.page_size;
    };

    // SLIM: position of the next unit to issue as (page-table entry, token offset inside the page), advanced by
    // WARPS units per issue -- the divisions happen once here instead of once per unit
    int walk_page = 0, walk_off = 0, adv_page = 0, adv_off = 0;
    if constexpr (SLIM) {
        const int pos0 = (u0 + warp) * UNIT;
        if (p.cache_mode == 0) {
            walk_off = pos0;            // slot = cache_indices[b] + position: one "page" of unbounded size
            adv_off = WARPS * UNIT;
        } else {
            walk_page = pos0 / p.page_size;
            walk_off = pos0 % p.page_size;
            adv_page = (WARPS * UNIT) / p.page_size;
            adv_off = (WARPS * UNIT) % p.page_size;
        }
    }
    auto walk_slot0 = [&]() -> int64_t {  // first cache slot of the unit the walk points at
        if (p.cache_mode == 0) return p.cache_indices[b] + walk_off;
        return p.cache_indices[(int64_t)b * p.max_pages + walk_page] + walk_off;
    };
    auto walk_advance = [&]() {
        walk_page += adv_page;
        walk_off += adv_off;
        if (p.cache_mode != 0 && walk_off >= p.page_size) {
            walk_off -= p.page_size;
            ++walk_page;
        }
    };

    // ---- loader: fill stage `st` of this warp's ring with unit u
    constexpr int ESZ = KV16 ? 2 : 1;  // bytes per cache element
    const int8_t* kbase = p.cache + hk * p.cs.head * ESZ;
    const int8_t* vbase = kbase + p.cs.kv * ESZ;
    const __half* ksbase = KV16 ? nullptr : p.scale + hk * p.cs.head / 8;
    const __half* vsbase = KV16 ? nullptr : ksbase + p.cs.kv / 8;
    auto load_unit = [&](int u, int st) {
        const uint32_t sK = wsm_u32 + st * STG, sV = sK + K_BYTES, sKS = sV + K_BYTES, sVS = sKS + S_BYTES;
        if constexpr (TMA && KV16) {
            // tiles K lo | V lo | K hi | V hi: the box is 128 B wide (one swizzle span), a row of the fp16 cache 256 B
            if (lane == 0) {
                const int s0 = SLIM ? (int)walk_slot0() : (int)unit_slot0(u);
                const uint32_t bar = wbar + 8u * st;
                mbar_expect_tx(bar, STAGE16);
                if constexpr (MERGED) {  // {128 B, 16 tokens, 1 head, (k, v)}: K and V halves in one load
                    if (tc.tok_dim == 1) {
                        tma_load_4d(sK, &map_kv, bar, 0, s0, hk, tc.k_fixed);
                        tma_load_4d(sK + 2 * K_BYTES, &map_kv, bar, 128, s0, hk, tc.k_fixed);
                    } else {
                        tma_load_4d(sK, &map_kv, bar, 0, hk, s0, tc.k_fixed);
                        tma_load_4d(sK + 2 * K_BYTES, &map_kv, bar, 128, hk, s0, tc.k_fixed);
                    }
                } else {
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        const uint32_t dk = sK + hf * 2 * K_BYTES, dv = dk + K_BYTES;
                        if (tc.tok_dim == 1) {
                            tma_load_3d(dk, &map_kv, bar, 128 * hf, tc.k_tok0 + s0, tc.k_fixed + hk);
                            tma_load_3d(dv, &map_kv, bar, 128 * hf, tc.v_tok0 + s0, tc.v_fixed + hk);
                        } else {
                            tma_load_3d(dk, &map_kv, bar, 128 * hf, tc.k_fixed + hk, tc.k_tok0 + s0);
                            tma_load_3d(dv, &map_kv, bar, 128 * hf, tc.v_fixed + hk, tc.v_tok0 + s0);
                        }
                    }
                }
            }
        } else if constexpr (KV16) {
            // cp.async: 16 tokens x (16 K chunks + 16 V chunks) of 16 B; lane = chunk column, one token row per pass
            int64_t myslot = -1;
            {
                const int pos = u * UNIT + (lane & 15);
                if (pos < kv_len) myslot = kv_slot(p.cache_indices, p.cache_mode, p.page_size, p.max_pages, b, pos);
            }
            const int kvsel = lane >> 4, c = lane & 15;  // lanes 0-15: K chunks, 16-31: V chunks
            const uint32_t dst_tile = sK + kvsel * K_BYTES + (c >> 3) * 2 * K_BYTES;
            const int8_t* src_base = (kvsel ? vbase : kbase) + c * 16;
#pragma unroll 4
            for (int r = 0; r < UNIT; ++r) {
                const int64_t slot = __shfl_sync(0xffffffffu, myslot, r);
                const int ok = slot >= 0 ? 16 : 0;
                cp_async16(dst_tile + tile_off(r, c & 7), src_base + (slot >= 0 ? slot : 0) * p.cs.tok * 2, ok);
            }
        } else if constexpr (TMA) {
            if (lane == 0) {
                const int s0 = SLIM ? (int)walk_slot0() : (int)unit_slot0(u);
                const uint32_t bar = wbar + 8u * st;
                mbar_expect_tx(bar, STAGE);
                if constexpr (MERGED) {
                    if (tc.tok_dim == 1) {
                        tma_load_4d(sK, &map_kv, bar, 0, s0, hk, tc.k_fixed);
                        tma_load_4d(sKS, &map_sc, bar, 0, s0, hk, tc.k_fixed);
                    } else {
                        tma_load_4d(sK, &map_kv, bar, 0, hk, s0, tc.k_fixed);
                        tma_load_4d(sKS, &map_sc, bar, 0, hk, s0, tc.k_fixed);
                    }
                } else if (tc.tok_dim == 1) {
                    tma_load_3d(sK, &map_kv, bar, 0, tc.k_tok0 + s0, tc.k_fixed + hk);
                    tma_load_3d(sV, &map_kv, bar, 0, tc.v_tok0 + s0, tc.v_fixed + hk);
                    tma_load_3d(sKS, &map_sc, bar, 0, tc.k_tok0 + s0, tc.k_fixed + hk);
                    tma_load_3d(sVS, &map_sc, bar, 0, tc.v_tok0 + s0, tc.v_fixed + hk);
                } else {
                    tma_load_3d(sK, &map_kv, bar, 0, tc.k_fixed + hk, tc.k_tok0 + s0);
                    tma_load_3d(sV, &map_kv, bar, 0, tc.v_fixed + hk, tc.v_tok0 + s0);
                    tma_load_3d(sKS, &map_sc, bar, 0, tc.k_fixed + hk, tc.k_tok0 + s0);
                    tma_load_3d(sVS, &map_sc, bar, 0, tc.v_fixed + hk, tc.v_tok0 + s0);
                }
            }
        } else {
            // lanes 0..15 look up the slot of token u*16 + lane
            int64_t myslot = -1;
            {
                const int pos = u * UNIT + (lane & 15);
                if (pos < kv_len) myslot = kv_slot(p.cache_indices, p.cache_mode, p.page_size, p.max_pages, b, pos);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = (lane >> 3) + 4 * i, c = lane & 7;
                const int64_t slot = __shfl_sync(0xffffffffu, myslot, r);
                const int ok = slot >= 0 ? 16 : 0;
                const int64_t off = (slot >= 0 ? slot : 0) * p.cs.tok + c * 16;
                cp_async16(sK + tile_off(r, c), kbase + off, ok);
                cp_async16(sV + tile_off(r, c), vbase + off, ok);
            }
            {
                const int r = lane >> 1, c = lane & 1;
                const int64_t slot = __shfl_sync(0xffffffffu, myslot, r);
                const int ok = slot >= 0 ? 16 : 0;
                const int64_t off = (slot >= 0 ? slot : 0) * (p.cs.tok / 8) + c * 8;  // fp16 elements
                cp_async16(sKS + r * 32 + c * 16, ksbase + off, ok);
                cp_async16(sVS + r * 32 + c * 16, vsbase + off, ok);
            }
        }
    };

    WarpState st;
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int r = 0; r < 4; ++r) st.o[j][r] = 0.f;
    st.m_run[0] = st.m_run[1] = -INFINITY;
    st.l_run[0] = st.l_run[1] = 0.f;  // per-lane partials of the sums of heads 2t, 2t + 1
    const float sl2 = p.sm_scale * 1.4426950408889634f;

    // prologue: NSTAGE - 1 units in flight
    int u_issue = u0 + warp;
#pragma unroll
    for (int s = 0; s < NSTAGE - 1; ++s) {
        if (u_issue < u1) load_unit(u_issue, s);
        if constexpr (!TMA) cp_async_commit();
        if constexpr (SLIM) walk_advance();
        u_issue += WARPS;
    }
    int stg = 0;
    uint32_t phase = 0;
    const unsigned long long tr1 = p.trace ? global_timer_ns() : 0ull;
    for (int u = u0 + warp; u < u1; u += WARPS) {
        if constexpr (TMA) {
            mbar_wait(wbar + 8u * stg, phase);
        } else {
            cp_async_wait<NSTAGE - 2>();
        }
        __syncwarp();
        {   // refill the stage consumed in the previous iteration
            const int st_next = (stg + NSTAGE - 1) % NSTAGE;
            if (u_issue < u1) load_unit(u_issue, st_next);
            if constexpr (!TMA) cp_async_commit();
            if constexpr (SLIM) walk_advance();
            u_issue += WARPS;
        }
        uint8_t* stage = wsm + stg * STG;
        if constexpr (KV16) {
            // the tail unit's rows past kv_len hold whatever the cache holds there: zero their V rows
            // (their scores are masked to -inf in softmax_unit; 0 * NaN must not reach the accumulator)
            const int valid = kv_len - u * UNIT;
            if (valid < UNIT) {
                for (int idx = lane; idx < (UNIT - valid) * 16; idx += 32) {
                    const int r = valid + (idx >> 4), c = idx & 15;
                    *reinterpret_cast<uint4*>(stage + K_BYTES + (c >> 3) * 2 * K_BYTES + tile_off(r, c & 7)) = make_uint4(0, 0, 0, 0);
                }
                __syncwarp();
            }
            process_unit16<SLIM>(wsm_u32 + stg * STG, qa, st, u * UNIT, kv_len, sl2, g, lane);
            if (++stg == NSTAGE) { stg = 0; phase ^= 1; }
            continue;
        }
        if constexpr (TMA) {
            // the tail unit's rows past kv_len hold whatever the cache holds there: neutralise their V scales
            // (their scores are masked to -inf below; 0 * NaN must not reach the accumulator)
            const int valid = kv_len - u * UNIT;
            if (valid < UNIT) {
                if (lane >= valid && lane < UNIT) {
                    uint4* row = reinterpret_cast<uint4*>(stage + 2 * K_BYTES + S_BYTES + lane * 32);
                    row[0] = make_uint4(0, 0, 0, 0);
                    row[1] = make_uint4(0, 0, 0, 0);
                }
                __syncwarp();
            }
        }
        process_unit<SLIM>(stage, qa, st, u * UNIT, kv_len, sl2, g, t);
        if (++stg == NSTAGE) { stg = 0; phase ^= 1; }
    }
    if constexpr (!TMA) cp_async_wait<0>();

    // ---- reduce the row sums over the 8 g-lanes; lane (g, t) then owns d = [16 g, 16 g + 16) of heads 2t, 2t + 1
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        st.l_run[c] += __shfl_xor_sync(0xffffffffu, st.l_run[c], 4);
        st.l_run[c] += __shfl_xor_sync(0xffffffffu, st.l_run[c], 8);
        st.l_run[c] += __shfl_xor_sync(0xffffffffu, st.l_run[c], 16);
    }
    if constexpr (WARPS == 1) {
        // single warp: normalise and store straight from the accumulator fragments (32 contiguous bytes per lane)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int rq = 2 * t + c;
            if (rq >= nrow) continue;
            const int hq = hq0 + rq;
            if (p.nsplit == 1 && KV16) {
                const float inv = 1.f / st.l_run[c];
                __half* dst = p.out + tok * (int64_t)p.nq * 128 + (int64_t)hq * 128;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    dst[acc_dim<true>(i, g, 0)] = __float2half_rn(st.o[i][c] * inv);
                    dst[acc_dim<true>(i, g, 1)] = __float2half_rn(st.o[i][2 + c] * inv);
                }
            } else if (p.nsplit == 1) {
                const float inv = 1.f / st.l_run[c];
                uint32_t pk[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    __half2 lo = __floats2half2_rn(st.o[2 * i][c] * inv, st.o[2 * i + 1][c] * inv);
                    __half2 hi = __floats2half2_rn(st.o[2 * i][2 + c] * inv, st.o[2 * i + 1][2 + c] * inv);
                    pk[i] = *reinterpret_cast<uint32_t*>(&lo);
                    pk[4 + i] = *reinterpret_cast<uint32_t*>(&hi);
                }
                uint4* dst = reinterpret_cast<uint4*>(p.out + tok * (int64_t)p.nq * 128 + (int64_t)hq * 128 + 16 * g);
                dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            } else {
                float* wrow = p.ws + (((int64_t)b * p.nq + hq) * p.nsplit + split) * 130;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    wrow[acc_dim<KV16>(i, g, 0)] = st.o[i][c];
                    wrow[acc_dim<KV16>(i, g, 1)] = st.o[i][2 + c];
                }
                if (g == 0) {
                    wrow[128] = st.m_run[c];
                    wrow[129] = st.l_run[c];
                }
            }
        }
        trace_cta(p.trace, tr0, tr1);
        return;
    }
    __syncthreads();  // all rings are dead from here on
    float* red = reinterpret_cast<float*>(smem);  // [WARPS][G][130]: 128 o, m, l
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const int rq = 2 * t + c;
        if (rq >= G) continue;
        float* row = red + (warp * G + rq) * 130;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            row[acc_dim<KV16>(i, g, 0)] = st.o[i][c];
            row[acc_dim<KV16>(i, g, 1)] = st.o[i][2 + c];
        }
        if (g == 0) {
            row[128] = st.m_run[c];
            row[129] = st.l_run[c];
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < nrow * 128; idx += WARPS * 32) {
        const int rq = idx >> 7, d = idx & 127;
        float mm = -INFINITY;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) mm = fmaxf(mm, red[(w * G + rq) * 130 + 128]);
        const float ms = mm == -INFINITY ? 0.f : mm;
        float acc = 0.f, ll = 0.f;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const float* row = red + (w * G + rq) * 130;
            const float f = exp2f(row[128] - ms);
            acc += row[d] * f;
            ll += row[129] * f;
        }
        const int hq = hq0 + rq;
        if (p.nsplit == 1) {
            p.out[tok * (int64_t)p.nq * 128 + (int64_t)hq * 128 + d] = __float2half_rn(acc / ll);
        } else {
            float* wrow = p.ws + (((int64_t)b * p.nq + hq) * p.nsplit + split) * 130;
            wrow[d] = acc;
            if (d == 0) {
                wrow[128] = mm;
                wrow[129] = ll;
            }
        }
    }
    trace_cta(p.trace, tr0, tr1);
}

// merge split partials: one warp per (sequence, q head).  The per-split maxima / sums are read lane-parallel (one split per
// lane) and the partial rows by an unrolled loop of 8-byte loads, so the loads of many splits are in flight at once: the
// first version walked the splits in two scalar loops, a chain of ~2 x nsplit memory latencies that made the merge ~20 us of
// a 137 us launch at 13 splits (round 2 run 41).  Both sums still run in ascending split order: bit-identical results.
__global__ void __launch_bounds__(128) attn_merge_kernel(const float* __restrict__ ws, int nq, int nsplit, int64_t rows,
                                                        __half* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (w >= rows) return;
    const float* base = ws + w * nsplit * 130;
    float mm = -INFINITY;
    for (int s = lane; s < nsplit; s += 32) mm = fmaxf(mm, __ldcg(base + s * 130 + 128));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mm = fmaxf(mm, __shfl_xor_sync(0xffffffffu, mm, o));
    const float ms = mm == -INFINITY ? 0.f : mm;
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, ll = 0.f;
    for (int s0 = 0; s0 < nsplit; s0 += 32) {
        float f_l = 0.f, l_l = 0.f;   // lane j: rescaling factor and row sum of split s0 + j
        if (s0 + lane < nsplit) {
            const float2 ml = __ldcg(reinterpret_cast<const float2*>(base + (s0 + lane) * 130 + 128));
            f_l = exp2f(ml.x - ms);
            l_l = ml.y;
        }
        const int cnt = min(32, nsplit - s0);
#pragma unroll 8
        for (int j = 0; j < cnt; ++j) {
            const float f = __shfl_sync(0xffffffffu, f_l, j);
            ll += __shfl_sync(0xffffffffu, l_l, j) * f;   // same fused multiply-add, same order as the scalar loop it replaces
            const float2* row = reinterpret_cast<const float2*>(base + (s0 + j) * 130 + lane * 4);   // 520-byte rows: 8-byte aligned
            const float2 a = __ldcg(row), b = __ldcg(row + 1);
            acc[0] += a.x * f;
            acc[1] += a.y * f;
            acc[2] += b.x * f;
            acc[3] += b.y * f;
        }
    }
    // w = b * nq + hq and decode token index == b
    const __half2 h0 = __floats2half2_rn(acc[0] / ll, acc[1] / ll), h1 = __floats2half2_rn(acc[2] / ll, acc[3] / ll);
    *reinterpret_cast<uint2*>(out + w * 128 + lane * 4) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
}

// How many KV splits and how many warps per CTA.  The kernel's unit of residency is a warp (12 warp slots per SM,
// 148 SMs): a launch of W = ctas * warps warps runs in ceil(W / 1776) waves and the last, partial wave costs as much as
// a full one.  Round 2 run 6 showed what that does away from the 1-GPU benchmark shape: 70B TP = 8 (256 sequences x 1 kv
// head x 8192 tokens) ran 512 CTAs x 4 warps = 1.15 waves at 0.60 of the HBM peak, 7B TP = 8 (4096 CTAs) 2.3 waves at
// 0.76.  So: the smallest splits x warps whose wave efficiency reaches 0.93 (else the most efficient one), keeping at
// least 8 units (128 tokens) per warp and the split partials inside the workspace (attention_workspace_rows).
struct DecodePlan {
    int nsplit, warps;
};
constexpr int64_t kWarpSlots = 148 * kCtasPerSm;

int64_t attention_workspace_rows(int64_t batch) { return std::max<int64_t>(2 * batch + 444, 4096); }

DecodePlan plan_decode(int64_t base_ctas, int64_t batch, int64_t max_kv_len) {
    const int64_t units = (max_kv_len + UNIT - 1) / UNIT;
    auto eff = [&](int64_t n, int64_t w) {
        const double waves = (double)(base_ctas * n * w) / (double)kWarpSlots;
        return waves / ceil(waves);
    };
    DecodePlan best{1, 1};
    double best_eff = eff(1, 1);
    if (best_eff >= 0.90) return best;
    // ONE wave when the sequences can be cut into between 0.75 and 1.0 of the warp slots: every CTA starts at once and they
    // finish together -- no second-wave ramp, no staggered drain, the fewest partials to merge.  ~1500 one-warp CTAs already
    // saturate HBM (per-CTA timelines, profiles/r2_attention_cta_timeline_run41.txt), so the idle slots cost nothing.  Measured
    // at the 70B / TP 8 per-rank shape (256 base CTAs, kv 8192; round 2 run 45): 6 splits = 0.86 of one wave 0.125 ms per
    // launch; 13 splits = 0.94 of two waves (what the efficiency search below picks) 0.144 ms; 5 / 4 splits 0.142 / 0.140 ms.
    if (base_ctas < kWarpSlots) {
        int64_t nw = kWarpSlots / base_ctas;                              // the finest cut that still fits one wave
        nw = std::min(nw, units / 8);                                      // at least 8 units per warp
        nw = std::min(nw, attention_workspace_rows(batch) / std::max<int64_t>(1, batch));   // one-warp CTAs: splits = nw
        if (nw >= 2 && (double)(base_ctas * nw) >= 0.75 * (double)kWarpSlots) return DecodePlan{(int)nw, 1};
    }
    for (int64_t nw = 2; nw <= 64; ++nw) {       // total ways a sequence's KV range is cut (splits x warps)
        if (units / nw < 8) break;
        for (int w : {4, 2, 1}) {                // prefer warps of one CTA (merged in shared memory) over splits
            if (nw % w) continue;
            const int64_t n = nw / w;
            if (batch * n > attention_workspace_rows(batch)) continue;
            const double e = eff(n, w);
            if (e > best_eff + 1e-9) {
                best_eff = e;
                best = DecodePlan{(int)n, w};
            }
            break;
        }
        if (best_eff >= 0.93) break;
    }
    return best;
}

AttnParams make_params(const AttnArgs& a) {
    AttnParams p{};
    p.qkv = a.qkv;
    p.seq_starts = a.step->seq_starts;
    p.start_pos = a.step->start_pos;
    p.cache_indices = a.step->cache_indices;
    p.batch = (int)a.step->batch;
    p.decoding_batches = (int)a.step->decoding_batches;
    p.max_pages = a.step->max_pages;
    p.nq = a.num_heads;
    p.nkv = a.geom.num_kv_heads;
    p.D = a.geom.head_dim;
    p.cache_mode = a.geom.cache_mode;
    p.page_size = a.geom.page_size;
    p.group = a.geom.quant_group;
    p.cache_prefill = a.step->cache_prefill;
    p.cs = kv_strides(a.geom);
    p.kv16 = a.geom.quant_group == 1 ? 1 : 0;
    p.cache = a.kv_cache + (int64_t)a.layer * p.cs.layer * (p.kv16 ? 2 : 1);
    p.scale = p.kv16 ? nullptr : a.kv_scale + (int64_t)a.layer * p.cs.layer / a.geom.quant_group;
    p.sm_scale = 1.0f / sqrtf((float)a.geom.head_dim);
    p.out = a.out;
    p.ws = reinterpret_cast<float*>(a.workspace);
    p.nsplit = 1;
    return p;
}

// ---- tensor maps over the whole cache / scale tensors, one pair per (pointer, geometry).
// Every layout of llm_engine.cc:118-169 is a 3-D byte tensor {row bytes, X, Y} whose boxes are
// 16 token rows of one (layer, k/v, head):
//   layout 3 [L,2,H,T,D]: {D, T, L*2*H}   box {D,16,1}   coords (0, slot, (l*2+kv)*H + h)
//   layout 2 [L,2,T,H,D]: {D, H, L*2*T}   box {D,1,16}   coords (0, h, (l*2+kv)*T + slot)
//   layout 1 [L,T,2,H,D]: {D, 2H, L*T}    box {D,1,16}   coords (0, kv*H + h, l*T + slot)
//   layout 0 [T,L,2,H,D]: {D, L*2*H, T}   box {D,1,16}   coords (0, (l*2+kv)*H + h, slot)
struct KvMaps {
    CUtensorMap kv, sc;
};
std::mutex g_kvmap_mutex;
std::map<std::tuple<const void*, const void*, int, int, int, int, uint64_t>, KvMaps> g_kvmap_cache;

bool make_kv_maps(const AttnArgs& a, KvMaps* out) {
    const b2llm_kv_geom& g = a.geom;
    const bool kv16 = g.quant_group == 1;
    auto key = std::make_tuple((const void*)a.kv_cache, (const void*)a.kv_scale, g.cache_layout, g.num_layers, g.num_kv_heads,
                               g.head_dim * (kv16 ? 2 : 1), (uint64_t)g.max_tokens);
    std::lock_guard<std::mutex> lk(g_kvmap_mutex);
    auto it = g_kvmap_cache.find(key);
    if (it != g_kvmap_cache.end()) {
        *out = it->second;
        return true;
    }
    const uint64_t L = g.num_layers, H = g.num_kv_heads, T = g.max_tokens;
    for (int which = 0; which < (kv16 ? 1 : 2); ++which) {
        // row bytes: int8 values (fp16 cache: 256 B rows read as two 128 B boxes) / fp16 scales
        const uint64_t rb = which == 0 ? (kv16 ? 256 : 128) : 32;
        uint64_t dims[3], strides[2];
        uint32_t box[3];
        dims[0] = rb;
        switch (g.cache_layout) {
            case 3: dims[1] = T; dims[2] = L * 2 * H; box[1] = UNIT; box[2] = 1; break;
            case 2: dims[1] = H; dims[2] = L * 2 * T; box[1] = 1; box[2] = UNIT; break;
            case 1: dims[1] = 2 * H; dims[2] = L * T; box[1] = 1; box[2] = UNIT; break;
            default: dims[1] = L * 2 * H; dims[2] = T; box[1] = 1; box[2] = UNIT; break;
        }
        box[0] = (uint32_t)(rb > 128 ? 128 : rb);
        strides[0] = rb;
        strides[1] = rb * dims[1];
        if (!tma_encode_bytes(which == 0 ? &out->kv : &out->sc, which == 0 ? (const void*)a.kv_cache : (const void*)a.kv_scale,
                              3, dims, strides, box, which == 0))
            return false;
    }
    if (kv16) out->sc = out->kv;  // no scale tensor: the kernel never touches this map
    if (g_kvmap_cache.size() > 64) g_kvmap_cache.clear();
    g_kvmap_cache[key] = *out;
    return true;
}

// 4-D maps for LOADER 3: k / v as the outermost box dimension (layouts 2 and 3 only)
//   layout 3 [L,2,H,T,D]: {D, T, H, L*2}   box {D,16,1,2}   coords (0, slot, h, l*2)
//   layout 2 [L,2,T,H,D]: {D, H, T, L*2}   box {D,1,16,2}   coords (0, h, slot, l*2)
std::map<std::tuple<const void*, const void*, int, int, int, int, uint64_t>, KvMaps> g_kvmap4_cache;

bool make_kv_maps_merged(const AttnArgs& a, KvMaps* out) {
    const b2llm_kv_geom& g = a.geom;
    if (g.cache_layout != 2 && g.cache_layout != 3) return false;
    const bool kv16 = g.quant_group == 1;
    auto key = std::make_tuple((const void*)a.kv_cache, (const void*)a.kv_scale, g.cache_layout, g.num_layers, g.num_kv_heads,
                               g.head_dim * (kv16 ? 2 : 1), (uint64_t)g.max_tokens);
    std::lock_guard<std::mutex> lk(g_kvmap_mutex);
    auto it = g_kvmap4_cache.find(key);
    if (it != g_kvmap4_cache.end()) {
        *out = it->second;
        return true;
    }
    const uint64_t L = g.num_layers, H = g.num_kv_heads, T = g.max_tokens;
    for (int which = 0; which < (kv16 ? 1 : 2); ++which) {
        const uint64_t rb = which == 0 ? (kv16 ? 256 : 128) : 32;
        uint64_t dims[4], strides[3];
        uint32_t box[4];
        dims[0] = rb;
        dims[3] = L * 2;
        box[0] = (uint32_t)(rb > 128 ? 128 : rb);
        box[3] = 2;
        if (g.cache_layout == 3) {
            dims[1] = T; dims[2] = H; box[1] = UNIT; box[2] = 1;
        } else {
            dims[1] = H; dims[2] = T; box[1] = 1; box[2] = UNIT;
        }
        strides[0] = rb;
        strides[1] = rb * dims[1];
        strides[2] = rb * dims[1] * dims[2];
        if (!tma_encode_bytes(which == 0 ? &out->kv : &out->sc, which == 0 ? (const void*)a.kv_cache : (const void*)a.kv_scale,
                              4, dims, strides, box, which == 0))
            return false;
    }
    if (kv16) out->sc = out->kv;
    if (g_kvmap4_cache.size() > 64) g_kvmap4_cache.clear();
    g_kvmap4_cache[key] = *out;
    return true;
}

DecodeTma make_tma_coords_merged(const AttnArgs& a) {
    DecodeTma c{};
    c.tok_dim = a.geom.cache_layout == 3 ? 1 : 2;
    c.k_fixed = a.layer * 2;  // the (layer, k/v) coordinate of K; V is the second element of the box
    return c;
}

DecodeTma make_tma_coords(const AttnArgs& a) {
    const b2llm_kv_geom& g = a.geom;
    const int L = a.layer, H = g.num_kv_heads;
    const int64_t T = (int64_t)g.max_tokens;
    DecodeTma c{};
    switch (g.cache_layout) {
        case 3: c.tok_dim = 1; c.k_fixed = (L * 2 + 0) * H; c.v_fixed = (L * 2 + 1) * H; c.k_tok0 = 0; c.v_tok0 = 0; break;
        case 2: c.tok_dim = 2; c.k_fixed = 0; c.v_fixed = 0; c.k_tok0 = (int)((L * 2 + 0) * T); c.v_tok0 = (int)((L * 2 + 1) * T); break;
        case 1: c.tok_dim = 2; c.k_fixed = 0; c.v_fixed = H; c.k_tok0 = (int)(L * T); c.v_tok0 = (int)(L * T); break;
        default: c.tok_dim = 2; c.k_fixed = (L * 2 + 0) * H; c.v_fixed = (L * 2 + 1) * H; c.k_tok0 = 0; c.v_tok0 = 0; break;
    }
    return c;
}

int g_attn_warps_override = -1, g_attn_tma_override = -1, g_attn_slim = -1;
std::once_flag g_attn_env_once;

}  // namespace

void attention_decode_plan(int64_t base_ctas, int64_t batch, int64_t max_kv_len, int* nsplit, int* warps) {
    const DecodePlan p = plan_decode(base_ctas, batch, max_kv_len);
    *nsplit = p.nsplit;
    *warps = p.warps;
}

int64_t attention_workspace_bytes(int64_t batch, int num_heads, int head_dim) {
    (void)head_dim;
    // rows = sequences * q heads * splits; plan_decode keeps sequences * splits <= attention_workspace_rows and the
    // "always split" mode (split_k == 2) needs 2 per sequence
    return attention_workspace_rows(batch) * num_heads * 130 * (int64_t)sizeof(float);
}

int32_t launch_attention_simple(cudaStream_t s, const AttnArgs& a, int64_t token_begin, int64_t token_end) {
    B2_REQUIRE(a.geom.quant_group == 8 || a.geom.quant_group == 1, B2LLM_ERR_UNSUPPORTED,
               "kv cache: int8 with quant group 8, or fp16 (quant group 1)");
    if (token_end <= token_begin) return B2LLM_OK;
    AttnParams p = make_params(a);
    p.token_begin = token_begin;
    p.token_end = token_end;
    const int64_t warps = (token_end - token_begin) * a.num_heads;
    const unsigned blocks = (unsigned)((warps + 3) / 4);
    if (a.geom.head_dim == 128)
        attn_simple_kernel<128><<<blocks, 128, 0, s>>>(p);
    else if (a.geom.head_dim == 64)
        attn_simple_kernel<64><<<blocks, 128, 0, s>>>(p);
    else {
        set_last_error("attention: head_dim must be 64 or 128");
        return B2LLM_ERR_UNSUPPORTED;
    }
    B2_LAUNCH_CHECK();
    return B2LLM_OK;
}

// debug: device buffer armed by b2llm_debug_attention_trace (4 x uint64 per CTA), nullptr = off
std::atomic<unsigned long long*> g_attn_trace{nullptr};
std::atomic<long long> g_attn_trace_ctas{0};

template <int G, int WARPS, int LOADER, bool KV16 = false>
static int32_t launch_decode(cudaStream_t s, AttnParams& p, const KvMaps& maps, const DecodeTma& tc) {
    auto kern = attn_decode_kernel<G, WARPS, LOADER, KV16>;
    constexpr int STG = KV16 ? STAGE16 : STAGE;
    constexpr int smem_bytes = WARPS * NSTAGE * STG + 1024 + 8 * WARPS * NSTAGE + 64 +
                               (WARPS * G * 130 * 4 > WARPS * NSTAGE * STG ? WARPS * G * 130 * 4 : 0);
    B2_ENSURE_DYN_SMEM(kern, smem_bytes);
    const int gq = p.nq / p.nkv;
    const int chunks = (gq + G - 1) / G;
    dim3 grid(p.nsplit, p.nkv * chunks, p.decoding_batches);
    unsigned long long* tr = g_attn_trace.load();
    p.trace = tr != nullptr && (long long)grid.x * grid.y * grid.z <= g_attn_trace_ctas.load() ? tr : nullptr;
    launch_kernel(kern, grid, dim3(WARPS * 32), smem_bytes, s, p, maps.kv, maps.sc, tc);
    B2_LAUNCH_CHECK();
    if (p.nsplit > 1) {
        const int64_t rows = (int64_t)p.decoding_batches * p.nq;
        launch_kernel(attn_merge_kernel, dim3((unsigned)((rows + 3) / 4)), dim3(128), 0, s, (const float*)p.ws, p.nq, p.nsplit, rows, p.out);
        B2_LAUNCH_CHECK();
    }
    return B2LLM_OK;
}

template <int G>
static int32_t dispatch_decode(cudaStream_t s, AttnParams& p, const KvMaps& maps, const DecodeTma& tc, int warps, bool tma,
                               int slim) {
    if (p.kv16) {  // fp16 cache: merged / slim TMA loaders or cp.async
        if (tma && slim == 2) {
            if (warps == 1) return launch_decode<G, 1, 3, true>(s, p, maps, tc);
            if (warps == 2) return launch_decode<G, 2, 3, true>(s, p, maps, tc);
            return launch_decode<G, 4, 3, true>(s, p, maps, tc);
        }
        if (tma) {
            if (warps == 1) return launch_decode<G, 1, 2, true>(s, p, maps, tc);
            if (warps == 2) return launch_decode<G, 2, 2, true>(s, p, maps, tc);
            return launch_decode<G, 4, 2, true>(s, p, maps, tc);
        }
        if (warps == 1) return launch_decode<G, 1, 0, true>(s, p, maps, tc);
        if (warps == 2) return launch_decode<G, 2, 0, true>(s, p, maps, tc);
        return launch_decode<G, 4, 0, true>(s, p, maps, tc);
    }
    if (tma && slim == 2) {
        if (warps == 1) return launch_decode<G, 1, 3>(s, p, maps, tc);
        if (warps == 2) return launch_decode<G, 2, 3>(s, p, maps, tc);
        return launch_decode<G, 4, 3>(s, p, maps, tc);
    }
    if (tma && slim) {
        if (warps == 1) return launch_decode<G, 1, 2>(s, p, maps, tc);
        if (warps == 2) return launch_decode<G, 2, 2>(s, p, maps, tc);
        return launch_decode<G, 4, 2>(s, p, maps, tc);
    }
    if (tma) {
        if (warps == 1) return launch_decode<G, 1, 1>(s, p, maps, tc);
        if (warps == 2) return launch_decode<G, 2, 1>(s, p, maps, tc);
        return launch_decode<G, 4, 1>(s, p, maps, tc);
    }
    if (warps == 1) return launch_decode<G, 1, 0>(s, p, maps, tc);
    if (warps == 2) return launch_decode<G, 2, 0>(s, p, maps, tc);
    return launch_decode<G, 4, 0>(s, p, maps, tc);
}

int32_t launch_attention_decode_mma(cudaStream_t s, const AttnArgs& a) {
    B2_REQUIRE(a.geom.head_dim == 128 && (a.geom.quant_group == 8 || a.geom.quant_group == 1), B2LLM_ERR_UNSUPPORTED,
               "attention (tensor-core path): head_dim 128; int8 group-8 or fp16 cache");
    B2_REQUIRE(a.step->decoding_batches <= 65535, B2LLM_ERR_INVALID_VALUE, "too many decoding sequences");
    if (a.step->decoding_batches == 0) return B2LLM_OK;
    std::call_once(g_attn_env_once, [] {
        if (const char* e = getenv("B2LLM_ATTN_WARPS")) g_attn_warps_override = atoi(e);
        if (const char* e = getenv("B2LLM_ATTN_TMA")) g_attn_tma_override = atoi(e);
        if (const char* e = getenv("B2LLM_ATTN_SLIM")) g_attn_slim = atoi(e);
    });
    AttnParams p = make_params(a);
    const int gq = p.nq / p.nkv;
    const int64_t max_kv = a.step->max_kv_len > 0 ? a.step->max_kv_len : 1;
    const int G = gq == 1 ? 1 : (gq <= 4 ? 4 : 8);
    const int chunks = (gq + G - 1) / G;
    const int64_t base_ctas = (int64_t)p.nkv * chunks * p.decoding_batches;
    const DecodePlan plan = plan_decode(base_ctas, p.decoding_batches, max_kv);
    p.nsplit = plan.nsplit;
    int warps = plan.warps;
    if (a.split_k == 0) p.nsplit = 1;     // ENGINE_CONF_DECODING_ATTN_SPLIT_K 0: never split (warps of one CTA still share a sequence)
    if (a.split_k == 2 && p.nsplit < 2 && max_kv > UNIT) p.nsplit = 2;
    if (g_attn_warps_override == 1 || g_attn_warps_override == 2 || g_attn_warps_override == 4) warps = g_attn_warps_override;
    {   // B2LLM_ATTN_SPLITS: force the number of KV splits (experiments; must fit the workspace)
        static const int splits_override = [] { const char* e = getenv("B2LLM_ATTN_SPLITS"); return e ? atoi(e) : 0; }();
        if (splits_override >= 1 && splits_override <= 64 &&
            (int64_t)p.decoding_batches * splits_override <= attention_workspace_rows(p.decoding_batches))
            p.nsplit = splits_override;
    }
    // TMA loader: units must be 16 contiguous slots
    bool tma = tma_available() && (p.cache_mode == 0 || p.page_size % UNIT == 0) &&
               (int64_t)a.geom.max_tokens * a.geom.num_layers * 2 < (1ll << 31);
    if (g_attn_tma_override == 0) tma = false;
    KvMaps maps{};
    DecodeTma tc{};
    if (tma) {
        if (!make_kv_maps(a, &maps)) tma = false;
        else tc = make_tma_coords(a);
    }
    // loader defaults = what ran green on the device: the slim loader (run 17 of round 1) and, since round 2 run 1
    // (profiles/r2_bringup_run1.txt: all layouts x paged / indexed x page 16 / 64 / 128; in-step 0.849 vs 0.883 ms per
    // launch), the merged K + V loads (LOADER 3) wherever the layout allows them (2 and 3; others fall back to slim).
    // B2LLM_ATTN_SLIM = 0 dividing loader, 1 slim, 2 slim + merged loads.
    int slim = g_attn_slim < 0 ? 2 : g_attn_slim;
    if (a.loader >= 0) slim = a.loader;  // explicit choice of the caller (parity tests of every loader)
    if (tma && slim == 2) {  // merged K + V loads: layouts 2 / 3 only, else the plain slim loader
        KvMaps merged{};
        if (make_kv_maps_merged(a, &merged)) {
            maps = merged;
            tc = make_tma_coords_merged(a);
        } else {
            slim = 1;
        }
    }
    if (G == 1) return dispatch_decode<1>(s, p, maps, tc, warps, tma, slim);
    if (G == 4) return dispatch_decode<4>(s, p, maps, tc, warps, tma, slim);
    return dispatch_decode<8>(s, p, maps, tc, warps, tma, slim);
}

}  // namespace b2llm

extern "C" int32_t b2llm_debug_attention_trace(void* device_buf, int64_t capacity_ctas) {
    b2llm::g_attn_trace_ctas.store(device_buf != nullptr ? (long long)capacity_ctas : 0);
    b2llm::g_attn_trace.store(reinterpret_cast<unsigned long long*>(device_buf));
    return B2LLM_OK;
}
