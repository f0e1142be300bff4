// K6 (+K7 reference form) of SURVEY.md section 2.3: attention over the paged / indexed int8 KV cache.
//
//  * attn_decode_mma_kernel -- THE dominant kernel of the decode step (HBM-bound: it streams
//    kv_len * 320 B per (sequence, kv head)).  One CTA per (sequence, kv head, kv split); 4 warps,
//    each with a private 3-stage cp.async ring of 16-token units (K 2 KB + V 2 KB + scales 1 KB), so
//    warps never block each other and ~60 KB per CTA are in flight.  Rows of 128 B (one token-head)
//    are fetched as 8 coalesced 16 B chunks into XOR-swizzled shared memory.  int8 -> fp16 dequant
//    happens in registers (magic-number trick: (b ^ 0x80) | 0x6400 == 1024 + (b + 128), minus 1152,
//    times the group scale) directly into mma.sync m16n8k16 fragments:
//        S[q-head, token]  = Q[q-head, d]   . K^T[d, token]     (A = Q, B = K as stored: token-major)
//        O[q-head, d]      = P[q-head, tok] . V[tok, d]         (A = P from S's accumulator layout)
//    the head-dim and token orders inside a fragment are permuted so that every lane reads whole
//    16-byte chunks; fp32 accumulation, online softmax in the exp2 domain, split-KV partials merged
//    by attn_merge_kernel.  GQA packs up to 8 q-heads of a kv head into the MMA M dimension.
//  * attn_simple_kernel -- one warp per (token, q-head), straightforward fp32 math.  Reference form
//    used for prefill tokens (fresh fp16 K/V + optional cached prefix) and as the checker of the
//    MMA kernel (impl = 1).
//
// Numeric contract: oracle/llama_ref.py attention_decode / _attend.
#include "common.cuh"

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
    const int8_t* cache;   // layer offset applied
    const __half* scale;   // layer offset applied
    KvStrides cs;
    float sm_scale;        // 1 / sqrt(D)
    __half* out;           // [T, nq * D]
    float* ws;             // split partials
    int nsplit;
    int64_t token_begin, token_end;
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
            const int8_t* kr = p.cache + off;
            const int8_t* vr = p.cache + p.cs.kv + off;
            const float ks = __half2float(p.scale[off / p.group + g]);
            const float vs = __half2float(p.scale[(p.cs.kv + off) / p.group + g]);
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                kx[i] = __half2float(__float2half_rn(__fmul_rn((float)kr[lane * PER + i], ks)));
                vx[i] = __half2float(__float2half_rn(__fmul_rn((float)vr[lane * PER + i], vs)));
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
constexpr int NSTAGE = 3;
constexpr int K_BYTES = UNIT * 128;      // 2048
constexpr int S_BYTES = UNIT * 32;       // 512 (16 fp16 scales per token)
constexpr int STAGE = 2 * K_BYTES + 2 * S_BYTES;  // 5120
constexpr int WARPS = 4;
constexpr int ATT_SMEM = WARPS * NSTAGE * STAGE;  // 61440

__device__ __forceinline__ int swz_f(int r) { return (r & 6) ^ ((r & 1) << 2); }

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

template <int G>  // q heads per CTA (rows of the MMA M dimension in use), 1..8
__global__ void __launch_bounds__(WARPS * 32, 3) attn_decode_mma_kernel(AttnParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int split = blockIdx.x, b = blockIdx.z;
    const int gq = p.nq / p.nkv;                 // q heads per kv head
    const int chunks = (gq + G - 1) / G;          // CTAs per kv head along y
    const int hk = blockIdx.y / chunks;
    const int hq0 = hk * gq + (blockIdx.y % chunks) * G;
    const int nrow = min(G, hk * gq + gq - hq0);  // valid q-head rows

    const int64_t kv_len = p.start_pos[b] + 1;
    const int64_t units_total = (kv_len + UNIT - 1) / UNIT;
    const int64_t units_per_split = (units_total + p.nsplit - 1) / p.nsplit;
    const int64_t u0 = split * units_per_split;
    const int64_t u1 = min(units_total, u0 + units_per_split);

    const int heads = p.nq + 2 * p.nkv;
    const int64_t tok = b;  // decode sequences come first, one token each (seq_starts[b] == b)

    // ---- Q fragments (A operand), rows >= nrow are zero.  k-slot order follows the K byte order:
    // step j, lane t covers d = base(j,t) + {0,2} (A0) and {1,3} (A2), base = (j<4 ? 16t : 64+16t) + 4(j&3)
    uint32_t qa[8][2];
    {
        const __half* qrow = p.qkv + tok * (int64_t)heads * 128 + (int64_t)(hq0 + g) * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int base = (j < 4 ? 16 * t : 64 + 16 * t) + 4 * (j & 3);
            if (g < nrow) {
                const uint2 v = *reinterpret_cast<const uint2*>(qrow + base);  // halves d..d+3
                qa[j][0] = prmt(v.x, v.y, 0x5410);  // (d+0, d+2)
                qa[j][1] = prmt(v.x, v.y, 0x7632);  // (d+1, d+3)
            } else {
                qa[j][0] = 0u;
                qa[j][1] = 0u;
            }
        }
    }

    const int8_t* kbase = p.cache + hk * p.cs.head;
    const int8_t* vbase = kbase + p.cs.kv;
    const __half* ksbase = p.scale + hk * p.cs.head / 8;
    const __half* vsbase = ksbase + p.cs.kv / 8;
    uint8_t* wsm = smem + warp * (NSTAGE * STAGE);
    const uint32_t wsm_u32 = smem_u32(wsm);

    // issue the loads of unit u into stage st (warp-collective; 10 x 16 B per lane)
    auto load_unit = [&](int64_t u, int st) {
        const uint32_t sK = wsm_u32 + st * STAGE, sV = sK + K_BYTES, sKS = sV + K_BYTES, sVS = sKS + S_BYTES;
        // lanes 0..15 look up the slot of token u*16 + lane
        int64_t myslot = -1;
        {
            const int64_t pos = u * UNIT + (lane & 15);
            if (pos < kv_len) myslot = kv_slot(p.cache_indices, p.cache_mode, p.page_size, p.max_pages, b, pos);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = (lane >> 3) + 4 * i, c = lane & 7;
            const int64_t slot = __shfl_sync(0xffffffffu, myslot, r);
            const int ok = slot >= 0 ? 16 : 0;
            const int64_t off = (slot >= 0 ? slot : 0) * p.cs.tok + c * 16;
            const uint32_t d = r * 128 + ((c ^ swz_f(r)) << 4);
            cp_async16(sK + d, kbase + off, ok);
            cp_async16(sV + d, vbase + off, ok);
        }
        {
            const int r = lane >> 1, c = lane & 1;
            const int64_t slot = __shfl_sync(0xffffffffu, myslot, r);
            const int ok = slot >= 0 ? 16 : 0;
            const int64_t off = (slot >= 0 ? slot : 0) * (p.cs.tok / 8) + c * 8;  // fp16 elements
            cp_async16(sKS + r * 32 + c * 16, ksbase + off, ok);
            cp_async16(sVS + r * 32 + c * 16, vsbase + off, ok);
        }
    };

    float o[16][4];
#pragma unroll
    for (int j = 0; j < 16; ++j)
#pragma unroll
        for (int r = 0; r < 4; ++r) o[j][r] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;   // state of q-head row g (replicated over the 4 t-lanes; l is per-lane partial)
    const float sl2 = p.sm_scale * 1.4426950408889634f;

    // prologue
    int64_t u_issue = u0 + warp;
#pragma unroll
    for (int s = 0; s < NSTAGE - 1; ++s) {
        if (u_issue < u1) load_unit(u_issue, s);
        cp_async_commit();
        u_issue += WARPS;
    }
    int st = 0;
    for (int64_t u = u0 + warp; u < u1; u += WARPS) {
        cp_async_wait<NSTAGE - 2>();
        __syncwarp();
        {   // refill the stage consumed in the previous iteration
            const int st_next = (st + NSTAGE - 1) % NSTAGE;
            if (u_issue < u1) load_unit(u_issue, st_next);
            cp_async_commit();
            u_issue += WARPS;
        }
        const uint8_t* sK = wsm + st * STAGE;
        const uint8_t* sV = sK + K_BYTES;
        const uint8_t* sKS = sV + K_BYTES;
        const uint8_t* sVS = sKS + S_BYTES;

        // ---- S = Q K^T for 2 n-tiles of 8 tokens
        float s_acc[2][4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
            for (int r = 0; r < 4; ++r) s_acc[nt][r] = 0.f;
            const int r = 8 * nt + g;
            const uint4 ca = *reinterpret_cast<const uint4*>(sK + r * 128 + ((t ^ swz_f(r)) << 4));
            const uint4 cb = *reinterpret_cast<const uint4*>(sK + r * 128 + (((4 + t) ^ swz_f(r)) << 4));
            const __half2 sa = *reinterpret_cast<const __half2*>(sKS + r * 32 + 4 * t);        // groups 2t, 2t+1
            const __half2 sb = *reinterpret_cast<const __half2*>(sKS + r * 32 + 16 + 4 * t);   // groups 8+2t, 8+2t+1
            const uint32_t wds[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
            const __half2 scs[4] = {__low2half2(sa), __high2half2(sa), __low2half2(sb), __high2half2(sb)};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t wd = wds[j];
                const uint32_t b0 = deq2(lop3_and_xor(wd, 0x00FF00FFu, 0x64806480u), scs[j >> 1]);       // bytes 0,2
                const uint32_t b1 = deq2(lop3_and_xor(wd >> 8, 0x00FF00FFu, 0x64806480u), scs[j >> 1]);  // bytes 1,3
                mma_f16_16816(s_acc[nt], qa[j][0], 0u, qa[j][1], 0u, b0, b1);
            }
        }

        // ---- online softmax for q-head row g over tokens {2t, 2t+1, 8+2t, 9+2t} of this unit
        const int64_t tbase = u * UNIT;
        float sv[4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int64_t pos = tbase + 8 * nt + 2 * t + c;
                sv[nt * 2 + c] = pos < kv_len ? s_acc[nt][c] * sl2 : -INFINITY;
            }
        float mx = fmaxf(fmaxf(sv[0], sv[1]), fmaxf(sv[2], sv[3]));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        const float m_new = fmaxf(m_run, mx);
        const float m_safe = m_new == -INFINITY ? 0.f : m_new;
        const float corr = exp2f(m_run - m_safe);
        float pv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) pv[i] = exp2f(sv[i] - m_safe);
        l_run = l_run * corr + (pv[0] + pv[1]) + (pv[2] + pv[3]);
        m_run = m_new;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            o[j][0] *= corr;
            o[j][1] *= corr;
        }
        // P enters the tensor core as hi + lo fp16 halves (~22 significant bits), so the only fp16
        // roundings on this path are the dequantised K / V values, which the oracle reproduces
        __half2 p01 = __floats2half2_rn(pv[0], pv[1]), p23 = __floats2half2_rn(pv[2], pv[3]);
        const float2 f01 = __half22float2(p01), f23 = __half22float2(p23);
        __half2 q01 = __floats2half2_rn(pv[0] - f01.x, pv[1] - f01.y), q23 = __floats2half2_rn(pv[2] - f23.x, pv[3] - f23.y);
        const uint32_t pa0 = *reinterpret_cast<uint32_t*>(&p01), pa2 = *reinterpret_cast<uint32_t*>(&p23);
        const uint32_t pl0 = *reinterpret_cast<uint32_t*>(&q01), pl2 = *reinterpret_cast<uint32_t*>(&q23);

        // ---- O += P V : lane (g, t) supplies column n = g of every n-tile, i.e. d = 16 g + j, for tokens
        // {2t, 2t+1} (B0) and {8+2t, 9+2t} (B1)
        {
            const int r0 = 2 * t, r1 = 2 * t + 1, r2 = 8 + 2 * t, r3 = 9 + 2 * t;
            const uint4 va = *reinterpret_cast<const uint4*>(sV + r0 * 128 + ((g ^ swz_f(r0)) << 4));
            const uint4 vb = *reinterpret_cast<const uint4*>(sV + r1 * 128 + ((g ^ swz_f(r1)) << 4));
            const uint4 vc = *reinterpret_cast<const uint4*>(sV + r2 * 128 + ((g ^ swz_f(r2)) << 4));
            const uint4 vd = *reinterpret_cast<const uint4*>(sV + r3 * 128 + ((g ^ swz_f(r3)) << 4));
            const uint32_t sA = *reinterpret_cast<const uint32_t*>(sVS + r0 * 32 + 4 * g);  // groups 2g, 2g+1
            const uint32_t sB = *reinterpret_cast<const uint32_t*>(sVS + r1 * 32 + 4 * g);
            const uint32_t sC = *reinterpret_cast<const uint32_t*>(sVS + r2 * 32 + 4 * g);
            const uint32_t sD = *reinterpret_cast<const uint32_t*>(sVS + r3 * 32 + 4 * g);
            uint32_t sc_ab[2] = {prmt(sA, sB, 0x5410), prmt(sA, sB, 0x7632)};  // (A.lo,B.lo), (A.hi,B.hi)
            uint32_t sc_cd[2] = {prmt(sC, sD, 0x5410), prmt(sC, sD, 0x7632)};
            const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
            const uint32_t wc[4] = {vc.x, vc.y, vc.z, vc.w}, wdd[4] = {vd.x, vd.y, vd.z, vd.w};
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int w = j >> 2, i = j & 3;
                const uint32_t sel = (uint32_t)(i | (i << 4) | ((4 + i) << 8) | ((4 + i) << 12));
                const __half2 s_ab = *reinterpret_cast<__half2*>(&sc_ab[j >> 3]);
                const __half2 s_cd = *reinterpret_cast<__half2*>(&sc_cd[j >> 3]);
                const uint32_t b0 = deq2(lop3_and_xor(prmt(wa[w], wb[w], sel), 0x00FF00FFu, 0x64806480u), s_ab);
                const uint32_t b1 = deq2(lop3_and_xor(prmt(wc[w], wdd[w], sel), 0x00FF00FFu, 0x64806480u), s_cd);
                mma_f16_16816(o[j], pa0, 0u, pa2, 0u, b0, b1);
                mma_f16_16816(o[j], pl0, 0u, pl2, 0u, b0, b1);
            }
        }
        st = (st + 1) % NSTAGE;
    }
    cp_async_wait<0>();

    // ---- reduce l over the 4 t-lanes, then merge the 4 warps through shared memory
    l_run += __shfl_xor_sync(0xffffffffu, l_run, 1);
    l_run += __shfl_xor_sync(0xffffffffu, l_run, 2);
    __syncthreads();  // all rings are dead from here on
    float* red = reinterpret_cast<float*>(smem);  // [WARPS][G][130]: 128 o, m, l
    if (g < G) {
        float* row = red + (warp * G + g) * 130;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            row[16 * (2 * t) + j] = o[j][0];
            row[16 * (2 * t + 1) + j] = o[j][1];
        }
        if (t == 0) {
            row[128] = m_run;
            row[129] = l_run;
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
}

// merge split partials: one warp per (sequence, q head)
__global__ void __launch_bounds__(128) attn_merge_kernel(const float* __restrict__ ws, int nq, int nsplit, int64_t rows,
                                                        __half* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (w >= rows) return;
    const float* base = ws + w * nsplit * 130;
    float mm = -INFINITY;
    for (int s = 0; s < nsplit; ++s) mm = fmaxf(mm, base[s * 130 + 128]);
    const float ms = mm == -INFINITY ? 0.f : mm;
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, ll = 0.f;
    for (int s = 0; s < nsplit; ++s) {
        const float f = exp2f(base[s * 130 + 128] - ms);
        ll += base[s * 130 + 129] * f;
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] += base[s * 130 + lane * 4 + i] * f;
    }
    // w = b * nq + hq and decode token index == b
#pragma unroll
    for (int i = 0; i < 4; ++i) out[w * 128 + lane * 4 + i] = __float2half_rn(acc[i] / ll);
}

int choose_splits(int64_t ctas_per_split, int64_t max_kv_len) {
    // fill ~3 CTAs/SM on 148 SMs; never make a split shorter than 4 units per warp
    const int64_t target = 148 * 3;
    int64_t n = (target + ctas_per_split - 1) / ctas_per_split;
    const int64_t max_by_len = (max_kv_len + 255) / 256;
    if (n > max_by_len) n = max_by_len;
    if (n < 1) n = 1;
    if (n > 64) n = 64;
    return (int)n;
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
    p.cache = a.kv_cache + (int64_t)a.layer * p.cs.layer;
    p.scale = a.kv_scale + (int64_t)a.layer * p.cs.layer / a.geom.quant_group;
    p.sm_scale = 1.0f / sqrtf((float)a.geom.head_dim);
    p.out = a.out;
    p.ws = reinterpret_cast<float*>(a.workspace);
    p.nsplit = 1;
    return p;
}

}  // namespace

int64_t attention_workspace_bytes(int64_t batch, int num_heads, int head_dim) {
    (void)head_dim;
    // rows = sequences * q heads * splits; choose_splits keeps sequences * splits <= sequences + 444
    return (batch + 444) * num_heads * 130 * (int64_t)sizeof(float);
}

int32_t launch_attention_simple(cudaStream_t s, const AttnArgs& a, int64_t token_begin, int64_t token_end) {
    B2_REQUIRE(a.geom.quant_group == 8, B2LLM_ERR_UNSUPPORTED, "kv cache: only int8 with quant group 8 is supported");
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

template <int G>
static int32_t launch_decode_g(cudaStream_t s, AttnParams& p, int64_t max_kv_len) {
    auto kern = attn_decode_mma_kernel<G>;
    static bool configured = false;
    if (!configured) {
        B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        configured = true;
    }
    const int gq = p.nq / p.nkv;
    const int chunks = (gq + G - 1) / G;
    const int64_t per_split = (int64_t)p.nkv * chunks * p.decoding_batches;
    p.nsplit = choose_splits(per_split, max_kv_len);
    dim3 grid(p.nsplit, p.nkv * chunks, p.decoding_batches);
    kern<<<grid, WARPS * 32, ATT_SMEM, s>>>(p);
    B2_LAUNCH_CHECK();
    if (p.nsplit > 1) {
        const int64_t rows = (int64_t)p.decoding_batches * p.nq;
        attn_merge_kernel<<<(unsigned)((rows + 3) / 4), 128, 0, s>>>(p.ws, p.nq, p.nsplit, rows, p.out);
        B2_LAUNCH_CHECK();
    }
    return B2LLM_OK;
}

int32_t launch_attention_decode_mma(cudaStream_t s, const AttnArgs& a) {
    B2_REQUIRE(a.geom.head_dim == 128 && a.geom.quant_group == 8, B2LLM_ERR_UNSUPPORTED,
               "attention (tensor-core path): head_dim 128 and int8 group-8 cache only");
    B2_REQUIRE(a.step->decoding_batches <= 65535, B2LLM_ERR_INVALID_VALUE, "too many decoding sequences");
    if (a.step->decoding_batches == 0) return B2LLM_OK;
    AttnParams p = make_params(a);
    const int gq = p.nq / p.nkv;
    const int64_t max_kv = a.step->max_kv_len > 0 ? a.step->max_kv_len : 1;
    if (gq == 1) return launch_decode_g<1>(s, p, max_kv);
    if (gq <= 4) return launch_decode_g<4>(s, p, max_kv);
    return launch_decode_g<8>(s, p, max_kv);
}

}  // namespace b2llm
