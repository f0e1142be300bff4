// K4 + K5 of SURVEY.md section 2.3: rotary embedding on q,k (in place in the qkv activation) and the
// append of k,v into the KV cache, for every token of the step: int8 group-8 quantised (cache_quant_bit 8) or as the
// fp16 values they are (cache_quant_bit 0 / group 1, llm_generator.cc:131-136: no scale tensor).
// Eight lanes per (token, head); 16-byte accesses; no shuffles (see the kernel comment).
//
// Numeric contract (oracle/llama_ref.py: apply_rope, kv_quant):
//   rotate-half pairing (i, i + D/2); o1 = x1*c - x2*s, o2 = x2*c + x1*s with every product and sum
//   individually rounded to fp32, result rounded to fp16;
//   scale16 = fp16(max|x| / 127) per 8 elements, q = clamp(rint(x / fp32(scale16)), -127, 127).
#include "common.cuh"

namespace b2llm {

namespace {

struct RopeKvParams {
    __half* qkv;
    const int64_t* seq_starts;
    const int64_t* start_pos;
    const int64_t* cache_indices;
    int batch;
    int64_t decoding_batches;
    int64_t num_tokens;
    int64_t max_pages;
    int nq, nkv, D;
    int cache_mode, page_size, group;
    const float* cos_t;
    const float* sin_t;
    int8_t* cache;   // layer offset applied
    __half* scale;   // layer offset applied
    KvStrides cs;    // cache strides (elements); scale strides = cs / group
};

// 8 lanes per (token, head): lane j owns the 16-byte chunk j of the low half (dims [8j, 8j+8)) and the matching
// chunk of the high half (dims [D/2 + 8j, D/2 + 8j + 8)) -- exactly the rotate-half partners, and exactly one
// quantisation group each, so neither the rotation nor the group max needs a shuffle.  D = 128 uses all 8 lanes
// of the sub-group, D = 64 the first 4.  A warp covers 4 heads of one token.
template <int D, bool KV16>
__global__ void __launch_bounds__(256) rope_kv_append_kernel(RopeKvParams p) {
    constexpr int HALF = D / 2;
    constexpr int CHUNKS = HALF / 8;  // 16-byte chunks per half: 8 (D = 128) or 4 (D = 64)
    pdl_trigger();
    pdl_wait();
    const int sub = threadIdx.x & 7;
    const int heads = p.nq + 2 * p.nkv;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;  // (token, head) index
    if (w >= p.num_tokens * heads || sub >= CHUNKS) return;
    const int64_t t = w / heads;
    const int h = (int)(w - t * heads);
    // decode sequences come first with one token each (llm_generator.cc:229-242): token t < decoding_batches
    // belongs to sequence t; prefill tokens need the search
    const int b = t < p.decoding_batches ? (int)t : find_seq(p.seq_starts, p.batch, t);
    const int64_t pos = p.start_pos[b] + (t - p.seq_starts[b]);

    __half* row = p.qkv + (t * heads + h) * (int64_t)D;
    const uint4 ulo = *reinterpret_cast<const uint4*>(row + 8 * sub);
    const uint4 uhi = *reinterpret_cast<const uint4*>(row + HALF + 8 * sub);
    float lo[8], hi[8];
    {
        const __half2* a = reinterpret_cast<const __half2*>(&ulo);
        const __half2* c = reinterpret_cast<const __half2*>(&uhi);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 fa = __half22float2(a[i]), fc = __half22float2(c[i]);
            lo[2 * i] = fa.x; lo[2 * i + 1] = fa.y; hi[2 * i] = fc.x; hi[2 * i + 1] = fc.y;
        }
    }

    const bool is_v = h >= p.nq + p.nkv;
    if (!is_v) {
        const float4* cp = reinterpret_cast<const float4*>(p.cos_t + pos * HALF + 8 * sub);
        const float4* sp = reinterpret_cast<const float4*>(p.sin_t + pos * HALF + 8 * sub);
        const float4 c0 = cp[0], c1 = cp[1], s0 = sp[0], s1 = sp[1];
        const float cs[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        const float sn[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float o1 = __fsub_rn(__fmul_rn(lo[i], cs[i]), __fmul_rn(hi[i], sn[i]));
            const float o2 = __fadd_rn(__fmul_rn(hi[i], cs[i]), __fmul_rn(lo[i], sn[i]));
            lo[i] = __half2float(__float2half_rn(o1));
            hi[i] = __half2float(__float2half_rn(o2));
        }
        uint4 olo, ohi;
        __half2* a = reinterpret_cast<__half2*>(&olo);
        __half2* c = reinterpret_cast<__half2*>(&ohi);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            a[i] = __floats2half2_rn(lo[2 * i], lo[2 * i + 1]);
            c[i] = __floats2half2_rn(hi[2 * i], hi[2 * i + 1]);
        }
        *reinterpret_cast<uint4*>(row + 8 * sub) = olo;
        *reinterpret_cast<uint4*>(row + HALF + 8 * sub) = ohi;
    }
    if (h < p.nq) return;

    // ---- quantised append: this lane's two chunks are two whole quantisation groups
    const int kv = is_v ? 1 : 0;
    const int hk = h - p.nq - kv * p.nkv;
    const int64_t slot = kv_slot(p.cache_indices, p.cache_mode, p.page_size, p.max_pages, b, pos);
    const int64_t off = kv * p.cs.kv + hk * p.cs.head + slot * p.cs.tok;
    if constexpr (KV16) {  // fp16 cache: the (rotated) values as they are; lo / hi were re-rounded to fp16 above
        __half* crow16 = reinterpret_cast<__half*>(p.cache) + off;
        uint4 olo, ohi;
        __half2* a = reinterpret_cast<__half2*>(&olo);
        __half2* c = reinterpret_cast<__half2*>(&ohi);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            a[i] = __floats2half2_rn(lo[2 * i], lo[2 * i + 1]);
            c[i] = __floats2half2_rn(hi[2 * i], hi[2 * i + 1]);
        }
        *reinterpret_cast<uint4*>(crow16 + 8 * sub) = olo;
        *reinterpret_cast<uint4*>(crow16 + HALF + 8 * sub) = ohi;
        return;
    }
    int8_t* crow = p.cache + off;
    __half* srow = p.scale + off / 8;

    float mlo = 0.f, mhi = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        mlo = fmaxf(mlo, fabsf(lo[i]));
        mhi = fmaxf(mhi, fabsf(hi[i]));
    }
    const __half slo16 = __float2half_rn(__fdiv_rn(mlo, 127.0f)), shi16 = __float2half_rn(__fdiv_rn(mhi, 127.0f));
    const float slo = __half2float(slo16), shi = __half2float(shi16);
    uint32_t wlo[2] = {0u, 0u}, whi[2] = {0u, 0u};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int ql = slo > 0.f ? max(-127, min(127, __float2int_rn(__fdiv_rn(lo[i], slo)))) : 0;
        const int qh = shi > 0.f ? max(-127, min(127, __float2int_rn(__fdiv_rn(hi[i], shi)))) : 0;
        wlo[i >> 2] |= (uint32_t)(ql & 0xff) << (8 * (i & 3));
        whi[i >> 2] |= (uint32_t)(qh & 0xff) << (8 * (i & 3));
    }
    *reinterpret_cast<uint2*>(crow + 8 * sub) = make_uint2(wlo[0], wlo[1]);
    *reinterpret_cast<uint2*>(crow + HALF + 8 * sub) = make_uint2(whi[0], whi[1]);
    srow[sub] = slo16;
    srow[CHUNKS + sub] = shi16;
}

}  // namespace

int32_t launch_rope_kv_append(cudaStream_t s, __half* qkv, const b2llm_step* step, int num_heads,
                              const b2llm_kv_geom& geom, int layer, const float* cos_t, const float* sin_t,
                              int8_t* kv_cache, __half* kv_scale) {
    B2_REQUIRE(geom.quant_group == 8 || geom.quant_group == 1, B2LLM_ERR_UNSUPPORTED,
               "kv cache: int8 with quant group 8, or fp16 (quant group 1)");
    const bool kv16 = geom.quant_group == 1;
    B2_REQUIRE(kv16 || kv_scale != nullptr, B2LLM_ERR_INVALID_VALUE, "kv cache: the int8 cache needs its scale tensor");
    B2_REQUIRE(geom.head_dim == 128 || geom.head_dim == 64, B2LLM_ERR_UNSUPPORTED, "head_dim must be 64 or 128");
    if (step->num_tokens == 0) return B2LLM_OK;
    RopeKvParams p{};
    p.qkv = qkv;
    p.seq_starts = step->seq_starts;
    p.start_pos = step->start_pos;
    p.cache_indices = step->cache_indices;
    p.batch = (int)step->batch;
    p.num_tokens = step->num_tokens;
    p.decoding_batches = step->decoding_batches;
    p.max_pages = step->max_pages;
    p.nq = num_heads;
    p.nkv = geom.num_kv_heads;
    p.D = geom.head_dim;
    p.cache_mode = geom.cache_mode;
    p.page_size = geom.page_size;
    p.group = geom.quant_group;
    p.cos_t = cos_t;
    p.sin_t = sin_t;
    p.cs = kv_strides(geom);
    p.cache = kv_cache + (int64_t)layer * p.cs.layer * (kv16 ? 2 : 1);
    p.scale = kv16 ? nullptr : kv_scale + (int64_t)layer * p.cs.layer / geom.quant_group;
    const int64_t threads = step->num_tokens * (num_heads + 2 * geom.num_kv_heads) * 8;
    const unsigned blocks = (unsigned)((threads + 255) / 256);
    if (geom.head_dim == 128) {
        if (kv16) launch_kernel(rope_kv_append_kernel<128, true>, dim3(blocks), dim3(256), 0, s, p);
        else launch_kernel(rope_kv_append_kernel<128, false>, dim3(blocks), dim3(256), 0, s, p);
    } else {
        if (kv16) launch_kernel(rope_kv_append_kernel<64, true>, dim3(blocks), dim3(256), 0, s, p);
        else launch_kernel(rope_kv_append_kernel<64, false>, dim3(blocks), dim3(256), 0, s, p);
    }
    B2_LAUNCH_CHECK();
    return B2LLM_OK;
}

}  // namespace b2llm
