// K4 + K5 of SURVEY.md section 2.3: rotary embedding on q,k (in place in the qkv activation) and the
// int8 group-8 quantised append of k,v into the KV cache, for every token of the step.
// One warp per (token, head); 16-byte accesses; group max via 4-lane shuffles.
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

// D = 128: lane l owns dims [4l, 4l+4) of the low half when l < 16 ... simpler: lane l owns the pair
// columns i in {2l, 2l+1} of each half: elements (2l, 2l+1) and (D/2 + 2l, D/2 + 2l + 1).
template <int D>
__global__ void __launch_bounds__(128) rope_kv_append_kernel(RopeKvParams p) {
    constexpr int HALF = D / 2;
    constexpr int PER = HALF / 32;  // pair-columns per lane (2 for D = 128, 1 for D = 64)
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int heads = p.nq + 2 * p.nkv;
    if (w >= p.num_tokens * heads) return;
    const int64_t t = w / heads;
    const int h = (int)(w % heads);
    const int b = find_seq(p.seq_starts, p.batch, t);
    const int64_t pos = p.start_pos[b] + (t - p.seq_starts[b]);

    __half* row = p.qkv + t * (int64_t)heads * D + (int64_t)h * D;
    float lo[PER], hi[PER];
    if constexpr (PER == 2) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(row + 2 * lane));
        const float2 c = __half22float2(*reinterpret_cast<const __half2*>(row + HALF + 2 * lane));
        lo[0] = a.x; lo[1] = a.y; hi[0] = c.x; hi[1] = c.y;
    } else {
        lo[0] = __half2float(row[lane]);
        hi[0] = __half2float(row[HALF + lane]);
    }

    const bool is_v = h >= p.nq + p.nkv;
    if (!is_v) {
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int col = PER * lane + i;
            const float c = p.cos_t[pos * HALF + col], s = p.sin_t[pos * HALF + col];
            const float o1 = __fsub_rn(__fmul_rn(lo[i], c), __fmul_rn(hi[i], s));
            const float o2 = __fadd_rn(__fmul_rn(hi[i], c), __fmul_rn(lo[i], s));
            lo[i] = __half2float(__float2half_rn(o1));
            hi[i] = __half2float(__float2half_rn(o2));
        }
        if constexpr (PER == 2) {
            *reinterpret_cast<__half2*>(row + 2 * lane) = __floats2half2_rn(lo[0], lo[1]);
            *reinterpret_cast<__half2*>(row + HALF + 2 * lane) = __floats2half2_rn(hi[0], hi[1]);
        } else {
            row[lane] = __float2half_rn(lo[0]);
            row[HALF + lane] = __float2half_rn(hi[0]);
        }
    }
    if (h < p.nq) return;

    // ---- quantised append.  group of 8 dims = 8 / PER consecutive lanes in each half.
    const int kv = is_v ? 1 : 0;
    const int hk = h - p.nq - kv * p.nkv;
    const int64_t slot = kv_slot(p.cache_indices, p.cache_mode, p.page_size, p.max_pages, b, pos);
    int8_t* crow = p.cache + kv * p.cs.kv + hk * p.cs.head + slot * p.cs.tok;
    __half* srow = p.scale + (kv * p.cs.kv + hk * p.cs.head + slot * p.cs.tok) / p.group;

    float mlo = 0.f, mhi = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        mlo = fmaxf(mlo, fabsf(lo[i]));
        mhi = fmaxf(mhi, fabsf(hi[i]));
    }
    constexpr int LANES_PER_GROUP = 8 / PER;
#pragma unroll
    for (int o = 1; o < LANES_PER_GROUP; o <<= 1) {
        mlo = fmaxf(mlo, __shfl_xor_sync(0xffffffffu, mlo, o));
        mhi = fmaxf(mhi, __shfl_xor_sync(0xffffffffu, mhi, o));
    }
    const __half slo16 = __float2half_rn(__fdiv_rn(mlo, 127.0f)), shi16 = __float2half_rn(__fdiv_rn(mhi, 127.0f));
    const float slo = __half2float(slo16), shi = __half2float(shi16);
    int qlo[PER], qhi[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        qlo[i] = slo > 0.f ? max(-127, min(127, __float2int_rn(__fdiv_rn(lo[i], slo)))) : 0;
        qhi[i] = shi > 0.f ? max(-127, min(127, __float2int_rn(__fdiv_rn(hi[i], shi)))) : 0;
    }
    if constexpr (PER == 2) {
        *reinterpret_cast<uint16_t*>(crow + 2 * lane) = (uint16_t)((qlo[0] & 0xff) | ((qlo[1] & 0xff) << 8));
        *reinterpret_cast<uint16_t*>(crow + HALF + 2 * lane) = (uint16_t)((qhi[0] & 0xff) | ((qhi[1] & 0xff) << 8));
    } else {
        crow[lane] = (int8_t)qlo[0];
        crow[HALF + lane] = (int8_t)qhi[0];
    }
    if ((lane % LANES_PER_GROUP) == 0) {
        const int gidx = lane / LANES_PER_GROUP;  // group index inside the half
        srow[gidx] = slo16;
        srow[HALF / 8 + gidx] = shi16;
    }
}

}  // namespace

int32_t launch_rope_kv_append(cudaStream_t s, __half* qkv, const b2llm_step* step, int num_heads,
                              const b2llm_kv_geom& geom, int layer, const float* cos_t, const float* sin_t,
                              int8_t* kv_cache, __half* kv_scale) {
    B2_REQUIRE(geom.quant_group == 8, B2LLM_ERR_UNSUPPORTED, "kv cache: only int8 with quant group 8 is supported");
    B2_REQUIRE(geom.head_dim == 128 || geom.head_dim == 64, B2LLM_ERR_UNSUPPORTED, "head_dim must be 64 or 128");
    if (step->num_tokens == 0) return B2LLM_OK;
    RopeKvParams p{};
    p.qkv = qkv;
    p.seq_starts = step->seq_starts;
    p.start_pos = step->start_pos;
    p.cache_indices = step->cache_indices;
    p.batch = (int)step->batch;
    p.num_tokens = step->num_tokens;
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
    p.cache = kv_cache + (int64_t)layer * p.cs.layer;
    p.scale = kv_scale + (int64_t)layer * p.cs.layer / geom.quant_group;
    const int64_t warps = step->num_tokens * (num_heads + 2 * geom.num_kv_heads);
    const unsigned blocks = (unsigned)((warps + 3) / 4);
    if (geom.head_dim == 128)
        rope_kv_append_kernel<128><<<blocks, 128, 0, s>>>(p);
    else
        rope_kv_append_kernel<64><<<blocks, 128, 0, s>>>(p);
    B2_LAUNCH_CHECK();
    return B2LLM_OK;
}

}  // namespace b2llm
