// Weight preparation: the seeded synthetic generator (mirror of oracle/weights.py, same integer hash),
// the "online" per-output-channel int8 quantisation of projection weights
// (quant_method online_i8i8, src/backends/cuda/resource_manager.cc:51-52) and the gate/up row
// interleave that lets the SwiGLU epilogue see (gate_j, up_j) as one accumulator column pair.
#include "common.cuh"

namespace b2llm {

namespace {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// out[r, c] (local [rows, cols]) = synth(full index (row0 + r) * full_cols + col0 + c)
__global__ void synth_fp16_kernel(uint64_t base, int64_t rows, int64_t cols, int64_t row0, int64_t col0, int64_t full_cols,
                                  float mul, float mean, __half* __restrict__ out) {
    const int64_t n = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols, c = i - r * cols;
        const uint64_t idx = (uint64_t)((row0 + r) * full_cols + col0 + c);
        const uint64_t z = splitmix64(base + idx);
        const int s = (int)(z & 0xFFFF) + (int)((z >> 16) & 0xFFFF) + (int)((z >> 32) & 0xFFFF) + (int)(z >> 48);
        const float f = __int2float_rn(s - 131070);
        out[i] = __float2half_rn(__fadd_rn(__fmul_rn(f, mul), mean));
    }
}

__global__ void __launch_bounds__(256) quant_weight_kernel(const __half* __restrict__ w, int K, int8_t* __restrict__ q,
                                                          float* __restrict__ scale) {
    __shared__ float scratch[32];
    const int64_t n = blockIdx.x;
    const __half* row = w + n * K;
    float amax = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) amax = fmaxf(amax, fabsf(__half2float(row[k])));
    amax = block_max(amax, scratch);
    const float inv = amax > 0.f ? __fdiv_rn(127.0f, amax) : 0.f;
    if (threadIdx.x == 0) scale[n] = __fdiv_rn(amax, 127.0f);
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const int v = __float2int_rn(__fmul_rn(__half2float(row[k]), inv));
        q[n * K + k] = (int8_t)max(-127, min(127, v));
    }
}

// W4A16 (builder-defined: the reference cannot select it, resource_manager.cc:49-56): symmetric int4, one fp16 scale
// per 128 consecutive K elements of an output channel; scale16 = fp16(max|w| / 7), q = clamp(rint(w / scale16), -7, 7),
// stored as the nibble q + 8, element k in the low (k even) / high (k odd) half of byte k / 2.
// One warp per (row, group): lane l owns elements [4 l, 4 l + 4) of the group.
__global__ void __launch_bounds__(256) quant_weight_w4_kernel(const __half* __restrict__ w, int64_t groups_total, int K,
                                                             uint8_t* __restrict__ packed, __half* __restrict__ scale) {
    const int lane = threadIdx.x & 31;
    const int64_t gi = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (gi >= groups_total) return;
    const int gpr = K / 128;
    const int64_t n = gi / gpr;
    const int g = (int)(gi - n * gpr);
    const __half* src = w + n * K + g * 128 + 4 * lane;
    float v[4];
    float amax = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[i] = __half2float(src[i]);
        amax = fmaxf(amax, fabsf(v[i]));
    }
    amax = warp_max(amax);
    const __half s16 = __float2half_rn(__fdiv_rn(amax, 7.0f));
    const float sf = __half2float(s16);
    uint32_t nib[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int q = sf > 0.f ? max(-7, min(7, __float2int_rn(__fdiv_rn(v[i], sf)))) : 0;
        nib[i] = (uint32_t)(q + 8);
    }
    uint8_t* dst = packed + (n * K + g * 128 + 4 * lane) / 2;
    dst[0] = (uint8_t)(nib[0] | (nib[1] << 4));
    dst[1] = (uint8_t)(nib[2] | (nib[3] << 4));
    if (lane == 0) scale[gi] = s16;
}

// fp16 [N, K] <- fp16(q * scale): the operand the fp16 tensor-core GEMM consumes.  16 bytes (32 weights) per thread.
__global__ void __launch_bounds__(256) dequant_w4_kernel(const uint8_t* __restrict__ packed, const __half* __restrict__ scale,
                                                        int64_t chunks, __half* __restrict__ out) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // chunk of 32 weights
    if (c >= chunks) return;
    const uint4 raw = *reinterpret_cast<const uint4*>(packed + c * 16);
    const float sf = __half2float(scale[c / 4]);  // 4 chunks per 128-element group
    const uint32_t wds[4] = {raw.x, raw.y, raw.z, raw.w};
    uint4 o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t* ow = reinterpret_cast<uint32_t*>(&o[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int q0 = (int)((wds[i] >> (8 * j)) & 0xF) - 8, q1 = (int)((wds[i] >> (8 * j + 4)) & 0xF) - 8;
            const __half2 h = __floats2half2_rn(__fmul_rn((float)q0, sf), __fmul_rn((float)q1, sf));
            ow[j] = *reinterpret_cast<const uint32_t*>(&h);
        }
    }
    uint4* dst = reinterpret_cast<uint4*>(out + c * 32);
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = o[i];
}

// out[2 r] = a[r], out[2 r + 1] = b[r]
__global__ void interleave_rows_kernel(const __half* __restrict__ a, const __half* __restrict__ b, int cols,
                                       __half* __restrict__ out) {
    const int64_t r = blockIdx.x;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
        out[(2 * r) * cols + c] = a[r * cols + c];
        out[(2 * r + 1) * cols + c] = b[r * cols + c];
    }
}

}  // namespace

uint64_t synth_base(uint64_t seed, uint64_t tid) {
    return seed * 0x9E3779B97F4A7C15ull + tid * 0xD1B54A32D192ED03ull;
}

int32_t launch_synth_fp16_2d(cudaStream_t s, uint64_t seed, uint64_t tid, int64_t rows, int64_t cols, int64_t row0,
                             int64_t col0, int64_t full_cols, float std, float mean, __half* out) {
    if (rows * cols == 0) return B2LLM_OK;
    const float mul = (float)((double)std / (65535.0 / sqrt(3.0)));
    const int64_t n = rows * cols;
    const unsigned blocks = (unsigned)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    synth_fp16_kernel<<<blocks, 256, 0, s>>>(synth_base(seed, tid), rows, cols, row0, col0, full_cols, mul, mean, out);
    B2_LAUNCH_CHECK();
    return B2LLM_OK;
}

int32_t launch_synth_fp16(cudaStream_t s, uint64_t seed, uint64_t tid, uint64_t n, float std, float mean, __half* out) {
    return launch_synth_fp16_2d(s, seed, tid, 1, (int64_t)n, 0, 0, (int64_t)n, std, mean, out);
}

int32_t launch_quant_weight(cudaStream_t s, const __half* w, int N, int K, int8_t* q, float* scale) {
    if (N == 0) return B2LLM_OK;
    quant_weight_kernel<<<N, 256, 0, s>>>(w, K, q, scale);
    B2_LAUNCH_CHECK();
    return B2LLM_OK;
}

int32_t launch_quant_weight_w4(cudaStream_t s, const __half* w, int N, int K, uint8_t* packed, __half* scale) {
    B2_REQUIRE(K % 128 == 0, B2LLM_ERR_UNSUPPORTED, "W4A16: K must be a multiple of the quantisation group (128)");
    if (N == 0) return B2LLM_OK;
    const int64_t groups = (int64_t)N * (K / 128);
    quant_weight_w4_kernel<<<(unsigned)((groups + 7) / 8), 256, 0, s>>>(w, groups, K, packed, scale);
    B2_LAUNCH_CHECK();
    return B2LLM_OK;
}

int32_t launch_dequant_w4(cudaStream_t s, const uint8_t* packed, const __half* scale, int N, int K, __half* out) {
    B2_REQUIRE(K % 128 == 0, B2LLM_ERR_UNSUPPORTED, "W4A16: K must be a multiple of the quantisation group (128)");
    if (N == 0) return B2LLM_OK;
    const int64_t chunks = (int64_t)N * K / 32;
    dequant_w4_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, s>>>(packed, scale, chunks, out);
    B2_LAUNCH_CHECK();
    return B2LLM_OK;
}

int32_t launch_interleave_rows(cudaStream_t s, const __half* a, const __half* b, int rows, int cols, __half* out) {
    if (rows == 0) return B2LLM_OK;
    interleave_rows_kernel<<<rows, 256, 0, s>>>(a, b, cols, out);
    B2_LAUNCH_CHECK();
    return B2LLM_OK;
}

}  // namespace b2llm
