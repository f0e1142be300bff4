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

int32_t launch_interleave_rows(cudaStream_t s, const __half* a, const __half* b, int rows, int cols, __half* out) {
    if (rows == 0) return B2LLM_OK;
    interleave_rows_kernel<<<rows, 256, 0, s>>>(a, b, cols, out);
    B2_LAUNCH_CHECK();
    return B2LLM_OK;
}

}  // namespace b2llm
