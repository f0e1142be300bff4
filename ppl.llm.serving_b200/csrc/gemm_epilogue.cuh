// Epilogues shared by the mma.sync and tcgen05 GEMMs.  Values arrive as (m, n) / (m, n + 1) column
// pairs.  Numeric contract: oracle/llama_ref.py dequant_acc / silu_mul / the residual add in
// LlamaOracle.forward.
#pragma once
#include "common.cuh"

namespace b2llm {

enum { EPI_F16 = 0, EPI_RESIDUAL = 1, EPI_SWIGLU = 2, EPI_F32 = 3 };

template <bool I8>
struct AccT {
    using type = float;
};
template <>
struct AccT<true> {
    using type = int;
};

#ifdef __CUDACC__
// (float(acc) * a_scale[m]) * w_scale[n]: two individually rounded fp32 multiplies
__device__ __forceinline__ float dequant(int acc, float sa, float sw) {
    return __fmul_rn(__fmul_rn(__int2float_rn(acc), sa), sw);
}

// ldc: row stride of `out` in elements (fp16 for EPI 0/1/2 -- for SWIGLU the output has N/2 columns)
template <int EPI>
__device__ __forceinline__ void store_pair(void* out, int64_t ldc, int N, int m, int n, float v0, float v1) {
    if constexpr (EPI == EPI_F16) {
        *reinterpret_cast<__half2*>(reinterpret_cast<__half*>(out) + (int64_t)m * ldc + n) = __floats2half2_rn(v0, v1);
    } else if constexpr (EPI == EPI_RESIDUAL) {
        __half2* p = reinterpret_cast<__half2*>(reinterpret_cast<__half*>(out) + (int64_t)m * ldc + n);
        const float2 r = __half22float2(*p);
        *p = __floats2half2_rn(__fadd_rn(r.x, v0), __fadd_rn(r.y, v1));
    } else if constexpr (EPI == EPI_SWIGLU) {
        // interleaved weight rows: even column = gate_j, odd column = up_j
        reinterpret_cast<__half*>(out)[(int64_t)m * ldc + (n >> 1)] = __float2half_rn(silu_mul_f32(v0, v1));
    } else {
        *reinterpret_cast<float2*>(reinterpret_cast<float*>(out) + (int64_t)m * ldc + n) = make_float2(v0, v1);
    }
}
#endif

}  // namespace b2llm
