// K1/K2 of SURVEY.md section 2.3: embedding gather, (skip-)RMSNorm + per-token int8 quantisation,
// per-token int8 quantisation of plain rows.  HBM-bound row kernels: one CTA per row, 128-bit
// loads, warp-shuffle + shared-memory reductions, the row stays in L1 between passes.
//
// Numeric contract (oracle/llama_ref.py: rmsnorm_f32, quant_rows):
//   var = mean(x^2) fp32;  inv = 1 / sqrt(var + eps);  y = (x * inv) * gamma;
//   amax = max|y|;  scale = amax / 127;  q = clamp(rint(y * (127 / amax)), -127, 127).
#include "common.cuh"

namespace b2llm {

namespace {

constexpr int kThreads = 256;

struct alignas(16) Half8 {
    __half2 v[4];
};

__device__ __forceinline__ Half8 ld8(const __half* p) { return *reinterpret_cast<const Half8*>(p); }
__device__ __forceinline__ void st8(__half* p, const Half8& v) { *reinterpret_cast<Half8*>(p) = v; }

__device__ __forceinline__ double block_sum_f64(double v, double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double r = (lane < nw) ? scratch[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    return r;
}

__device__ __forceinline__ int q8(float y, float inv_scale) {
    int q = __float2int_rn(__fmul_rn(y, inv_scale));
    return max(-127, min(127, q));
}

// x: [rows, hidden] fp16 (updated in place when skip != nullptr)
__global__ void __launch_bounds__(kThreads) rmsnorm_quant_kernel(__half* __restrict__ x, const __half* __restrict__ skip,
                                                                const __half* __restrict__ gamma, float eps, int hidden,
                                                                int8_t* __restrict__ q, float* __restrict__ scale,
                                                                __half* __restrict__ y_out) {
    __shared__ float scratch[32];
    __shared__ double dscratch[32];
    pdl_trigger();
    pdl_wait();
    const int64_t row = blockIdx.x;
    __half* xr = x + row * hidden;
    const int nvec = hidden >> 3;

    // pass 1: residual join (optional) + sum of squares.  Accumulated in fp64 so that the variance is
    // independent of the reduction order (the oracle uses float64 too) and the int8 codes downstream
    // are reproducible bit for bit.
    double ss = 0.0;
    for (int v = threadIdx.x; v < nvec; v += kThreads) {
        Half8 a = ld8(xr + v * 8);
        if (skip != nullptr) {
            const Half8 b = ld8(skip + row * hidden + v * 8);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 fa = __half22float2(a.v[i]), fb = __half22float2(b.v[i]);
                a.v[i] = __floats2half2_rn(__fadd_rn(fa.x, fb.x), __fadd_rn(fa.y, fb.y));
            }
            st8(xr + v * 8, a);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(a.v[i]);
            ss += (double)f.x * (double)f.x + (double)f.y * (double)f.y;
        }
    }
    ss = block_sum_f64(ss, dscratch);
    const float var = (float)(ss / (double)hidden);
    const float inv = __fdiv_rn(1.0f, sqrtf(__fadd_rn(var, eps)));

    if (q == nullptr) {  // plain fp16 output
        for (int v = threadIdx.x; v < nvec; v += kThreads) {
            const Half8 a = ld8(xr + v * 8), g = ld8(gamma + v * 8);
            Half8 o;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 fa = __half22float2(a.v[i]), fg = __half22float2(g.v[i]);
                o.v[i] = __floats2half2_rn(__fmul_rn(__fmul_rn(fa.x, inv), fg.x), __fmul_rn(__fmul_rn(fa.y, inv), fg.y));
            }
            st8(y_out + row * hidden + v * 8, o);
        }
        return;
    }

    // pass 2: row max of |y|
    float amax = 0.f;
    for (int v = threadIdx.x; v < nvec; v += kThreads) {
        const Half8 a = ld8(xr + v * 8), g = ld8(gamma + v * 8);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 fa = __half22float2(a.v[i]), fg = __half22float2(g.v[i]);
            amax = fmaxf(amax, fabsf(__fmul_rn(__fmul_rn(fa.x, inv), fg.x)));
            amax = fmaxf(amax, fabsf(__fmul_rn(__fmul_rn(fa.y, inv), fg.y)));
        }
    }
    amax = block_max(amax, scratch);
    const float inv_scale = amax > 0.f ? __fdiv_rn(127.0f, amax) : 0.f;
    if (threadIdx.x == 0) scale[row] = __fdiv_rn(amax, 127.0f);

    // pass 3: quantise, 8 bytes per store
    for (int v = threadIdx.x; v < nvec; v += kThreads) {
        const Half8 a = ld8(xr + v * 8), g = ld8(gamma + v * 8);
        uint32_t w[2] = {0, 0};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 fa = __half22float2(a.v[i]), fg = __half22float2(g.v[i]);
            const int q0 = q8(__fmul_rn(__fmul_rn(fa.x, inv), fg.x), inv_scale);
            const int q1 = q8(__fmul_rn(__fmul_rn(fa.y, inv), fg.y), inv_scale);
            w[i >> 1] |= (uint32_t)(q0 & 0xff) << (16 * (i & 1));
            w[i >> 1] |= (uint32_t)(q1 & 0xff) << (16 * (i & 1) + 8);
        }
        *reinterpret_cast<uint2*>(q + row * hidden + v * 8) = make_uint2(w[0], w[1]);
    }
}

__global__ void __launch_bounds__(kThreads) quant_rows_kernel(const __half* __restrict__ x, int cols,
                                                             int8_t* __restrict__ q, float* __restrict__ scale) {
    __shared__ float scratch[32];
    pdl_trigger();
    pdl_wait();
    const int64_t row = blockIdx.x;
    const __half* xr = x + row * cols;
    const int nvec = cols >> 3;
    float amax = 0.f;
    for (int v = threadIdx.x; v < nvec; v += kThreads) {
        const Half8 a = ld8(xr + v * 8);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(a.v[i]);
            amax = fmaxf(amax, fmaxf(fabsf(f.x), fabsf(f.y)));
        }
    }
    amax = block_max(amax, scratch);
    const float inv_scale = amax > 0.f ? __fdiv_rn(127.0f, amax) : 0.f;
    if (threadIdx.x == 0) scale[row] = __fdiv_rn(amax, 127.0f);
    for (int v = threadIdx.x; v < nvec; v += kThreads) {
        const Half8 a = ld8(xr + v * 8);
        uint32_t w[2] = {0, 0};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(a.v[i]);
            w[i >> 1] |= (uint32_t)(q8(f.x, inv_scale) & 0xff) << (16 * (i & 1));
            w[i >> 1] |= (uint32_t)(q8(f.y, inv_scale) & 0xff) << (16 * (i & 1) + 8);
        }
        *reinterpret_cast<uint2*>(q + row * cols + v * 8) = make_uint2(w[0], w[1]);
    }
}

// ------------------------------------------------------------------------------------------
// Register-resident row kernels.  The CTA-per-row kernels above re-read the row for every pass and synchronise
// the block four times; at 1024 rows x 4096 columns that costs ~11 us for 12 MB (profiles/r1_ncu_full_step_run11.txt).
// Here W warps own one row (W = 1: no block synchronisation at all), every lane keeps its V 16-byte vectors of the
// row in registers across the passes, and reductions are warp shuffles (+ one shared-memory exchange when W > 1).
// Same arithmetic, same fp64 variance, same rounding points as the kernels above -> bit-identical results.
template <int W>
__device__ __forceinline__ float rowgroup_max(float v, float* scratch, int w) {
    v = warp_max(v);
    if constexpr (W > 1) {
        __syncthreads();
        if ((threadIdx.x & 31) == 0) scratch[w] = v;
        __syncthreads();
        v = scratch[0];
#pragma unroll
        for (int i = 1; i < W; ++i) v = fmaxf(v, scratch[i]);
    }
    return v;
}
template <int W>
__device__ __forceinline__ double rowgroup_sum_f64(double v, double* scratch, int w) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if constexpr (W > 1) {
        __syncthreads();
        if ((threadIdx.x & 31) == 0) scratch[w] = v;
        __syncthreads();
        v = scratch[0];
#pragma unroll
        for (int i = 1; i < W; ++i) v += scratch[i];
    }
    return v;
}

// V vectors per lane, W warps per row.  W == 1: a 128-thread CTA handles 4 rows; W > 1: a CTA of W warps handles one row.
template <int V, int W, bool NORM>
__global__ void __launch_bounds__(W == 1 ? 128 : 32 * W)
    row_quant_reg_kernel(__half* __restrict__ x, const __half* __restrict__ skip, const __half* __restrict__ gamma, float eps,
                         int64_t rows, int cols, int8_t* __restrict__ q, float* __restrict__ scale, __half* __restrict__ y_out) {
    __shared__ float fscratch[W > 1 ? W : 1];
    __shared__ double dscratch[W > 1 ? W : 1];
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = W == 1 ? (int64_t)blockIdx.x * 4 + warp : (int64_t)blockIdx.x;
    const int w = W == 1 ? 0 : warp;
    if (row >= rows) return;  // W == 1 only: whole warps leave, no block-level sync is used in that configuration
    const int nvec = cols >> 3;
    __half* xr = x + row * cols;
    Half8 a[V];
    double ss = 0.0;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const int v = i * (32 * W) + w * 32 + lane;
        if (v < nvec) {
            a[i] = ld8(xr + v * 8);
            if constexpr (NORM) {
                if (skip != nullptr) {
                    const Half8 b = ld8(skip + row * cols + v * 8);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 fa = __half22float2(a[i].v[j]), fb = __half22float2(b.v[j]);
                        a[i].v[j] = __floats2half2_rn(__fadd_rn(fa.x, fb.x), __fadd_rn(fa.y, fb.y));
                    }
                    st8(xr + v * 8, a[i]);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __half22float2(a[i].v[j]);
                    ss += (double)f.x * (double)f.x + (double)f.y * (double)f.y;
                }
            }
        }
    }
    float inv = 1.f;
    if constexpr (NORM) {
        ss = rowgroup_sum_f64<W>(ss, dscratch, w);
        const float var = (float)(ss / (double)cols);
        inv = __fdiv_rn(1.0f, sqrtf(__fadd_rn(var, eps)));
    }
    // y = (x * inv) * gamma kept as fp32 pairs only transiently: recomputed per pass from the register-resident x
    auto yv = [&](const Half8& xa, const Half8& g, int j) -> float2 {
        const float2 fa = __half22float2(xa.v[j]);
        if constexpr (NORM) {
            const float2 fg = __half22float2(g.v[j]);
            return make_float2(__fmul_rn(__fmul_rn(fa.x, inv), fg.x), __fmul_rn(__fmul_rn(fa.y, inv), fg.y));
        } else {
            return fa;
        }
    };
    if (NORM && q == nullptr) {  // plain fp16 output
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int v = i * (32 * W) + w * 32 + lane;
            if (v < nvec) {
                const Half8 g = ld8(gamma + v * 8);
                Half8 o;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = yv(a[i], g, j);
                    o.v[j] = __floats2half2_rn(f.x, f.y);
                }
                st8(y_out + row * cols + v * 8, o);
            }
        }
        return;
    }
    float amax = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const int v = i * (32 * W) + w * 32 + lane;
        if (v < nvec) {
            Half8 g;
            if constexpr (NORM) g = ld8(gamma + v * 8);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = yv(a[i], g, j);
                amax = fmaxf(amax, fmaxf(fabsf(f.x), fabsf(f.y)));
            }
        }
    }
    amax = rowgroup_max<W>(amax, fscratch, w);
    const float inv_scale = amax > 0.f ? __fdiv_rn(127.0f, amax) : 0.f;
    if (lane == 0 && w == 0) scale[row] = __fdiv_rn(amax, 127.0f);
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const int v = i * (32 * W) + w * 32 + lane;
        if (v < nvec) {
            Half8 g;
            if constexpr (NORM) g = ld8(gamma + v * 8);
            uint32_t wd[2] = {0, 0};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = yv(a[i], g, j);
                wd[j >> 1] |= (uint32_t)(q8(f.x, inv_scale) & 0xff) << (16 * (j & 1));
                wd[j >> 1] |= (uint32_t)(q8(f.y, inv_scale) & 0xff) << (16 * (j & 1) + 8);
            }
            *reinterpret_cast<uint2*>(q + row * cols + v * 8) = make_uint2(wd[0], wd[1]);
        }
    }
}

template <bool NORM>
bool launch_row_reg(cudaStream_t s, __half* x, const __half* skip, const __half* gamma, float eps, int64_t rows, int cols,
                    int8_t* q, float* scale, __half* y) {
    const int nvec = cols >> 3;
    // warps per row: enough rows-in-flight parallelism matters more than avoiding the block sync -- one warp per
    // 4096-column row measured SLOWER in the step than the CTA-per-row kernels (run 12: 7 warps per SM, long dependent
    // chains), so wide rows are spread over 4 warps
    const int W = nvec < 256 ? 1 : (nvec < 512 ? 2 : 4);
    const int need = (nvec + 32 * W - 1) / (32 * W);
    if (need > 32) return false;  // > 32768 columns: CTA-per-row kernel
    const unsigned blocks = W == 1 ? (unsigned)((rows + 3) / 4) : (unsigned)rows;
#define B2_ROW_LAUNCH(VV, WW) launch_kernel(row_quant_reg_kernel<VV, WW, NORM>, dim3(blocks), dim3(WW == 1 ? 128 : 32 * WW), 0, s, x, skip, gamma, eps, rows, cols, q, scale, y)
#define B2_ROW_V(WW)                                           \
    if (need <= 4) B2_ROW_LAUNCH(4, WW);                       \
    else if (need <= 8) B2_ROW_LAUNCH(8, WW);                  \
    else if (need <= 16) B2_ROW_LAUNCH(16, WW);                \
    else if (need <= 24) B2_ROW_LAUNCH(24, WW);                \
    else B2_ROW_LAUNCH(32, WW)
    if (W == 1) { B2_ROW_V(1); } else if (W == 2) { B2_ROW_V(2); } else { B2_ROW_V(4); }
#undef B2_ROW_V
#undef B2_ROW_LAUNCH
    return true;
}

// out[i, :] = table[ids[i], :]
__global__ void embedding_kernel(const int64_t* __restrict__ ids, const __half* __restrict__ table, int hidden, int vocab,
                                 __half* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    const int64_t i = blockIdx.x;
    int64_t id = ids[i];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const int nvec = hidden >> 3;
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) st8(out + i * hidden + v * 8, ld8(table + id * hidden + v * 8));
}

// out[b, :] = x[seq_starts[b + 1] - 1, :]   (last token of every sequence)
__global__ void gather_rows_kernel(const __half* __restrict__ x, const int64_t* __restrict__ seq_starts, int hidden,
                                   __half* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    const int64_t b = blockIdx.x;
    const int64_t t = seq_starts[b + 1] - 1;
    const int nvec = hidden >> 3;
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) st8(out + b * hidden + v * 8, ld8(x + t * hidden + v * 8));
}

// all-gathered logits [parts, rows, cols] (rank-major, what ncclAllGather delivers) -> [rows, parts * cols]
__global__ void __launch_bounds__(256) interleave_blocks_kernel(const float4* __restrict__ src, int parts, int64_t rows, int cols4,
                                                               float4* __restrict__ dst) {
    pdl_trigger();
    pdl_wait();
    const int64_t n = (int64_t)parts * rows * cols4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = i % cols4, r = (i / cols4) % rows, p = i / (cols4 * rows);
        dst[(r * parts + p) * cols4 + c] = src[i];
    }
}

// Which row kernel: the register-resident warp-group-per-row variants from 256 rows up -- at prefill sizes they save
// 16 % of the norm / quant / rope time (65 536 rows: 74 vs 89 ms per 32-layer step, round 1 run 16), at the decode size
// (1024 rows) 0.4 % of the step (37.36 vs 37.52 ms, same box back to back, round 2 run 16; round 1 had measured them
// slower there, before PDL); tiny steps keep the CTA-per-row kernels.  Results are bit-identical (tests cover both at
// both sizes).  B2LLM_ROW_KERNELS=reg / legacy forces one of them.
bool use_row_reg(int64_t rows) {
    static const int mode = [] {
        const char* e = getenv("B2LLM_ROW_KERNELS");
        return e == nullptr ? 0 : (e[0] == 'r' ? 1 : (e[0] == 'l' ? 2 : 0));
    }();
    return mode == 1 || (mode == 0 && rows >= 256);
}

}  // namespace

int32_t launch_rmsnorm_quant(cudaStream_t s, __half* x, const __half* skip, const __half* gamma, float eps, int64_t rows,
                             int hidden, int8_t* q, float* scale, __half* y) {
    B2_REQUIRE(hidden % 8 == 0, B2LLM_ERR_INVALID_VALUE, "rmsnorm: hidden must be a multiple of 8");
    B2_REQUIRE((q != nullptr && scale != nullptr) || y != nullptr, B2LLM_ERR_INVALID_VALUE, "rmsnorm: no output");
    if (rows == 0) return B2LLM_OK;
    if (use_row_reg(rows) && launch_row_reg<true>(s, x, skip, gamma, eps, rows, hidden, q, scale, y)) {
        B2_LAUNCH_CHECK();
        return B2LLM_OK;
    }
    launch_kernel(rmsnorm_quant_kernel, dim3((unsigned)rows), dim3(kThreads), 0, s, x, skip, gamma, eps, hidden, q, scale, y);
    B2_LAUNCH_CHECK();
    return B2LLM_OK;
}

int32_t launch_quant_rows(cudaStream_t s, const __half* x, int64_t rows, int cols, int8_t* q, float* scale) {
    B2_REQUIRE(cols % 8 == 0, B2LLM_ERR_INVALID_VALUE, "quant_rows: cols must be a multiple of 8");
    if (rows == 0) return B2LLM_OK;
    if (use_row_reg(rows) && launch_row_reg<false>(s, const_cast<__half*>(x), nullptr, nullptr, 0.f, rows, cols, q, scale, nullptr)) {
        B2_LAUNCH_CHECK();
        return B2LLM_OK;
    }
    launch_kernel(quant_rows_kernel, dim3((unsigned)rows), dim3(kThreads), 0, s, x, cols, q, scale);
    B2_LAUNCH_CHECK();
    return B2LLM_OK;
}

int32_t launch_embedding(cudaStream_t s, const int64_t* ids, const __half* table, int64_t n, int hidden, int vocab,
                         __half* out) {
    if (n == 0) return B2LLM_OK;
    launch_kernel(embedding_kernel, dim3((unsigned)n), dim3(128), 0, s, ids, table, hidden, vocab, out);
    B2_LAUNCH_CHECK();
    return B2LLM_OK;
}

int32_t launch_gather_rows(cudaStream_t s, const __half* x, const int64_t* seq_starts, int64_t batch, int hidden,
                           __half* out) {
    if (batch == 0) return B2LLM_OK;
    launch_kernel(gather_rows_kernel, dim3((unsigned)batch), dim3(128), 0, s, x, seq_starts, hidden, out);
    B2_LAUNCH_CHECK();
    return B2LLM_OK;
}

int32_t launch_interleave_blocks(cudaStream_t s, const float* src, int parts, int64_t rows, int cols, float* dst) {
    B2_REQUIRE(cols % 4 == 0, B2LLM_ERR_INVALID_VALUE, "interleave_blocks: cols must be a multiple of 4");
    const int64_t n = (int64_t)parts * rows * (cols / 4);
    if (n == 0) return B2LLM_OK;
    const unsigned blocks = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8);
    launch_kernel(interleave_blocks_kernel, dim3(blocks), dim3(256), 0, s, reinterpret_cast<const float4*>(src), parts, rows, cols / 4,
                  reinterpret_cast<float4*>(dst));
    B2_LAUNCH_CHECK();
    return B2LLM_OK;
}

}  // namespace b2llm
