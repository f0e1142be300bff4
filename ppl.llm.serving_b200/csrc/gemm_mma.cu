// Baseline tensor-core GEMM (legacy mma.sync path): C[M,N] = A[M,K] * W[N,K]^T for int8 (W8A8, exact
// int32 accumulation, K3/K8/K9/K10 of SURVEY.md section 2.3) and fp16 (lm_head K11, and the
// quant_method "none" mode).  This is the bit-exact checker and fallback for the tcgen05 kernel in
// gemm_tcgen05.cu; both share the epilogues in gemm_epilogue.cuh.
//
// Tiling: CTA 128x128, 8 warps (2 x 4) of 64x32, K step 64 bytes, 4-stage cp.async pipeline,
// XOR-swizzled 16-byte chunks so ldmatrix is conflict-free.  The int8 m16n8k32 and fp16 m16n8k16
// fragments have the same byte layout, so the whole main loop is written in bytes.
#include "common.cuh"
#include "gemm_epilogue.cuh"

namespace b2llm {

namespace {

constexpr int BM = 128, BN = 128, BKB = 64;  // BKB: bytes of K per stage
constexpr int STAGES = 4;
constexpr int THREADS = 256;
constexpr int STAGE_BYTES = (BM + BN) * BKB;

__device__ __forceinline__ uint32_t swz(int row, int chunk) { return row * BKB + ((chunk ^ ((row >> 1) & 3)) << 4); }

template <bool I8, int EPI>
__global__ void __launch_bounds__(THREADS, 2)
    gemm_mma_kernel(const uint8_t* __restrict__ A, const float* __restrict__ a_scale, const uint8_t* __restrict__ W,
                    const float* __restrict__ w_scale, int M, int N, int Kb /* bytes per row */, void* __restrict__ out,
                    int64_t ldc) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;  // warp tile origin: (wm * 64, wn * 32)
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const uint32_t smem_base = smem_u32(smem);

    auto load_stage = [&](int stage, int kb) {
        const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + BM * BKB;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int c = tid + i * THREADS, row = c >> 2, ch = c & 3;
            const int gm = m0 + row;
            const uint8_t* src = A + (int64_t)(gm < M ? gm : 0) * Kb + kb + ch * 16;
            cp_async16(sa + swz(row, ch), src, gm < M ? 16 : 0);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int c = tid + i * THREADS, row = c >> 2, ch = c & 3;
            const int gn = n0 + row;
            const uint8_t* src = W + (int64_t)(gn < N ? gn : 0) * Kb + kb + ch * 16;
            cp_async16(sb + swz(row, ch), src, gn < N ? 16 : 0);
        }
    };

    typename AccT<I8>::type acc[4][4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[i][j][r] = 0;

    const int nk = Kb / BKB;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) load_stage(s, s * BKB);
        cp_async_commit();
    }

    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nxt = kt + STAGES - 1;
            if (nxt < nk) load_stage(nxt % STAGES, nxt * BKB);
            cp_async_commit();
        }
        const uint32_t sa = smem_base + (kt % STAGES) * STAGE_BYTES, sb = sa + BM * BKB;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {  // two 32-byte MMA k-steps per stage
            uint32_t af[4][4], bf[4][2];
            const int j = lane >> 3, r = lane & 7;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = wm * 64 + i * 16 + (j & 1) * 8 + r;
                ldmatrix_x4(af[i][0], af[i][1], af[i][2], af[i][3], sa + swz(row, ks * 2 + (j >> 1)));
            }
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const int row = wn * 32 + p * 16 + (j >> 1) * 8 + r;
                ldmatrix_x4(bf[2 * p][0], bf[2 * p][1], bf[2 * p + 1][0], bf[2 * p + 1][1],
                            sb + swz(row, ks * 2 + (j & 1)));
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jn = 0; jn < 4; ++jn) {
                    if constexpr (I8)
                        mma_s8_16832(acc[i][jn], af[i][0], af[i][1], af[i][2], af[i][3], bf[jn][0], bf[jn][1]);
                    else
                        mma_f16_16816(acc[i][jn], af[i][0], af[i][1], af[i][2], af[i][3], bf[jn][0], bf[jn][1]);
                }
        }
    }
    cp_async_wait<0>();

    // epilogue straight from the accumulator fragments
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int m = m0 + wm * 64 + i * 16 + g + h * 8;
            if (m >= M) continue;
            const float sa_m = I8 ? a_scale[m] : 1.f;
#pragma unroll
            for (int jn = 0; jn < 4; ++jn) {
                const int n = n0 + wn * 32 + jn * 8 + 2 * t;
                if (n >= N) continue;
                float v0, v1;
                if constexpr (I8) {
                    v0 = dequant(acc[i][jn][2 * h], sa_m, w_scale[n]);
                    v1 = dequant(acc[i][jn][2 * h + 1], sa_m, w_scale[n + 1]);
                } else {
                    v0 = acc[i][jn][2 * h];
                    v1 = acc[i][jn][2 * h + 1];
                }
                store_pair<EPI>(out, ldc, N, m, n, v0, v1);
            }
        }
}

template <bool I8, int EPI>
int32_t launch(cudaStream_t s, const void* a, const float* a_scale, const void* w, const float* w_scale, int64_t M,
               int N, int Kb, void* out, int64_t ldc) {
    auto kern = gemm_mma_kernel<I8, EPI>;
    const int smem_bytes = STAGES * STAGE_BYTES;
    B2_ENSURE_DYN_SMEM(kern, smem_bytes);
    dim3 grid((N + BN - 1) / BN, (unsigned)((M + BM - 1) / BM));
    kern<<<grid, THREADS, smem_bytes, s>>>((const uint8_t*)a, a_scale, (const uint8_t*)w, w_scale, (int)M, N, Kb, out,
                                           ldc);
    B2_LAUNCH_CHECK();
    return B2LLM_OK;
}

}  // namespace

int32_t launch_gemm_mma(cudaStream_t s, bool is_i8, const void* a, const float* a_scale, const void* w,
                        const float* w_scale, int64_t M, int N, int K, int epilogue, void* out, int64_t ldc) {
    const int Kb = is_i8 ? K : 2 * K;
    B2_REQUIRE(Kb % BKB == 0, B2LLM_ERR_INVALID_VALUE, "gemm: K must be a multiple of 64 (int8) / 32 (fp16)");
    B2_REQUIRE(N % 2 == 0, B2LLM_ERR_INVALID_VALUE, "gemm: N must be even");
    B2_REQUIRE(M < (1ll << 31), B2LLM_ERR_INVALID_VALUE, "gemm: M too large");
    if (M == 0) return B2LLM_OK;
    if (is_i8) {
        B2_REQUIRE(a_scale && w_scale, B2LLM_ERR_INVALID_VALUE, "gemm_w8a8: scales required");
        switch (epilogue) {
            case EPI_F16: return launch<true, EPI_F16>(s, a, a_scale, w, w_scale, M, N, Kb, out, ldc);
            case EPI_RESIDUAL: return launch<true, EPI_RESIDUAL>(s, a, a_scale, w, w_scale, M, N, Kb, out, ldc);
            case EPI_SWIGLU: return launch<true, EPI_SWIGLU>(s, a, a_scale, w, w_scale, M, N, Kb, out, ldc);
            default: break;
        }
    } else {
        switch (epilogue) {
            case EPI_F16: return launch<false, EPI_F16>(s, a, nullptr, w, nullptr, M, N, Kb, out, ldc);
            case EPI_RESIDUAL: return launch<false, EPI_RESIDUAL>(s, a, nullptr, w, nullptr, M, N, Kb, out, ldc);
            case EPI_SWIGLU: return launch<false, EPI_SWIGLU>(s, a, nullptr, w, nullptr, M, N, Kb, out, ldc);
            case EPI_F32: return launch<false, EPI_F32>(s, a, nullptr, w, nullptr, M, N, Kb, out, ldc);
            default: break;
        }
    }
    set_last_error("gemm: unsupported epilogue");
    return B2LLM_ERR_UNSUPPORTED;
}

}  // namespace b2llm
