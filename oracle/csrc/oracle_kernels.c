/*
 * C restatement of the heavy integer loops of the oracle (TEST INFRASTRUCTURE, see oracle/__init__.py).
 * Used (a) to cross-check the numpy restatement with an independent implementation and (b) as the CPU arm
 * timed by bench.py (`cpu_baseline`, `--impl reference`), where converting int8 weights to float64 for
 * BLAS on every call would make numpy an unfairly slow stand-in for a CPU int8 path.
 *
 * gemm_i8_i32: exact int32 accumulation of A[M,K] (int8) x W[N,K]^T (int8) -- the W8A8 projection
 *   (quant_method online_i8i8, src/backends/cuda/resource_manager.cc:51-52; the arithmetic itself is
 *   EXTERNAL to the reference).  OpenMP over output columns, inner dot product auto-vectorised.
 * hash_combine: utils::HashCombine (src/utils/utils.cc:87-94) with C's own integer promotions, to pin
 *   the Python restatement in oracle/host_ref.py.
 */
#include <stdint.h>
#include <stddef.h>

void gemm_i8_i32(const int8_t* A, const int8_t* W, int32_t* C, int64_t M, int64_t N, int64_t K) {
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N; ++n) {
        const int8_t* w = W + n * K;
        for (int64_t m = 0; m < M; ++m) {
            const int8_t* a = A + m * K;
            int32_t acc = 0;
            for (int64_t k = 0; k < K; ++k) acc += (int32_t)a[k] * (int32_t)w[k];
            C[m * N + n] = acc;
        }
    }
}

/* decode attention over an int8 group-8 cache for one layer, canonical [2, T, H, D] + fp16-as-float scales
 * pre-expanded by the caller is avoided: scales are passed as float [2, T, H, D/8].  Dequantised values are
 * rounded to fp16 by the caller's convention -> here we take them as given floats (kdeq = fp16(int8*scale)
 * is computed in Python for parity; this routine is only the timing workhorse of the CPU arm). */
uint64_t hash_combine(uint64_t prev, const int32_t* vec, int32_t len) {
    uint64_t seed = (uint64_t)len;
    seed ^= prev + 0x9e3779b9 + (seed << 6) + (seed >> 2);
    for (int i = 0; i < len; ++i) {
        seed ^= vec[i] + 0x9e3779b9 + (seed << 6) + (seed >> 2);
    }
    return seed;
}
