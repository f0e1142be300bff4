// Tensor-parallel residual join, fused:  all-reduce of the row-parallel GEMM partials  +  residual add  +  RMSNorm of
// the next block  +  per-token int8 quantisation (or fp16 output)  in ONE kernel over NVLink peer memory.
//
// The reference's tensor parallelism (--tensor-parallel-size; resource_manager.cc:392-422, llm_engine.cc:124) exchanges
// activations at exactly two points per layer: after o_proj and after down_proj (the row-parallel GEMMs).  ppl.nn does
// that with ncclAllReduce followed by separate residual / norm / quant kernels.  Here every rank's communication buffer
// is mapped into all peers (CUDA IPC or plain peer access, exchanged once through the NCCL communicator the reference
// hands over: tp_comm_exchange) and the join is a reduce-scatter by ROWS fused with everything that follows it:
//
//   rank r owns rows {r, r + tp, r + 2 tp, ...}.  For each of its rows it
//     1. reads the fp16 partial of that row from every rank's buffer (P2P loads) and sums them in fp32 in rank order,
//        rounds to fp16                                   -- oracle/llama_ref.py: LlamaOracle._row_parallel
//     2. adds the residual row (fp16(x + o)); rows are owned statically, so only the owner ever needs x -- it is
//        broadcast only at the last join of the step, for the lm head
//     3. RMSNorm with the next block's gain (fp64 variance) and per-token int8 quantisation (W8A8) or fp16 output
//        -- same arithmetic, bit for bit, as rmsnorm_quant_kernel
//     4. writes the int8 row + scale (or the fp16 row) into EVERY rank's buffer (P2P stores): the "all-gather" half
//   Two flag barriers through peer memory bracket the kernel: "my partials are complete" before the loads, "my rows have
//   landed everywhere" before anybody consumes them.  Flags carry the call's epoch, so nothing is ever reset.
//
// One launch replaces ncclAllReduce + rmsnorm_quant (2 x 32 per 7B step), moves 12 KB instead of 16 KB per row and rank
// over NVLink at TP = 8, norms each row once per GROUP instead of once per rank -- and its summation order is the
// oracle's, so tensor-parallel logits agree with the oracle as tightly as single-GPU ones.
#include <unistd.h>

#include <vector>

#include "common.cuh"

namespace b2llm {

namespace {

// 16-byte vectors per thread.  The kernel is latency-bound (NVLink round trips, two block reductions), so what counts
// is ROWS IN FLIGHT: a row needs (TP + 3) x hidden x 2 B of registers whatever the CTA shape, hence few fat threads per
// row -- 128 threads for hidden 4096 at TP = 2 (4-5 CTAs per SM: the 512 rows a rank owns at B = 1024 run as ONE wave;
// with 256-thread CTAs they took two, 25 us per join, round 2 run 8).  TP >= 4 keeps 2 vectors per thread: TP x 4 peer
// loads in flight would not fit the register file, and a rank owns at most 256 rows there anyway.
template <int TP>
struct JoinCfg {
    static constexpr int kIter = TP == 2 ? 4 : 2;
    static constexpr int kMaxThreads = TP == 2 ? 256 : 512;   // hidden <= 8192
};

struct alignas(16) Half8 {
    __half2 v[4];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// release / acquire fence at system scope (lighter than __threadfence_system(), which is sequentially consistent)
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// peer data must come from the owner's memory, never from a stale line of this SM's L1
__device__ __forceinline__ Half8 ld8_peer(const __half* p) {
    Half8 r;
    uint4 u = __ldcg(reinterpret_cast<const uint4*>(p));
    r = *reinterpret_cast<Half8*>(&u);
    return r;
}

// wait until flags[i] has reached `epoch` for every i < n (threads 0..n-1 spin, then the CTA syncs); gives up after
// ~10 s so that a rank that died cannot wedge the other GPUs (the step's results are then garbage and *fault is set)
__device__ __forceinline__ void wait_flags(const uint32_t* flags, int n, uint32_t epoch, unsigned* fault) {
    if ((int)threadIdx.x < n) {
        const uint64_t t0 = globaltimer_ns();
        while ((int32_t)(ld_acquire_sys(flags + threadIdx.x) - epoch) < 0) {
            if (globaltimer_ns() - t0 > 10000000000ull) {
                atomicExch(fault, 1u);
                break;
            }
        }
    }
    __syncthreads();
}

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

// MODE 0: residual join only; 1: + RMSNorm -> int8 row + fp32 scale; 2: + RMSNorm -> fp16 row.
// TP is a template parameter so that the loop over the ranks unrolls: ALL peer loads of a row (TP x kMaxIter 16-byte
// loads per thread) are in flight together.  The first version looped over a run-time rank count and consumed every
// load before issuing the next -- one NVLink round trip per rank and vector: 25 us per join at TP = 4, 49 us at TP = 8
// with hidden 8192 (round 2 run 6: profiles/r2_bench_run6_n{4,8}_tp.json, fused_join_us_per_call).
template <int MODE, int TP>
__global__ void __launch_bounds__(JoinCfg<TP>::kMaxThreads, TP == 2 ? 2 : 1)
    tp_join_kernel(TpPeers c, TpLayout L, const __half* __restrict__ gamma, float eps, int rows, int hidden, int bcast_x,
                   uint32_t epoch) {
    constexpr int kMaxIter = JoinCfg<TP>::kIter;
    const int kThreads = blockDim.x;
    __shared__ float scratch[32];
    __shared__ double dscratch[32];
    __shared__ int s_last;
    pdl_trigger();
    pdl_wait();
    uint8_t* mine = c.base[c.rank];
    uint32_t* my_flags = reinterpret_cast<uint32_t*>(mine + L.flags);  // [0, 8): partials ready; [32, 40): rows delivered
    unsigned* counter = reinterpret_cast<unsigned*>(mine + L.flags) + 64;
    unsigned* fault = counter + 1;
    // phase clocks (ns, %globaltimer) summed over the calls: [0] calls, [1] waiting for the peers' partials, [2] the rows
    // (loads, norm, stores), [3] waiting for the peers' rows to land here; [4] scratch: when barrier 1 opened; CTA 0's first
    // row in detail: [5] peer loads, [6] reductions + quantisation + issuing the peer stores, [7] the system-scope fence
    unsigned long long* stats = reinterpret_cast<unsigned long long*>(mine + L.flags + 512);
    uint64_t t_start = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) t_start = globaltimer_ns();

    // barrier 1: the GEMM before this kernel completed my partials -> tell every rank, wait for every rank
    // (the partials were written by the kernel before this one: complete and visible; the release store orders them)
    if (blockIdx.x == 0 && (int)threadIdx.x < TP)
        st_release_sys(reinterpret_cast<uint32_t*>(c.base[threadIdx.x] + L.flags) + c.rank, epoch);
    wait_flags(my_flags, TP, epoch, fault);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const uint64_t t1 = globaltimer_ns();
        stats[0] += 1;
        stats[1] += t1 - t_start;
        stats[4] = t1;
    }

    const int nvec = hidden >> 3;
    const __half* x_mine = reinterpret_cast<const __half*>(mine + L.x);
    const __half* part[TP];
#pragma unroll
    for (int r = 0; r < TP; ++r) part[r] = reinterpret_cast<const __half*>(c.base[r] + L.partial);
    for (int64_t row = c.rank + (int64_t)TP * blockIdx.x; row < rows; row += (int64_t)TP * gridDim.x) {
        Half8 xn[kMaxIter];
        double ss = 0.0;
        // every peer load of the row first (TP x kMaxIter independent 16-byte loads per thread) ...
        Half8 pv[kMaxIter][TP], xold[kMaxIter];
#pragma unroll
        for (int it = 0; it < kMaxIter; ++it) {
            const int v = threadIdx.x + it * kThreads;
            if (v < nvec) {
                const int64_t off = row * hidden + v * 8;
#pragma unroll
                for (int r = 0; r < TP; ++r) pv[it][r] = ld8_peer(part[r] + off);
                xold[it] = *reinterpret_cast<const Half8*>(x_mine + off);
            }
        }
        // ... then the arithmetic
        const bool clocked = blockIdx.x == 0 && threadIdx.x == 0 && row == c.rank;  // CTA 0's first row carries the fine clocks
        uint64_t tc0 = 0, tc1 = 0, tc2 = 0;
        if (clocked) tc0 = globaltimer_ns();
#pragma unroll
        for (int it = 0; it < kMaxIter; ++it) {
            const int v = threadIdx.x + it * kThreads;
            if (v >= nvec) break;
            const int64_t off = row * hidden + v * 8;
            float acc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
            for (int r = 0; r < TP; ++r) {  // fp32 sum in rank order
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = __half22float2(pv[it][r].v[i]);
                    acc[2 * i] = __fadd_rn(acc[2 * i], f.x);
                    acc[2 * i + 1] = __fadd_rn(acc[2 * i + 1], f.y);
                }
            }
            const Half8 xo = xold[it];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 o = __half22float2(__floats2half2_rn(acc[2 * i], acc[2 * i + 1]));  // the all-reduced value is fp16
                const float2 fx = __half22float2(xo.v[i]);
                xn[it].v[i] = __floats2half2_rn(__fadd_rn(fx.x, o.x), __fadd_rn(fx.y, o.y));
                const float2 f = __half22float2(xn[it].v[i]);
                ss += (double)f.x * (double)f.x + (double)f.y * (double)f.y;
            }
            if (bcast_x) {
#pragma unroll
                for (int r = 0; r < TP; ++r) *reinterpret_cast<Half8*>(reinterpret_cast<__half*>(c.base[r] + L.x) + off) = xn[it];
            } else {
                *reinterpret_cast<Half8*>(reinterpret_cast<__half*>(mine + L.x) + off) = xn[it];
            }
        }
        if constexpr (MODE == 0) continue;
        if (clocked) tc1 = globaltimer_ns();   // the peer loads have arrived (the sums above consumed them)
        ss = block_sum_f64(ss, dscratch);
        const float var = (float)(ss / (double)hidden);
        const float inv = __fdiv_rn(1.0f, sqrtf(__fadd_rn(var, eps)));
        float y[kMaxIter][8];
        float amax = 0.f;
#pragma unroll
        for (int it = 0; it < kMaxIter; ++it) {
            const int v = threadIdx.x + it * kThreads;
            if (v >= nvec) break;
            const Half8 g = *reinterpret_cast<const Half8*>(gamma + v * 8);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 fa = __half22float2(xn[it].v[i]), fg = __half22float2(g.v[i]);
                y[it][2 * i] = __fmul_rn(__fmul_rn(fa.x, inv), fg.x);
                y[it][2 * i + 1] = __fmul_rn(__fmul_rn(fa.y, inv), fg.y);
                amax = fmaxf(amax, fmaxf(fabsf(y[it][2 * i]), fabsf(y[it][2 * i + 1])));
            }
        }
        if constexpr (MODE == 2) {
#pragma unroll
            for (int it = 0; it < kMaxIter; ++it) {
                const int v = threadIdx.x + it * kThreads;
                if (v >= nvec) break;
                Half8 o;
#pragma unroll
                for (int i = 0; i < 4; ++i) o.v[i] = __floats2half2_rn(y[it][2 * i], y[it][2 * i + 1]);
#pragma unroll
                for (int r = 0; r < TP; ++r)
                    *reinterpret_cast<Half8*>(reinterpret_cast<__half*>(c.base[r] + L.y) + row * hidden + v * 8) = o;
            }
        } else {
            amax = block_max(amax, scratch);
            const float inv_scale = amax > 0.f ? __fdiv_rn(127.0f, amax) : 0.f;
            if ((int)threadIdx.x < TP) reinterpret_cast<float*>(c.base[threadIdx.x] + L.qscale)[row] = __fdiv_rn(amax, 127.0f);
#pragma unroll
            for (int it = 0; it < kMaxIter; ++it) {
                const int v = threadIdx.x + it * kThreads;
                if (v >= nvec) break;
                uint32_t w[2] = {0, 0};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int q = max(-127, min(127, __float2int_rn(__fmul_rn(y[it][i], inv_scale))));
                    w[i >> 2] |= (uint32_t)(q & 0xff) << (8 * (i & 3));
                }
#pragma unroll
                for (int r = 0; r < TP; ++r)
                    *reinterpret_cast<uint2*>(c.base[r] + L.q + row * hidden + v * 8) = make_uint2(w[0], w[1]);
            }
        }
        if (clocked) {
            tc2 = globaltimer_ns();            // norm, quantisation and the peer stores are issued
            stats[5] += tc1 - tc0;
            stats[6] += tc2 - tc1;
        }
    }

    // barrier 2: once every CTA of this rank has pushed its rows out, tell every rank; the last CTA stays until every
    // rank's rows have landed here, so the kernel's end means "the joined activations are complete on this GPU"
    // one system-scope fence per CTA: the CTA barrier orders every thread's peer stores before thread 0's fence, which is
    // cumulative (the grid-sync idiom, at system scope) -- 255 fewer fences per CTA than fencing in every thread
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint64_t tf0 = blockIdx.x == 0 ? globaltimer_ns() : 0;
        fence_acq_rel_sys();
        if (blockIdx.x == 0) stats[7] += globaltimer_ns() - tf0;   // how long the system-scope fence holds CTA 0
        s_last = atomicAdd(counter, 1u) == gridDim.x - 1 ? 1 : 0;
        if (s_last) fence_acq_rel_sys();  // acquire side of the other CTAs' "fence; atomicAdd"
    }
    __syncthreads();
    if (!s_last) return;
    uint64_t t_done = 0;
    if (threadIdx.x == 0) t_done = globaltimer_ns();
    // every CTA fenced its peer stores at system scope before its atomicAdd, and this CTA observed all of them through the
    // counter: a release store is all the flag needs
    if ((int)threadIdx.x < TP)
        st_release_sys(reinterpret_cast<uint32_t*>(c.base[threadIdx.x] + L.flags) + 32 + c.rank, epoch);
    wait_flags(my_flags + 32, TP, epoch, fault);
    if (threadIdx.x == 0) {
        *counter = 0;
        const uint64_t t_end = globaltimer_ns();
        const uint64_t t1 = *reinterpret_cast<volatile unsigned long long*>(stats + 4);
        stats[2] += t_done > t1 ? t_done - t1 : 0;
        stats[3] += t_end - t_done;
    }
}

}  // namespace

TpLayout tp_layout(int64_t max_tokens, int hidden, int act_cols) {
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    TpLayout L{};
    size_t o = 1024;  // flags (2 x 32 words), CTA counter, fault word
    L.flags = 0;
    L.partial = o; o = up(o + (size_t)max_tokens * hidden * 2);
    L.x = o;       o = up(o + (size_t)max_tokens * hidden * 2);
    L.q = o;       o = up(o + (size_t)max_tokens * act_cols);
    L.qscale = o;  o = up(o + (size_t)max_tokens * 4);
    L.y = o;       o = up(o + (size_t)max_tokens * hidden * 2);
    L.total = o;
    return L;
}

int32_t launch_tp_join(cudaStream_t s, const TpPeers& peers, const TpLayout& L, int mode, bool bcast_x, const __half* gamma,
                       float eps, int64_t rows, int hidden, uint32_t epoch) {
    B2_REQUIRE(hidden % 8 == 0 && hidden <= 8192, B2LLM_ERR_UNSUPPORTED, "tp join: hidden must be a multiple of 8, <= 8192");
    B2_REQUIRE(peers.tp == 2 || peers.tp == 4 || peers.tp == 8, B2LLM_ERR_UNSUPPORTED, "tp join: 2, 4 or 8 ranks");
    B2_REQUIRE(mode >= 0 && mode <= 2, B2LLM_ERR_INVALID_VALUE, "tp join: bad mode");
    const int64_t owned = (rows - peers.rank + peers.tp - 1) / peers.tp;
    // every rank launches, also one that owns no row of a small step: the flag barriers are collective
    const dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(owned, 148 * 4)));
    const int iters = peers.tp == 2 ? JoinCfg<2>::kIter : JoinCfg<8>::kIter;
    const dim3 block((unsigned)((((hidden / 8) + iters - 1) / iters + 31) / 32 * 32));
    const int bx = bcast_x ? 1 : 0;
#define B2_JOIN(MM, TT) launch_kernel(tp_join_kernel<MM, TT>, grid, block, 0, s, peers, L, gamma, eps, (int)rows, hidden, bx, epoch)
#define B2_JOIN_TP(MM)                                   \
    do {                                                 \
        if (peers.tp == 2) B2_JOIN(MM, 2);               \
        else if (peers.tp == 4) B2_JOIN(MM, 4);          \
        else B2_JOIN(MM, 8);                             \
    } while (0)
    if (mode == 0) B2_JOIN_TP(0);
    else if (mode == 1) B2_JOIN_TP(1);
    else B2_JOIN_TP(2);
#undef B2_JOIN_TP
#undef B2_JOIN
    B2_LAUNCH_CHECK();
    return B2LLM_OK;
}

// ---- peer mapping.  What each rank publishes about its buffer; all-gathered through the engine's NCCL communicator.
struct TpXchg {
    cudaIpcMemHandle_t handle;  // 64 bytes
    uint64_t ptr;
    int32_t pid, dev;
};

int32_t tp_comm_exchange(cudaStream_t s, void* nccl_comm, int (*allgather)(const void*, void*, size_t, int, void*, cudaStream_t),
                         int rank, int tp, void* local, TpPeers* out, std::vector<void*>* ipc_opened) {
    B2_REQUIRE(tp <= kTpMaxRanks, B2LLM_ERR_UNSUPPORTED, "tp join: at most 8 ranks");
    TpXchg mine{};
    B2_CHECK_CUDA(cudaIpcGetMemHandle(&mine.handle, local));
    mine.ptr = (uint64_t)(uintptr_t)local;
    mine.pid = (int32_t)getpid();
    B2_CHECK_CUDA(cudaGetDevice(&mine.dev));
    void* dbuf = nullptr;
    B2_CHECK_CUDA(cudaMalloc(&dbuf, sizeof(TpXchg) * (tp + 1)));
    std::vector<TpXchg> all(tp);
    cudaError_t ce = cudaMemcpyAsync(dbuf, &mine, sizeof(TpXchg), cudaMemcpyHostToDevice, s);
    int nr = 0;
    if (ce == cudaSuccess) nr = allgather(dbuf, (char*)dbuf + sizeof(TpXchg), sizeof(TpXchg), 0 /* ncclInt8 */, nccl_comm, s);
    if (ce == cudaSuccess && nr == 0)
        ce = cudaMemcpyAsync(all.data(), (char*)dbuf + sizeof(TpXchg), sizeof(TpXchg) * tp, cudaMemcpyDeviceToHost, s);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
    cudaFree(dbuf);
    B2_REQUIRE(nr == 0, B2LLM_ERR_DEVICE, "tp join: ncclAllGather of the peer handles failed with code " + std::to_string(nr));
    B2_CHECK_CUDA(ce);
    TpPeers p{};
    p.tp = tp;
    p.rank = rank;
    for (int r = 0; r < tp; ++r) {
        if (r == rank) {
            p.base[r] = (uint8_t*)local;
        } else if (all[r].pid == mine.pid) {  // one process, one thread per GPU (the reference's host): plain peer access
            int can = 0;
            B2_CHECK_CUDA(cudaDeviceCanAccessPeer(&can, mine.dev, all[r].dev));
            B2_REQUIRE(can, B2LLM_ERR_UNSUPPORTED, "tp join: no peer access between the GPUs of the group");
            const cudaError_t e = cudaDeviceEnablePeerAccess(all[r].dev, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) B2_CHECK_CUDA(e);
            cudaGetLastError();
            p.base[r] = (uint8_t*)(uintptr_t)all[r].ptr;
        } else {  // one process per GPU (torchrun): CUDA IPC
            void* mapped = nullptr;
            B2_CHECK_CUDA(cudaIpcOpenMemHandle(&mapped, all[r].handle, cudaIpcMemLazyEnablePeerAccess));
            ipc_opened->push_back(mapped);
            p.base[r] = (uint8_t*)mapped;
        }
    }
    *out = p;
    return B2LLM_OK;
}

}  // namespace b2llm
