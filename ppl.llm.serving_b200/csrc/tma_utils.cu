#include "tma_utils.cuh"

#include <stdlib.h>

#include <mutex>

namespace b2llm {

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_num_sms = 0;
std::once_flag g_once;

void init_once() {
    std::call_once(g_once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            g_encode = (EncodeTiledFn)fn;
        else
            cudaGetLastError();
        int dev = 0;
        cudaDeviceProp prop;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess && prop.major == 10)
            g_num_sms = prop.multiProcessorCount;
        else
            cudaGetLastError();
    });
}
}  // namespace

bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("B2LLM_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}

bool tma_available() {
    init_once();
    return g_encode != nullptr && g_num_sms > 0;
}

int device_num_sms() {
    init_once();
    return g_num_sms;
}

bool tma_encode_bytes(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides,
                      const uint32_t* box, bool swizzle128) {
    if (!tma_available() || rank < 2 || rank > 5) return false;
    cuuint64_t d[5], st[4];
    cuuint32_t b[5], es[5];
    for (int i = 0; i < rank; ++i) {
        d[i] = dims[i];
        b[i] = box[i];
        es[i] = 1;
        if (i + 1 < rank) st[i] = strides[i];
    }
    return g_encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, (cuuint32_t)rank, const_cast<void*>(ptr), d, st, b, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace b2llm
