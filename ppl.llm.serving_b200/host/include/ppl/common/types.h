// data type / data format tags and the 2-byte float16_t storage type the reference's sources name
// (resource_manager.cc:41,387 sizeof(float16_t); utils.cc:96-99 datatype table).
#ifndef B2LLM_SHIM_PPL_COMMON_TYPES_H_
#define B2LLM_SHIM_PPL_COMMON_TYPES_H_

#include <stdint.h>

namespace ppl { namespace common {

struct float16_t {
    uint16_t bits;
};
static_assert(sizeof(float16_t) == 2, "float16_t must be 2 bytes");

enum {
    DATATYPE_UNKNOWN = 0,
    DATATYPE_UINT8,
    DATATYPE_UINT16,
    DATATYPE_UINT32,
    DATATYPE_UINT64,
    DATATYPE_FLOAT16,
    DATATYPE_FLOAT32,
    DATATYPE_FLOAT64,
    DATATYPE_BFLOAT16,
    DATATYPE_INT4B,
    DATATYPE_INT8,
    DATATYPE_INT16,
    DATATYPE_INT32,
    DATATYPE_INT64,
    DATATYPE_BOOL,
    DATATYPE_MAX,
};
typedef uint32_t datatype_t;

enum {
    DATAFORMAT_UNKNOWN = 0,
    DATAFORMAT_NDARRAY,
    DATAFORMAT_MAX,
};
typedef uint32_t dataformat_t;

inline uint32_t GetSizeOfDataType(datatype_t dt) {
    switch (dt) {
        case DATATYPE_UINT8: case DATATYPE_INT8: case DATATYPE_BOOL: return 1;
        case DATATYPE_UINT16: case DATATYPE_INT16: case DATATYPE_FLOAT16: case DATATYPE_BFLOAT16: return 2;
        case DATATYPE_UINT32: case DATATYPE_INT32: case DATATYPE_FLOAT32: return 4;
        case DATATYPE_UINT64: case DATATYPE_INT64: case DATATYPE_FLOAT64: return 8;
        default: return 0;
    }
}

const char* GetDataTypeStr(datatype_t);

}} // namespace ppl::common

#endif
