// ppl::common::RetCode -- the error convention of the reference: an int enum, RC_SUCCESS == 0,
// no exceptions.  ppl.common itself is EXTERNAL to /root/reference (hpcc @ master, cmake/deps.cmake:32-58);
// this header provides the names the reference's sources use, with values shared with include/b2llm.h.
#ifndef B2LLM_SHIM_PPL_COMMON_RETCODE_H_
#define B2LLM_SHIM_PPL_COMMON_RETCODE_H_

#include <stdint.h>

namespace ppl { namespace common {

enum {
    RC_SUCCESS = 0,
    RC_OTHER_ERROR = 1,          // == B2LLM_ERR_OTHER
    RC_INVALID_VALUE = 2,        // == B2LLM_ERR_INVALID_VALUE
    RC_OUT_OF_MEMORY = 3,        // == B2LLM_ERR_OUT_OF_MEMORY
    RC_UNSUPPORTED = 4,          // == B2LLM_ERR_UNSUPPORTED
    RC_DEVICE_RUNTIME_ERROR = 5, // == B2LLM_ERR_DEVICE
    RC_DEVICE_MEMORY_ERROR = 6,  // == B2LLM_ERR_DEVICE_MEMORY
    RC_NOT_FOUND = 7,
    RC_EXISTS = 8,
    RC_OUT_OF_RANGE = 9,
    RC_PERMISSION_DENIED = 10,
    RC_SIGN_IN = 11,
};
typedef uint32_t RetCode;

const char* GetRetCodeStr(RetCode);

}} // namespace ppl::common

#endif
