// ppl::nn::DeviceContext (EXTERNAL): where a tensor's buffer lives; the reference identifies the CUDA one by
// its 8-byte type tag and asks it for its stream (resource_manager.cc:182-211).
#ifndef B2LLM_SHIM_PPL_NN_COMMON_DEVICE_CONTEXT_H_
#define B2LLM_SHIM_PPL_NN_COMMON_DEVICE_CONTEXT_H_

#include "ppl/common/retcode.h"

#include <stdint.h>
#include <string.h>

namespace ppl { namespace nn {

class DeviceContext {
public:
    struct Type final {
        char str[8];
        Type() {
            memset(str, 0, sizeof(str));
        }
        bool operator==(const Type& rhs) const {
            return memcmp(str, rhs.str, sizeof(str)) == 0;
        }
        bool operator!=(const Type& rhs) const {
            return !(*this == rhs);
        }
    };

    virtual ~DeviceContext() {}
    virtual const Type& GetType() const = 0;
    virtual ppl::common::RetCode Configure(uint32_t option, ...) = 0;
};

}} // namespace ppl::nn

#endif
