// ppl::nn::Tensor (EXTERNAL): the subset the reference's sources call
// (llm_engine.h:124-147, llm_engine.cc:31-110,206-222, resource_manager.cc:297-304, utils.cc:124-163).
#ifndef B2LLM_SHIM_PPL_NN_RUNTIME_TENSOR_H_
#define B2LLM_SHIM_PPL_NN_RUNTIME_TENSOR_H_

#include "ppl/nn/common/device_context.h"
#include "ppl/nn/common/tensor_shape.h"

namespace ppl { namespace nn {

class Tensor {
public:
    virtual ~Tensor() {}
    virtual const char* GetName() const = 0;
    virtual TensorShape* GetShape() const = 0;

    virtual ppl::common::RetCode SetDeviceContext(DeviceContext*) = 0;
    virtual DeviceContext* GetDeviceContext() const = 0;

    /** points the tensor at caller-owned memory (the KV cache, llm_engine.h:142-146) */
    virtual void SetBufferPtr(void*) = 0;
    virtual void* GetBufferPtr() const = 0;
    virtual void FreeBuffer() = 0;
    virtual ppl::common::RetCode ReallocBuffer() = 0;

    /** copies GetShape()->CalcBytesIncludingPadding() bytes; async on the device context's stream */
    virtual ppl::common::RetCode CopyFromHostAsync(const void* src) = 0;
    virtual ppl::common::RetCode CopyFromHost(const void* src) = 0;
    virtual ppl::common::RetCode CopyToHostAsync(void* dst) const = 0;
    virtual ppl::common::RetCode CopyToHost(void* dst) const = 0;
    virtual ppl::common::RetCode ConvertToHost(void* dst, const TensorShape& dst_desc) const = 0;
};

}} // namespace ppl::nn

#endif
