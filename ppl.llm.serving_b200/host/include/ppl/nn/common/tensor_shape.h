// ppl::nn::TensorShape (EXTERNAL): the subset the reference's sources call
// (llm_engine.cc:38-101,118-169 Reshape({...}); :220 GetDim(1); utils.cc:107-135).
#ifndef B2LLM_SHIM_PPL_NN_COMMON_TENSOR_SHAPE_H_
#define B2LLM_SHIM_PPL_NN_COMMON_TENSOR_SHAPE_H_

#include "ppl/common/retcode.h"
#include "ppl/common/types.h"

#include <stdint.h>
#include <string>
#include <vector>

namespace ppl { namespace nn {

class TensorShape final {
public:
    void Reshape(const std::vector<int64_t>& dims) {
        dims_ = dims;
    }
    void Reshape(const int64_t* dims, uint32_t count) {
        dims_.assign(dims, dims + count);
    }
    void ReshapeAsScalar() {
        dims_.clear();
    }
    uint32_t GetDimCount() const {
        return (uint32_t)dims_.size();
    }
    uint32_t GetRealDimCount() const {
        return (uint32_t)dims_.size();
    }
    int64_t GetDim(uint32_t i) const {
        return dims_[i];
    }
    const int64_t* GetDims() const {
        return dims_.data();
    }
    bool IsScalar() const {
        return dims_.empty();
    }
    uint64_t CalcElementsIncludingPadding() const {
        uint64_t n = 1;
        for (auto d : dims_) {
            n *= (uint64_t)d;
        }
        return n;
    }
    uint64_t CalcElementsExcludingPadding() const {
        return CalcElementsIncludingPadding();
    }
    uint64_t CalcBytesIncludingPadding() const {
        return CalcElementsIncludingPadding() * ppl::common::GetSizeOfDataType(data_type_);
    }
    uint64_t CalcBytesExcludingPadding() const {
        return CalcBytesIncludingPadding();
    }
    void SetDataType(ppl::common::datatype_t dt) {
        data_type_ = dt;
    }
    ppl::common::datatype_t GetDataType() const {
        return data_type_;
    }
    void SetDataFormat(ppl::common::dataformat_t df) {
        data_format_ = df;
    }
    ppl::common::dataformat_t GetDataFormat() const {
        return data_format_;
    }

private:
    std::vector<int64_t> dims_;
    ppl::common::datatype_t data_type_ = ppl::common::DATATYPE_UNKNOWN;
    ppl::common::dataformat_t data_format_ = ppl::common::DATAFORMAT_NDARRAY;
};

template <typename T>
inline std::string ToString(const T& v) {
    return std::to_string(v);
}

}} // namespace ppl::nn

#endif
