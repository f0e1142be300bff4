// An ONNX ModelProto as far as a ppl.pmx LLaMA export needs it: the node list with attributes, and the initializers
// with their payload bytes (raw_data inside the mmap'ed .onnx, typed *_data fields, or ONNX "external data" files in
// the same directory).  Reference: the model the tools load is <model-dir>/model_slice_<rank>/model.onnx
// (src/backends/cuda/resource_manager.cc:280-290), schema src/onnx/onnx.proto.
#ifndef B2_ONNX_MODEL_H_
#define B2_ONNX_MODEL_H_

#include <stdint.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

namespace b2onnx {

// TensorProto.DataType (onnx.proto:480-507)
enum { DT_FLOAT = 1, DT_UINT8 = 2, DT_INT8 = 3, DT_INT32 = 6, DT_INT64 = 7, DT_FLOAT16 = 10, DT_DOUBLE = 11, DT_BFLOAT16 = 16 };
// AttributeProto.AttributeType (onnx.proto:118-134)
enum { AT_FLOAT = 1, AT_INT = 2, AT_STRING = 3, AT_TENSOR = 4, AT_FLOATS = 6, AT_INTS = 7, AT_STRINGS = 8 };

struct Tensor {
    std::string name;
    int32_t data_type = 0;
    std::vector<int64_t> dims;
    const uint8_t* data = nullptr; // payload in the element type's little-endian layout (fp16: 2 bytes / element)
    uint64_t bytes = 0;
    bool external = false;
    std::string location; // external data file, relative to the model's directory
    std::vector<uint8_t> owned; // payload rebuilt from float_data / int32_data / int64_data

    uint64_t NumElements() const {
        uint64_t n = 1;
        for (int64_t d : dims) n *= (uint64_t)d;
        return n;
    }
};

struct Attribute {
    std::string name;
    int32_t type = 0;
    int64_t i = 0;
    float f = 0.f;
    std::string s;
    std::vector<int64_t> ints;
    std::vector<float> floats;
};

struct Node {
    std::string name, op_type, domain;
    std::vector<std::string> inputs, outputs;
    std::vector<Attribute> attrs;

    const Attribute* Find(const char* attr_name) const;
    int64_t Int(const char* attr_name, int64_t dflt) const;
    float Float(const char* attr_name, float dflt) const;
    std::string Str(const char* attr_name, const std::string& dflt) const;
};

class Model {
public:
    Model() = default;
    ~Model();
    Model(const Model&) = delete;
    Model& operator=(const Model&) = delete;

    // false + *err on a missing / truncated / malformed file or unreadable external data
    bool Load(const std::string& path, std::string* err);

    const Tensor* FindInitializer(const std::string& name) const;

    std::string path, dir, producer_name, producer_version, graph_name;
    int64_t ir_version = 0;
    std::map<std::string, int64_t> opsets; // domain -> version
    std::vector<Node> nodes;
    std::vector<Tensor> initializers;
    std::vector<std::string> graph_inputs, graph_outputs;

private:
    struct Mapping {
        void* addr;
        uint64_t len;
    };
    bool MapFile(const std::string& file, const uint8_t** p, uint64_t* n, std::string* err);
    std::map<std::string, Mapping> maps_;
    std::map<std::string, size_t> by_name_;
};

} // namespace b2onnx
#endif
