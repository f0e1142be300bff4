// ONNX ModelProto reader over the wire format (no libprotobuf).  Field numbers: src/onnx/onnx.proto of the reference
// (ModelProto :340-412, GraphProto :439-470, NodeProto :196-213, AttributeProto :114-174, TensorProto :479-602,
// StringStringEntryProto :417-420, OperatorSetIdProto :744-755).
#include "onnx_model.h"

#include "onnx_wire.h"

#include <fcntl.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>

namespace b2onnx {

namespace {

std::string Str(const View& v) {
    return std::string((const char*)v.p, (size_t)v.n);
}

// repeated scalar that may arrive packed (wire 2) or one element per tag (wire 0 / 5)
bool ReadVarints(WireReader& r, uint32_t wire, std::vector<int64_t>* out) {
    if (wire == 0) {
        uint64_t v;
        if (!r.Varint(&v)) return false;
        out->push_back((int64_t)v);
        return true;
    }
    if (wire != 2) return false;
    View pk;
    if (!r.Bytes(&pk)) return false;
    WireReader pr(pk);
    while (!pr.AtEnd()) {
        uint64_t v;
        if (!pr.Varint(&v)) return false;
        out->push_back((int64_t)v);
    }
    return true;
}

bool ReadFloats(WireReader& r, uint32_t wire, std::vector<float>* out) {
    if (wire == 5) {
        uint32_t u;
        if (!r.Fixed32(&u)) return false;
        float f;
        memcpy(&f, &u, 4);
        out->push_back(f);
        return true;
    }
    if (wire != 2) return false;
    View pk;
    if (!r.Bytes(&pk)) return false;
    if (pk.n % 4) return false;
    const size_t old = out->size();
    out->resize(old + pk.n / 4);
    memcpy(out->data() + old, pk.p, pk.n);
    return true;
}

struct RawTensor { // TensorProto before its payload is resolved
    Tensor t;
    std::vector<float> float_data;
    std::vector<int64_t> int32_data, int64_data;
    int64_t ext_offset = 0, ext_length = -1;
    bool has_raw = false;
};

bool ParseTensor(const View& msg, RawTensor* out) {
    WireReader r(msg);
    uint32_t f, w;
    while (r.Next(&f, &w)) {
        View v;
        uint64_t u;
        switch (f) {
            case 1: // dims
                if (!ReadVarints(r, w, &out->t.dims)) return false;
                break;
            case 2: // data_type
                if (w != 0 || !r.Varint(&u)) return false;
                out->t.data_type = (int32_t)u;
                break;
            case 4: // float_data
                if (!ReadFloats(r, w, &out->float_data)) return false;
                break;
            case 5: // int32_data (also carries fp16 / bf16 bit patterns, int8, uint8 ...)
                if (!ReadVarints(r, w, &out->int32_data)) return false;
                break;
            case 7: // int64_data
                if (!ReadVarints(r, w, &out->int64_data)) return false;
                break;
            case 8: // name
                if (w != 2 || !r.Bytes(&v)) return false;
                out->t.name = Str(v);
                break;
            case 9: // raw_data
                if (w != 2 || !r.Bytes(&v)) return false;
                out->t.data = v.p;
                out->t.bytes = v.n;
                out->has_raw = true;
                break;
            case 13: { // external_data: StringStringEntryProto
                if (w != 2 || !r.Bytes(&v)) return false;
                WireReader er(v);
                uint32_t ef, ew;
                std::string key, value;
                while (er.Next(&ef, &ew)) {
                    View ev;
                    if (ew == 2 && (ef == 1 || ef == 2)) {
                        if (!er.Bytes(&ev)) return false;
                        (ef == 1 ? key : value) = Str(ev);
                    } else if (!er.Skip(ew)) {
                        return false;
                    }
                }
                if (!er.ok()) return false;
                if (key == "location") out->t.location = value;
                else if (key == "offset") out->ext_offset = atoll(value.c_str());
                else if (key == "length") out->ext_length = atoll(value.c_str());
                break;
            }
            case 14: // data_location
                if (w != 0 || !r.Varint(&u)) return false;
                out->t.external = (u == 1);
                break;
            default:
                if (!r.Skip(w)) return false;
        }
    }
    return r.ok();
}

bool ParseAttribute(const View& msg, Attribute* a) {
    WireReader r(msg);
    uint32_t f, w;
    while (r.Next(&f, &w)) {
        View v;
        uint64_t u;
        uint32_t u32;
        switch (f) {
            case 1:
                if (w != 2 || !r.Bytes(&v)) return false;
                a->name = Str(v);
                break;
            case 2:
                if (w != 5 || !r.Fixed32(&u32)) return false;
                memcpy(&a->f, &u32, 4);
                break;
            case 3:
                if (w != 0 || !r.Varint(&u)) return false;
                a->i = (int64_t)u;
                break;
            case 4:
                if (w != 2 || !r.Bytes(&v)) return false;
                a->s = Str(v);
                break;
            case 7:
                if (!ReadFloats(r, w, &a->floats)) return false;
                break;
            case 8:
                if (!ReadVarints(r, w, &a->ints)) return false;
                break;
            case 20:
                if (w != 0 || !r.Varint(&u)) return false;
                a->type = (int32_t)u;
                break;
            default: // tensors, graphs, strings, doc strings: not needed for a pmx LLaMA graph
                if (!r.Skip(w)) return false;
        }
    }
    return r.ok();
}

bool ParseNode(const View& msg, Node* n) {
    WireReader r(msg);
    uint32_t f, w;
    while (r.Next(&f, &w)) {
        View v;
        if (w == 2 && (f == 1 || f == 2 || f == 3 || f == 4 || f == 5 || f == 7)) {
            if (!r.Bytes(&v)) return false;
            switch (f) {
                case 1: n->inputs.push_back(Str(v)); break;
                case 2: n->outputs.push_back(Str(v)); break;
                case 3: n->name = Str(v); break;
                case 4: n->op_type = Str(v); break;
                case 7: n->domain = Str(v); break;
                case 5: {
                    Attribute a;
                    if (!ParseAttribute(v, &a)) return false;
                    n->attrs.push_back(std::move(a));
                    break;
                }
            }
        } else if (!r.Skip(w)) {
            return false;
        }
    }
    return r.ok();
}

bool ParseValueInfoName(const View& msg, std::string* name) {
    WireReader r(msg);
    uint32_t f, w;
    while (r.Next(&f, &w)) {
        View v;
        if (f == 1 && w == 2) {
            if (!r.Bytes(&v)) return false;
            *name = Str(v);
        } else if (!r.Skip(w)) {
            return false;
        }
    }
    return r.ok();
}

uint32_t ElementSize(int32_t dt) {
    switch (dt) {
        case DT_FLOAT: case DT_INT32: return 4;
        case DT_UINT8: case DT_INT8: return 1;
        case DT_INT64: case DT_DOUBLE: return 8;
        case DT_FLOAT16: case DT_BFLOAT16: return 2;
        default: return 0;
    }
}

} // namespace

const Attribute* Node::Find(const char* attr_name) const {
    for (const auto& a : attrs) {
        if (a.name == attr_name) return &a;
    }
    return nullptr;
}
int64_t Node::Int(const char* attr_name, int64_t dflt) const {
    const Attribute* a = Find(attr_name);
    return a ? a->i : dflt;
}
float Node::Float(const char* attr_name, float dflt) const {
    const Attribute* a = Find(attr_name);
    return a ? a->f : dflt;
}
std::string Node::Str(const char* attr_name, const std::string& dflt) const {
    const Attribute* a = Find(attr_name);
    return a ? a->s : dflt;
}

Model::~Model() {
    for (auto& kv : maps_) munmap(kv.second.addr, kv.second.len);
}

bool Model::MapFile(const std::string& file, const uint8_t** p, uint64_t* n, std::string* err) {
    auto it = maps_.find(file);
    if (it == maps_.end()) {
        const int fd = open(file.c_str(), O_RDONLY);
        if (fd < 0) {
            *err = "cannot open [" + file + "]";
            return false;
        }
        struct stat st;
        if (fstat(fd, &st) != 0 || st.st_size == 0) {
            close(fd);
            *err = "[" + file + "] is empty or unreadable";
            return false;
        }
        void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        close(fd);
        if (m == MAP_FAILED) {
            *err = "mmap of [" + file + "] failed";
            return false;
        }
        it = maps_.emplace(file, Mapping{m, (uint64_t)st.st_size}).first;
    }
    *p = (const uint8_t*)it->second.addr;
    *n = it->second.len;
    return true;
}

const Tensor* Model::FindInitializer(const std::string& name) const {
    auto it = by_name_.find(name);
    return it == by_name_.end() ? nullptr : &initializers[it->second];
}

bool Model::Load(const std::string& model_path, std::string* err) {
    path = model_path;
    const auto slash = path.rfind('/');
    dir = slash == std::string::npos ? "." : path.substr(0, slash);
    const uint8_t* base = nullptr;
    uint64_t size = 0;
    if (!MapFile(path, &base, &size, err)) return false;

    View graph;
    bool have_graph = false;
    {
        WireReader r(base, size);
        uint32_t f, w;
        while (r.Next(&f, &w)) {
            View v;
            uint64_t u;
            if (f == 1 && w == 0) {
                if (!r.Varint(&u)) break;
                ir_version = (int64_t)u;
            } else if ((f == 2 || f == 3) && w == 2) {
                if (!r.Bytes(&v)) break;
                (f == 2 ? producer_name : producer_version) = Str(v);
            } else if (f == 7 && w == 2) {
                if (!r.Bytes(&graph)) break;
                have_graph = true;
            } else if (f == 8 && w == 2) {
                if (!r.Bytes(&v)) break;
                WireReader orr(v);
                uint32_t of, ow;
                std::string domain;
                int64_t version = 0;
                while (orr.Next(&of, &ow)) {
                    View ov;
                    if (of == 1 && ow == 2) {
                        if (!orr.Bytes(&ov)) break;
                        domain = Str(ov);
                    } else if (of == 2 && ow == 0) {
                        if (!orr.Varint(&u)) break;
                        version = (int64_t)u;
                    } else if (!orr.Skip(ow)) {
                        break;
                    }
                }
                opsets[domain] = version;
            } else if (!r.Skip(w)) {
                break;
            }
        }
        if (!r.ok() || !have_graph) {
            *err = "[" + path + "] is not an ONNX ModelProto (" + (r.ok() ? "no graph" : "malformed protobuf") + ")";
            return false;
        }
    }

    std::vector<RawTensor> raw;
    {
        WireReader r(graph);
        uint32_t f, w;
        while (r.Next(&f, &w)) {
            View v;
            if (w == 2 && (f == 1 || f == 2 || f == 5 || f == 11 || f == 12)) {
                if (!r.Bytes(&v)) break;
                bool ok = true;
                if (f == 1) {
                    nodes.emplace_back();
                    ok = ParseNode(v, &nodes.back());
                } else if (f == 2) {
                    graph_name = Str(v);
                } else if (f == 5) {
                    raw.emplace_back();
                    ok = ParseTensor(v, &raw.back());
                } else {
                    std::string name;
                    ok = ParseValueInfoName(v, &name);
                    (f == 11 ? graph_inputs : graph_outputs).push_back(name);
                }
                if (!ok) {
                    *err = "[" + path + "]: malformed " + (f == 1 ? "NodeProto" : f == 5 ? "TensorProto" : "ValueInfoProto");
                    return false;
                }
            } else if (!r.Skip(w)) {
                break;
            }
        }
        if (!r.ok()) {
            *err = "[" + path + "]: malformed GraphProto";
            return false;
        }
    }

    // resolve payloads
    initializers.reserve(raw.size());
    for (auto& rt : raw) {
        Tensor& t = rt.t;
        const uint32_t es = ElementSize(t.data_type);
        const uint64_t n = t.NumElements();
        if (t.external) {
            if (t.location.empty() || t.location.find("..") != std::string::npos || t.location[0] == '/') {
                *err = "initializer [" + t.name + "]: bad external data location [" + t.location + "]";
                return false;
            }
            const uint8_t* p = nullptr;
            uint64_t len = 0;
            if (!MapFile(dir + "/" + t.location, &p, &len, err)) {
                *err = "initializer [" + t.name + "]: " + *err;
                return false;
            }
            const uint64_t off = (uint64_t)rt.ext_offset;
            const uint64_t want = rt.ext_length >= 0 ? (uint64_t)rt.ext_length : len - std::min(off, len);
            if (off > len || want > len - off) {
                *err = "initializer [" + t.name + "]: external data [" + t.location + "] is shorter than offset + length";
                return false;
            }
            t.data = p + off;
            t.bytes = want;
        } else if (!rt.has_raw) {
            if (!rt.float_data.empty() && t.data_type == DT_FLOAT) {
                t.owned.resize(rt.float_data.size() * 4);
                memcpy(t.owned.data(), rt.float_data.data(), t.owned.size());
            } else if (!rt.int64_data.empty() && t.data_type == DT_INT64) {
                t.owned.resize(rt.int64_data.size() * 8);
                memcpy(t.owned.data(), rt.int64_data.data(), t.owned.size());
            } else if (!rt.int32_data.empty() && es && es <= 4) {
                t.owned.resize(rt.int32_data.size() * es);
                for (size_t i = 0; i < rt.int32_data.size(); ++i) {
                    const uint32_t u = (uint32_t)rt.int32_data[i];
                    memcpy(t.owned.data() + i * es, &u, es); // little endian: low bytes carry the value
                }
            }
            t.data = t.owned.data();
            t.bytes = t.owned.size();
        }
        if (es && t.bytes != n * es) {
            *err = "initializer [" + t.name + "]: " + std::to_string(t.bytes) + " payload bytes for " + std::to_string(n) +
                " elements of " + std::to_string(es) + " bytes";
            return false;
        }
        initializers.push_back(std::move(t));
        Tensor& kept = initializers.back();
        if (!kept.owned.empty()) kept.data = kept.owned.data(); // the vector moved
    }
    for (size_t i = 0; i < initializers.size(); ++i) by_name_[initializers[i].name] = i;
    return true;
}

} // namespace b2onnx
