// Minimal protobuf wire-format reader (proto3 encoding rules) -- just enough to walk an ONNX ModelProto
// without libprotobuf / protoc, neither of which exists in this image.  The reference generates
// onnx.pb.{h,cc} from src/onnx/onnx.proto with protoc (CMakeLists.txt:52-67) for ppl.nn's model loader; the field
// numbers used by onnx_model.cc are the ones of that file.
//
// Wire types: 0 varint, 1 64-bit, 2 length-delimited, 5 32-bit (3/4 = groups, not used by ONNX -> error).
#ifndef B2_ONNX_WIRE_H_
#define B2_ONNX_WIRE_H_

#include <stddef.h>
#include <stdint.h>
#include <string.h>

namespace b2onnx {

struct View {
    const uint8_t* p = nullptr;
    uint64_t n = 0;
};

class WireReader {
public:
    WireReader(const uint8_t* p, uint64_t n) : p_(p), end_(p + n) {}
    explicit WireReader(const View& v) : p_(v.p), end_(v.p + v.n) {}

    // advances to the next field; false at the end of the message OR on a malformed tag (check ok())
    bool Next(uint32_t* field, uint32_t* wire) {
        if (!ok_ || p_ >= end_) return false;
        uint64_t tag = 0;
        if (!Varint(&tag)) return false;
        *field = (uint32_t)(tag >> 3);
        *wire = (uint32_t)(tag & 7);
        if (*field == 0) ok_ = false;
        return ok_;
    }
    bool Varint(uint64_t* out) {
        uint64_t v = 0;
        for (int shift = 0; shift < 64; shift += 7) {
            if (p_ >= end_) return Fail();
            const uint8_t b = *p_++;
            v |= (uint64_t)(b & 0x7f) << shift;
            if (!(b & 0x80)) {
                *out = v;
                return true;
            }
        }
        return Fail(); // more than 10 bytes
    }
    bool Fixed32(uint32_t* out) {
        if ((uint64_t)(end_ - p_) < 4) return Fail();
        memcpy(out, p_, 4);
        p_ += 4;
        return true;
    }
    bool Fixed64(uint64_t* out) {
        if ((uint64_t)(end_ - p_) < 8) return Fail();
        memcpy(out, p_, 8);
        p_ += 8;
        return true;
    }
    bool Bytes(View* out) {
        uint64_t n = 0;
        if (!Varint(&n)) return false;
        if (n > (uint64_t)(end_ - p_)) return Fail();
        out->p = p_;
        out->n = n;
        p_ += n;
        return true;
    }
    bool Skip(uint32_t wire) {
        uint64_t u64;
        uint32_t u32;
        View v;
        switch (wire) {
            case 0: return Varint(&u64);
            case 1: return Fixed64(&u64);
            case 2: return Bytes(&v);
            case 5: return Fixed32(&u32);
            default: return Fail();
        }
    }
    bool ok() const {
        return ok_;
    }
    bool AtEnd() const {
        return p_ >= end_;
    }

private:
    bool Fail() {
        ok_ = false;
        return false;
    }
    const uint8_t* p_;
    const uint8_t* end_;
    bool ok_ = true;
};

} // namespace b2onnx
#endif
