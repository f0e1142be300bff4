// <sentencepiece_processor.h> for the reference's tokenizer front end (src/tokenizer/tokenizer_impl_sp.h:22-68):
// a from-scratch SentencePiece *inference* implementation (host/src/sentencepiece.cc) -- the OpenPPL sentencepiece
// fork the reference fetches (cmake/deps.cmake) is absent in this image and there is no network.
//
// What it reads and does ("tokenizer front end"):
//   * `tokenizer.model` = a serialized sentencepiece ModelProto (pieces with score / type, TrainerSpec, NormalizerSpec),
//     parsed with the same protobuf wire reader as the ONNX loader (host/src/onnx_wire.h);
//   * BPE (LLaMA / LLaMA-2: model_type BPE, byte_fallback, identity normaliser, dummy prefix) and unigram models;
//     normalisation = dummy prefix / extra-whitespace removal / whitespace escaping; a compiled character map
//     (nmt_nfkc ...) is NOT interpreted -- Load() refuses such a model instead of tokenising differently;
//   * Encode -> ids, Decode <- ids (byte pieces re-assembled to UTF-8, control pieces invisible, unknown -> " ⁇ ").
// Parity is pinned against the official `sentencepiece` Python package on models trained in the test
// (tests/test_tokenizer_cpu.py): identical ids and identical decoded text.
//
// A file whose first bytes are "b2llm-byte-level-tokenizer" selects a built-in byte-level vocabulary (id = 3 + byte;
// 0 <unk>, 1 <s>, 2 </s>) -- a test aid for running `offline_inference` without a trained model.
#ifndef B2LLM_SENTENCEPIECE_PROCESSOR_H_
#define B2LLM_SENTENCEPIECE_PROCESSOR_H_

#include "absl/strings/string_view.h"

#include <memory>
#include <string>
#include <vector>

namespace sentencepiece {

namespace util {
class Status final {
public:
    Status() {}
    explicit Status(const std::string& err) : ok_(false), msg_(err) {}
    bool ok() const {
        return ok_;
    }
    std::string ToString() const {
        return msg_;
    }

private:
    bool ok_ = true;
    std::string msg_;
};
} // namespace util

class SentencePieceProcessor final {
public:
    SentencePieceProcessor();
    ~SentencePieceProcessor();
    SentencePieceProcessor(const SentencePieceProcessor&) = delete;
    SentencePieceProcessor& operator=(const SentencePieceProcessor&) = delete;

    util::Status Load(absl::string_view filename);
    util::Status LoadFromSerializedProto(absl::string_view serialized);

    util::Status Encode(absl::string_view input, std::vector<int>* ids) const;
    util::Status Encode(absl::string_view input, std::vector<std::string>* pieces) const;
    // (ids, len) form: the OpenPPL fork's overload the reference calls (tokenizer_impl_sp.h:53)
    util::Status Decode(const int* ids, unsigned int len, std::string* out) const;
    util::Status Decode(const std::vector<int>& ids, std::string* out) const;

    int GetPieceSize() const;
    int PieceToId(absl::string_view piece) const;
    const std::string& IdToPiece(int id) const;
    float GetScore(int id) const;
    bool IsUnknown(int id) const;
    bool IsControl(int id) const;
    bool IsUnused(int id) const;
    bool IsByte(int id) const;
    int unk_id() const;
    int bos_id() const;
    int eos_id() const;
    int pad_id() const;

private:
    struct Impl;
    std::unique_ptr<Impl> impl_;
};

} // namespace sentencepiece

#endif
