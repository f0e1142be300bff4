// Stand-in for <sentencepiece_processor.h> (the OpenPPL sentencepiece fork is absent in this image and there
// is no network).  OUT OF SCOPE component (SURVEY.md section 2.1 "Tokenizer"): the hot path is driven with
// token-in/out requests that bypass the tokenizer (llm_generator.cc:790-801).  This stub lets the reference's
// tokenizer headers compile and lets `offline_inference` run its four text prompts end to end with a
// byte-level vocabulary: piece id = 3 + byte (0 <unk>, 1 <s>, 2 </s>), "Load" accepts any readable file.
#ifndef B2LLM_SHIM_SENTENCEPIECE_PROCESSOR_H_
#define B2LLM_SHIM_SENTENCEPIECE_PROCESSOR_H_

#include "absl/strings/string_view.h"

#include <fstream>
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
    util::Status Load(absl::string_view filename) {
        std::ifstream ifs{std::string(filename)};
        if (!ifs.is_open()) {
            return util::Status("cannot open tokenizer model [" + std::string(filename) + "]");
        }
        return util::Status();
    }
    util::Status Encode(absl::string_view input, std::vector<int>* ids) const {
        ids->clear();
        for (unsigned char c : input) {
            ids->push_back(3 + (int)c);
        }
        return util::Status();
    }
    util::Status Decode(const int* ids, unsigned int len, std::string* out) const {
        out->clear();
        for (unsigned int i = 0; i < len; ++i) {
            const int id = ids[i];
            if (id >= 3 && id < 3 + 256) {
                out->push_back((char)(id - 3));
            } else if (id >= 3 + 256) { // outside the byte range: printable placeholder
                *out += "<" + std::to_string(id) + ">";
            }
        }
        return util::Status();
    }
    util::Status Decode(const std::vector<int>& ids, std::string* out) const {
        return Decode(ids.data(), (unsigned int)ids.size(), out);
    }
    std::string IdToPiece(int id) const {
        if (id >= 3 && id < 3 + 256) {
            return std::string(1, (char)(id - 3));
        }
        return id == 1 ? "<s>" : id == 2 ? "</s>" : "<unk>";
    }
    int GetPieceSize() const {
        return 32000;
    }
    int bos_id() const {
        return 1;
    }
    int eos_id() const {
        return 2;
    }
    int pad_id() const {
        return -1;
    }
    int unk_id() const {
        return 0;
    }
};

} // namespace sentencepiece

#endif
