// Robustness of the two file parsers that read files from disk -- the ONNX model-slice loader and the sentencepiece model
// loader -- against damaged input: every file given on the command line is loaded `rounds` times with random byte
// mutations / truncations; built with -fsanitize=address,undefined (Makefile: build/fuzz_loaders).  A parser may refuse a
// file, it must never crash, read out of bounds or loop.   usage: fuzz_loaders <rounds> <seed> <model.onnx | *.model>...
#include "../src/pmx_llama.h"
#include "sentencepiece_processor.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <fstream>
#include <random>
#include <sstream>

extern "C" int32_t b2pplnn_inspect_model(const char* path, char* out, uint64_t cap);

static std::string Slurp(const char* p) {
    std::ifstream f(p, std::ios::binary);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const int rounds = atoi(argv[1]);
    std::mt19937_64 rng(strtoull(argv[2], nullptr, 0));
    std::vector<char> json(1 << 20);
    char tmpl[] = "/tmp/b2fuzzXXXXXX";
    if (!mkdtemp(tmpl)) return 2;
    const std::string dir = std::string(tmpl) + "/model_slice_0";
    if (system(("mkdir -p " + dir).c_str()) != 0) return 2;
    long loaded = 0, refused = 0;
    for (int a = 3; a < argc; ++a) {
        const std::string orig = Slurp(argv[a]);
        const bool onnx = std::string(argv[a]).size() > 5 && std::string(argv[a]).substr(std::string(argv[a]).size() - 5) == ".onnx";
        for (int r = 0; r < rounds; ++r) {
            std::string s = orig;
            const int kind = (int)(rng() % 4);
            if (kind == 0 && !s.empty()) s.resize(rng() % s.size());                       // truncation
            const int flips = kind == 3 ? 0 : 1 + (int)(rng() % 8);
            const size_t head = std::min<size_t>(s.size(), (r & 1) ? 4096 : s.size());     // half of the rounds: damage the headers
            for (int i = 0; i < flips && head; ++i) s[rng() % head] = (char)(rng() & 0xff);
            if (kind == 3 && s.size() > 16) {                                               // a huge length prefix somewhere
                const size_t at = rng() % (s.size() - 8);
                memcpy(&s[at], "\xff\xff\xff\xff\xff\xff\xff\x7f", 8);
            }
            if (onnx) {
                const std::string path = dir + "/model.onnx";
                std::ofstream(path, std::ios::binary).write(s.data(), (std::streamsize)s.size());
                (b2pplnn_inspect_model(path.c_str(), json.data(), json.size()) == 0 ? loaded : refused)++;
            } else {
                sentencepiece::SentencePieceProcessor sp;
                if (sp.LoadFromSerializedProto(s).ok()) {
                    ++loaded;
                    std::vector<int> ids;
                    sp.Encode("hello world \xe5\x8c\x97\xe4\xba\xac caf\xc3\xa9 \xff", &ids);
                    std::string text;
                    sp.Decode(ids, &text);
                } else {
                    ++refused;
                }
            }
        }
    }
    if (system(("rm -rf " + std::string(tmpl)).c_str()) != 0) return 2;
    printf("fuzz ok: %ld loaded, %ld refused\n", loaded, refused);
    return 0;
}
