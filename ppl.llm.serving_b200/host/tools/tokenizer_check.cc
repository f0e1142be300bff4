// The reference's own tokenizer classes (src/tokenizer/tokenizer_factory.h -> TokenizerImplSP -> LlamaTokenizer),
// compiled in place against this repo's sentencepiece implementation:
//     tokenizer_check <tokenizer.model> <text>
// prints   ids <bos> <id> ...        Tokenizer::Encode  (llama_tokenizer.h:36-39 prepends BOS)
//          bos <id> eos <id>
//          text <decoded>            Tokenizer::Decode of all ids at once
//          stream <decoded>          Tokenizer::Decode one token at a time, concatenated -- how the generator's
//                                    detokenize thread emits text (tokenizer_impl_sp.h:52-59 restores the leading space)
#include "tokenizer/tokenizer_factory.h"

#include <stdio.h>
#include <string.h>

#include <memory>

int main(int argc, char** argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s <tokenizer.model> <text>\n", argv[0]);
        return 2;
    }
    std::unique_ptr<ppl::llm::Tokenizer> tok(ppl::llm::TokenizerFactory::Create("llama", "sentencepiece", argv[1], ""));
    if (!tok) return 1;
    std::vector<int> ids;
    tok->Encode(argv[2], (uint32_t)strlen(argv[2]), &ids);
    printf("ids");
    for (int id : ids) printf(" %d", id);
    printf("\nbos %d eos %d\n", tok->GetBosId(), tok->GetEosId());
    std::string text;
    tok->Decode(ids.data(), (uint32_t)ids.size(), &text);
    printf("text %s\n", text.c_str());
    std::string stream;
    for (int id : ids) {
        std::string piece;
        tok->Decode(&id, 1, &piece);
        stream += piece;
    }
    printf("stream %s\n", stream.c_str());
    return 0;
}
