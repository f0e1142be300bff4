// SentencePiece inference (see include/sentencepiece_processor.h).  Written from the published algorithm of
// google/sentencepiece (Kudo & Richardson 2018; the pinned-by-nothing OpenPPL fork the reference fetches is absent here):
//   model file      sentencepiece_model.proto -- field numbers as in the reference's generated
//                   src/generated/onnx/v3.1.0/sentencepiece_model.pb.h (ModelProto.pieces = 1, trainer_spec = 2,
//                   normalizer_spec = 3; SentencePiece.{piece = 1, score = 2, type = 3}; TrainerSpec.{model_type = 3,
//                   treat_whitespace_as_suffix = 24, byte_fallback = 35, unk_id = 40 .. pad_id = 43, unk_surface = 44,
//                   unk_piece = 45 .. pad_piece = 48}; NormalizerSpec.{name = 1, precompiled_charsmap = 2,
//                   add_dummy_prefix = 3, remove_extra_whitespaces = 4, escape_whitespaces = 5})
//   normalisation   dummy prefix, extra-whitespace removal, U+2581 escaping; invalid UTF-8 bytes -> U+FFFD
//   BPE             agenda of adjacent pairs ordered by (score desc, position asc); unused pieces re-segmented
//   unigram         Viterbi over UTF-8 positions, unknown characters at min_score - 10
//   post-processing byte fallback (<0xXX> pieces) or merging of consecutive unknowns
//   decoding        control pieces invisible, unknown -> unk_surface, byte runs -> UTF-8 (invalid -> U+FFFD),
//                   leading U+2581 of the first visible piece dropped (dummy prefix)
// Checked against the official Python package in tests/test_tokenizer_cpu.py.
#include "sentencepiece_processor.h"

#include "onnx_wire.h"

#include <float.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <fstream>
#include <functional>
#include <queue>
#include <sstream>
#include <unordered_map>
#include <unordered_set>

namespace sentencepiece {

namespace {

using b2onnx::View;
using b2onnx::WireReader;

enum PieceType { NORMAL = 1, UNKNOWN = 2, CONTROL = 3, USER_DEFINED = 4, UNUSED = 5, BYTE = 6 };
enum ModelType { UNIGRAM = 1, BPE = 2, WORD = 3, CHAR = 4 };

const char kSpaceSymbol[] = "\xe2\x96\x81";   // U+2581
const char kReplacement[] = "\xef\xbf\xbd";   // U+FFFD
const char kByteLevelMagic[] = "b2llm-byte-level-tokenizer";

struct Piece {
    std::string piece;
    float score = 0.f;
    int type = NORMAL;
};

int OneCharLen(const char* src) {
    return "\1\1\1\1\1\1\1\1\1\1\1\1\2\2\3\4"[(*src & 0xFF) >> 4];
}

bool IsTrail(unsigned char c) {
    return (c & 0xC0) == 0x80;
}

// length of the valid UTF-8 character at the head of [p, p + n); 0 when it is not valid (as sentencepiece's
// DecodeUTF8: overlong forms, surrogates and > U+10FFFF are invalid)
int ValidUtf8Len(const unsigned char* p, size_t n) {
    if (n == 0) return 0;
    if (p[0] < 0x80) return 1;
    if (n >= 2 && (p[0] & 0xE0) == 0xC0) {
        const uint32_t cp = ((p[0] & 0x1F) << 6) | (p[1] & 0x3F);
        return IsTrail(p[1]) && cp >= 0x80 ? 2 : 0;
    }
    if (n >= 3 && (p[0] & 0xF0) == 0xE0) {
        const uint32_t cp = ((p[0] & 0x0F) << 12) | ((p[1] & 0x3F) << 6) | (p[2] & 0x3F);
        return IsTrail(p[1]) && IsTrail(p[2]) && cp >= 0x800 && (cp < 0xD800 || cp > 0xDFFF) ? 3 : 0;
    }
    if (n >= 4 && (p[0] & 0xF8) == 0xF0) {
        const uint32_t cp = ((p[0] & 0x07) << 18) | ((p[1] & 0x3F) << 12) | ((p[2] & 0x3F) << 6) | (p[3] & 0x3F);
        return IsTrail(p[1]) && IsTrail(p[2]) && IsTrail(p[3]) && cp >= 0x10000 && cp <= 0x10FFFF ? 4 : 0;
    }
    return 0;
}

std::string ByteToPiece(unsigned char b) {
    char buf[8];
    snprintf(buf, sizeof(buf), "<0x%02X>", b);
    return buf;
}

int PieceToByte(const std::string& p) { // "<0xAB>" -> 0xAB, -1 otherwise
    if (p.size() != 6 || p[0] != '<' || p[1] != '0' || p[2] != 'x' || p[5] != '>') return -1;
    auto hex = [](char c) { return c >= '0' && c <= '9' ? c - '0' : (c >= 'A' && c <= 'F' ? c - 'A' + 10 : -1); };
    const int h = hex(p[3]), l = hex(p[4]);
    return h < 0 || l < 0 ? -1 : h * 16 + l;
}

bool ConsumePrefix(std::string_view* s, std::string_view prefix) {
    if (s->substr(0, prefix.size()) != prefix) return false;
    s->remove_prefix(prefix.size());
    return true;
}

bool EndsWith(std::string_view s, std::string_view suffix) {
    return s.size() >= suffix.size() && s.substr(s.size() - suffix.size()) == suffix;
}

} // namespace

struct SentencePieceProcessor::Impl {
    std::vector<Piece> pieces;
    std::unordered_map<std::string, int> normal_ids;   // NORMAL, USER_DEFINED, UNUSED (the model's "pieces_")
    std::unordered_map<std::string, int> reserved_ids; // CONTROL, UNKNOWN, BYTE
    std::unordered_set<std::string> user_defined;
    size_t max_user_len = 0, max_piece_len = 0;
    int model_type = UNIGRAM;
    bool byte_fallback = false, treat_ws_as_suffix = false;
    bool add_dummy_prefix = true, remove_extra_ws = true, escape_ws = true;
    int unk = -1;
    std::string unk_piece = "<unk>", bos_piece = "<s>", eos_piece = "</s>", pad_piece = "<pad>";
    std::string unk_surface = " \xE2\x81\x87 ";
    float min_score = FLT_MAX, max_score = -FLT_MAX;
    bool byte_level = false; // built-in byte vocabulary (test aid)
    bool loaded = false;
    std::string empty;

    int Size() const {
        return byte_level ? 3 + 256 : (int)pieces.size();
    }
    int PieceToId(std::string_view p) const {
        const std::string key(p);
        auto it = reserved_ids.find(key);
        if (it != reserved_ids.end()) return it->second;
        auto it2 = normal_ids.find(key);
        if (it2 != normal_ids.end()) return it2->second;
        return unk;
    }
    int Type(int id) const {
        return id >= 0 && id < (int)pieces.size() ? pieces[id].type : 0;
    }

    util::Status Parse(std::string_view blob);
    void Normalize(std::string_view input, std::string* out) const;
    size_t PrefixMatch(std::string_view s, bool* found) const;
    void EncodeBpe(std::string_view norm, std::vector<std::pair<std::string_view, int>>* out) const;
    void EncodeUnigram(std::string_view norm, std::vector<std::pair<std::string_view, int>>* out) const;
    void EncodeChar(std::string_view norm, std::vector<std::pair<std::string_view, int>>* out) const;
    void EncodePieces(std::string_view input, std::vector<std::pair<std::string, int>>* out) const;
};

util::Status SentencePieceProcessor::Impl::Parse(std::string_view blob) {
    pieces.clear();
    normal_ids.clear();
    reserved_ids.clear();
    user_defined.clear();
    WireReader r((const uint8_t*)blob.data(), blob.size());
    uint32_t f, w;
    std::string charsmap;
    while (r.Next(&f, &w)) {
        View v;
        if (w != 2 || !(f == 1 || f == 2 || f == 3)) {
            if (!r.Skip(w)) break;
            continue;
        }
        if (!r.Bytes(&v)) break;
        WireReader m(v);
        uint32_t mf, mw;
        if (f == 1) { // SentencePiece
            Piece p;
            while (m.Next(&mf, &mw)) {
                View s;
                uint64_t u;
                uint32_t u32;
                if (mf == 1 && mw == 2) {
                    if (!m.Bytes(&s)) break;
                    p.piece.assign((const char*)s.p, s.n);
                } else if (mf == 2 && mw == 5) {
                    if (!m.Fixed32(&u32)) break;
                    memcpy(&p.score, &u32, 4);
                } else if (mf == 3 && mw == 0) {
                    if (!m.Varint(&u)) break;
                    p.type = (int)u;
                } else if (!m.Skip(mw)) {
                    break;
                }
            }
            if (!m.ok()) return util::Status("malformed SentencePiece entry in the model file");
            pieces.push_back(std::move(p));
        } else if (f == 2) { // TrainerSpec
            while (m.Next(&mf, &mw)) {
                View s;
                uint64_t u;
                if (mw == 0) {
                    if (!m.Varint(&u)) break;
                    if (mf == 3) model_type = (int)u;
                    else if (mf == 24) treat_ws_as_suffix = u != 0;
                    else if (mf == 35) byte_fallback = u != 0;
                } else if (mw == 2) {
                    if (!m.Bytes(&s)) break;
                    const std::string str((const char*)s.p, s.n);
                    if (mf == 44) unk_surface = str;
                    else if (mf == 45) unk_piece = str;
                    else if (mf == 46) bos_piece = str;
                    else if (mf == 47) eos_piece = str;
                    else if (mf == 48) pad_piece = str;
                } else if (!m.Skip(mw)) {
                    break;
                }
            }
            if (!m.ok()) return util::Status("malformed TrainerSpec in the model file");
        } else { // NormalizerSpec
            while (m.Next(&mf, &mw)) {
                View s;
                uint64_t u;
                if (mw == 0) {
                    if (!m.Varint(&u)) break;
                    if (mf == 3) add_dummy_prefix = u != 0;
                    else if (mf == 4) remove_extra_ws = u != 0;
                    else if (mf == 5) escape_ws = u != 0;
                } else if (mw == 2) {
                    if (!m.Bytes(&s)) break;
                    if (mf == 2) charsmap.assign((const char*)s.p, s.n);
                } else if (!m.Skip(mw)) {
                    break;
                }
            }
            if (!m.ok()) return util::Status("malformed NormalizerSpec in the model file");
        }
    }
    if (!r.ok() || pieces.empty()) return util::Status("not a sentencepiece model (malformed protobuf or no pieces)");
    if (!charsmap.empty()) {
        return util::Status("the model normalises with a compiled character map (e.g. nmt_nfkc); only identity-normalised "
                            "models (LLaMA, LLaMA-2: normalization_rule_name = identity) are supported");
    }
    if (model_type != BPE && model_type != UNIGRAM && model_type != CHAR) {
        return util::Status("model_type " + std::to_string(model_type) + " (word) is not supported: BPE, unigram and char are");
    }
    unk = -1;
    min_score = FLT_MAX;
    max_score = -FLT_MAX;
    int bytes_seen = 0;
    for (int i = 0; i < (int)pieces.size(); ++i) {
        const Piece& p = pieces[i];
        if (p.piece.empty()) return util::Status("piece " + std::to_string(i) + " is empty");
        const bool normal = p.type == NORMAL || p.type == USER_DEFINED || p.type == UNUSED;
        auto& map = normal ? normal_ids : reserved_ids;
        if (normal_ids.count(p.piece) || reserved_ids.count(p.piece)) return util::Status("piece [" + p.piece + "] is defined twice");
        map[p.piece] = i;
        if (p.type == UNKNOWN) {
            if (unk >= 0) return util::Status("unk is defined twice");
            unk = i;
        }
        if (p.type == USER_DEFINED) {
            user_defined.insert(p.piece);
            max_user_len = std::max(max_user_len, p.piece.size());
        }
        if (p.type == BYTE) {
            if (!byte_fallback) return util::Status("byte piece [" + p.piece + "] but byte_fallback is off");
            if (PieceToByte(p.piece) < 0) return util::Status("byte piece [" + p.piece + "] is not of the form <0xXX>");
            ++bytes_seen;
        }
        if (p.type == NORMAL) {
            min_score = std::min(min_score, p.score);
            max_score = std::max(max_score, p.score);
        }
        if (normal) max_piece_len = std::max(max_piece_len, p.piece.size());
    }
    if (unk < 0) return util::Status("unk is not defined");
    if (byte_fallback && bytes_seen != 256) return util::Status("byte_fallback is on but the model has " + std::to_string(bytes_seen) + " byte pieces, not 256");
    if (min_score == FLT_MAX) min_score = max_score = 0.f;
    loaded = true;
    return util::Status();
}

// google/sentencepiece normalizer.cc Normalizer::Normalize with an empty character map
void SentencePieceProcessor::Impl::Normalize(std::string_view input, std::string* out) const {
    out->clear();
    auto next_char = [](std::string_view s, std::string_view* piece) -> size_t { // NormalizePrefix: (normalised, consumed)
        const int n = ValidUtf8Len((const unsigned char*)s.data(), s.size());
        if (n == 0) {
            *piece = std::string_view(kReplacement, 3);
            return 1;
        }
        *piece = s.substr(0, n);
        return (size_t)n;
    };
    std::string_view sp;
    if (remove_extra_ws) {
        while (!input.empty()) {
            const size_t used = next_char(input, &sp);
            if (sp != " ") break;
            input.remove_prefix(used);
        }
    }
    if (input.empty()) return;
    out->reserve(input.size() * 3);
    const std::string_view space = escape_ws ? std::string_view(kSpaceSymbol, 3) : std::string_view(" ");
    if (!treat_ws_as_suffix && add_dummy_prefix) out->append(space);
    bool is_prev_space = remove_extra_ws;
    while (!input.empty()) {
        const size_t used = next_char(input, &sp);
        while (is_prev_space && ConsumePrefix(&sp, " ")) {
        }
        if (!sp.empty()) {
            for (char c : sp) {
                if (escape_ws && c == ' ') out->append(kSpaceSymbol, 3);
                else out->push_back(c);
            }
            is_prev_space = EndsWith(sp, " ");
        }
        input.remove_prefix(used);
        if (!remove_extra_ws) is_prev_space = false;
    }
    if (remove_extra_ws) {
        while (EndsWith(*out, space)) out->resize(out->size() - space.size());
    }
    if (treat_ws_as_suffix && add_dummy_prefix) out->append(space);
}

// longest user-defined symbol at the head of s, else one UTF-8 character
size_t SentencePieceProcessor::Impl::PrefixMatch(std::string_view s, bool* found) const {
    *found = false;
    for (size_t n = std::min(max_user_len, s.size()); n > 0; --n) {
        if (user_defined.count(std::string(s.substr(0, n)))) {
            *found = true;
            return n;
        }
    }
    return std::min<size_t>(OneCharLen(s.data()), s.size());
}

// google/sentencepiece bpe_model.cc Model::Encode (no dropout)
void SentencePieceProcessor::Impl::EncodeBpe(std::string_view norm, std::vector<std::pair<std::string_view, int>>* out) const {
    struct Symbol {
        int prev, next;
        bool freeze;
        std::string_view piece;
    };
    struct Pair {
        int left, right;
        float score;
        size_t size;
    };
    struct Cmp { // top = highest score, then leftmost
        bool operator()(const Pair& a, const Pair& b) const {
            return a.score < b.score || (a.score == b.score && a.left > b.left);
        }
    };
    std::vector<Symbol> sym;
    while (!norm.empty()) {
        bool found;
        const size_t n = PrefixMatch(norm, &found);
        Symbol s;
        s.piece = norm.substr(0, n);
        s.freeze = found;
        s.prev = (int)sym.size() - 1;
        norm.remove_prefix(n);
        s.next = norm.empty() ? -1 : (int)sym.size() + 1;
        sym.push_back(s);
    }
    if (sym.empty()) return;
    std::priority_queue<Pair, std::vector<Pair>, Cmp> agenda;
    std::unordered_map<std::string_view, std::pair<std::string_view, std::string_view>> rev_merge;
    auto maybe_add = [&](int left, int right) {
        if (left == -1 || right == -1 || sym[left].freeze || sym[right].freeze) return;
        const std::string_view piece(sym[left].piece.data(), sym[left].piece.size() + sym[right].piece.size());
        auto it = normal_ids.find(std::string(piece));
        if (it == normal_ids.end()) return;
        agenda.push(Pair{left, right, pieces[it->second].score, piece.size()});
        if (pieces[it->second].type == UNUSED) rev_merge[piece] = std::make_pair(sym[left].piece, sym[right].piece);
    };
    for (size_t i = 1; i < sym.size(); ++i) maybe_add((int)i - 1, (int)i);
    while (!agenda.empty()) {
        const Pair top = agenda.top();
        agenda.pop();
        if (sym[top.left].piece.empty() || sym[top.right].piece.empty() ||
            sym[top.left].piece.size() + sym[top.right].piece.size() != top.size)
            continue;
        sym[top.left].piece = std::string_view(sym[top.left].piece.data(), top.size);
        sym[top.left].next = sym[top.right].next;
        if (sym[top.right].next >= 0) sym[sym[top.right].next].prev = top.left;
        sym[top.right].piece = std::string_view();
        maybe_add(sym[top.left].prev, top.left);
        maybe_add(top.left, sym[top.left].next);
    }
    std::function<void(std::string_view)> resegment = [&](std::string_view w) {
        const int id = PieceToId(w);
        if (id == -1 || Type(id) != UNUSED) {
            out->emplace_back(w, id);
            return;
        }
        auto it = rev_merge.find(w);
        if (it == rev_merge.end()) { // unused piece that no merge produced (a single character): keep it
            out->emplace_back(w, id);
            return;
        }
        resegment(it->second.first);
        resegment(it->second.second);
    };
    for (int i = 0; i != -1; i = sym[i].next) resegment(sym[i].piece);
}

// google/sentencepiece unigram_model.cc Model::EncodeOptimized
void SentencePieceProcessor::Impl::EncodeUnigram(std::string_view norm, std::vector<std::pair<std::string_view, int>>* out) const {
    struct Node {
        int id = -1;
        float score = 0.f;
        int starts_at = -1;
    };
    const int size = (int)norm.size();
    if (size == 0) return;
    const float unk_score = min_score - 10.0f;
    std::vector<Node> best(size + 1);
    int starts_at = 0;
    std::string key;
    while (starts_at < size) {
        const float till_here = best[starts_at].score;
        bool has_single = false;
        const int mblen = std::min<int>(OneCharLen(norm.data() + starts_at), size - starts_at);
        const int max_end = std::min<int>(size, starts_at + (int)max_piece_len);
        for (int end = starts_at + 1; end <= max_end; ++end) { // increasing length == trie traversal order
            key.assign(norm.data() + starts_at, end - starts_at);
            auto it = normal_ids.find(key);
            if (it == normal_ids.end()) continue;
            const int id = it->second;
            if (pieces[id].type == UNUSED) continue;
            Node& target = best[end];
            const int length = end - starts_at;
            const double score = pieces[id].type == USER_DEFINED ? (length * max_score - 0.1) : pieces[id].score;
            const double cand = score + till_here;
            if (target.starts_at == -1 || cand > target.score) {
                target.score = (float)cand;
                target.starts_at = starts_at;
                target.id = id;
            }
            if (!has_single && length == mblen) has_single = true;
        }
        if (!has_single) {
            Node& target = best[starts_at + mblen];
            const float cand = unk_score + till_here;
            if (target.starts_at == -1 || cand > target.score) {
                target.score = cand;
                target.starts_at = starts_at;
                target.id = unk;
            }
        }
        starts_at += mblen;
    }
    std::vector<std::pair<std::string_view, int>> rev;
    int ends_at = size;
    while (ends_at > 0) {
        const Node& n = best[ends_at];
        rev.emplace_back(norm.substr(n.starts_at, ends_at - n.starts_at), n.id);
        ends_at = n.starts_at;
    }
    out->insert(out->end(), rev.rbegin(), rev.rend());
}

// google/sentencepiece char_model.cc: one piece per character (or user-defined symbol)
void SentencePieceProcessor::Impl::EncodeChar(std::string_view norm, std::vector<std::pair<std::string_view, int>>* out) const {
    while (!norm.empty()) {
        bool found;
        const size_t n = PrefixMatch(norm, &found);
        out->emplace_back(norm.substr(0, n), PieceToId(norm.substr(0, n)));
        norm.remove_prefix(n);
    }
}

// Encode + sentencepiece_processor.cc PopulateSentencePieceText: byte fallback, or merging of runs of unknowns
void SentencePieceProcessor::Impl::EncodePieces(std::string_view input, std::vector<std::pair<std::string, int>>* out) const {
    out->clear();
    if (byte_level) {
        for (unsigned char c : input) out->emplace_back(std::string(1, (char)c), 3 + (int)c);
        return;
    }
    std::string norm;
    Normalize(input, &norm);
    std::vector<std::pair<std::string_view, int>> raw;
    if (model_type == BPE) EncodeBpe(norm, &raw);
    else if (model_type == UNIGRAM) EncodeUnigram(norm, &raw);
    else EncodeChar(norm, &raw);
    bool prev_unk = false;
    for (const auto& p : raw) {
        const bool is_unk = p.second == unk;
        if (is_unk && byte_fallback) {
            for (unsigned char b : p.first) {
                const std::string bp = ByteToPiece(b);
                out->emplace_back(bp, PieceToId(bp));
            }
        } else if (is_unk && prev_unk) {
            out->back().first.append(p.first); // one unknown token for the whole run
        } else {
            out->emplace_back(std::string(p.first), p.second);
        }
        prev_unk = is_unk;
    }
}

// ------------------------------------------------------------------------------------------------ public
SentencePieceProcessor::SentencePieceProcessor() : impl_(new Impl()) {}
SentencePieceProcessor::~SentencePieceProcessor() {}

util::Status SentencePieceProcessor::Load(absl::string_view filename) {
    std::ifstream ifs{std::string(filename), std::ios::binary};
    if (!ifs.is_open()) return util::Status("cannot open tokenizer model [" + std::string(filename) + "]");
    std::stringstream ss;
    ss << ifs.rdbuf();
    const std::string blob = ss.str();
    auto st = LoadFromSerializedProto(blob);
    if (!st.ok()) return util::Status("tokenizer model [" + std::string(filename) + "]: " + st.ToString());
    return st;
}

util::Status SentencePieceProcessor::LoadFromSerializedProto(absl::string_view serialized) {
    impl_.reset(new Impl());
    if (serialized.substr(0, sizeof(kByteLevelMagic) - 1) == kByteLevelMagic) {
        impl_->byte_level = true;
        impl_->loaded = true;
        return util::Status();
    }
    return impl_->Parse(serialized);
}

util::Status SentencePieceProcessor::Encode(absl::string_view input, std::vector<int>* ids) const {
    ids->clear();
    if (!impl_->loaded) return util::Status("model is not loaded");
    std::vector<std::pair<std::string, int>> pcs;
    impl_->EncodePieces(input, &pcs);
    for (const auto& p : pcs) ids->push_back(p.second);
    return util::Status();
}

util::Status SentencePieceProcessor::Encode(absl::string_view input, std::vector<std::string>* out) const {
    out->clear();
    if (!impl_->loaded) return util::Status("model is not loaded");
    std::vector<std::pair<std::string, int>> pcs;
    impl_->EncodePieces(input, &pcs);
    for (auto& p : pcs) out->push_back(std::move(p.first));
    return util::Status();
}

util::Status SentencePieceProcessor::Decode(const std::vector<int>& ids, std::string* out) const {
    return Decode(ids.data(), (unsigned int)ids.size(), out);
}

// google/sentencepiece sentencepiece_processor.cc Decode(ids) -> Decode(pieces)
util::Status SentencePieceProcessor::Decode(const int* ids, unsigned int len, std::string* out) const {
    out->clear();
    const Impl& m = *impl_;
    if (!m.loaded) return util::Status("model is not loaded");
    if (m.byte_level) {
        for (unsigned int i = 0; i < len; ++i) {
            if (ids[i] >= 3 && ids[i] < 3 + 256) out->push_back((char)(ids[i] - 3));
            else if (ids[i] >= 3 + 256) *out += "<" + std::to_string(ids[i]) + ">"; // outside the byte range: placeholder
        }
        return util::Status();
    }
    for (unsigned int i = 0; i < len; ++i) {
        if (ids[i] < 0 || ids[i] >= (int)m.pieces.size()) return util::Status("Invalid id: " + std::to_string(ids[i]));
    }
    auto flush_bytes = [&](unsigned int begin, unsigned int end) {
        if (begin >= end) return;
        std::string bytes;
        for (unsigned int i = begin; i < end; ++i) bytes.push_back((char)PieceToByte(m.pieces[ids[i]].piece));
        size_t pos = 0;
        while (pos < bytes.size()) {
            const int n = ValidUtf8Len((const unsigned char*)bytes.data() + pos, bytes.size() - pos);
            if (n == 0) {
                out->append(kReplacement, 3);
                ++pos;
            } else {
                out->append(bytes, pos, n);
                pos += n;
            }
        }
    };
    unsigned int byte_start = 0;
    bool is_bos_ws = true, bos_ws_seen = false;
    for (unsigned int i = 0; i < len; ++i) {
        const int id = ids[i];
        if (m.Type(id) == BYTE) continue;
        flush_bytes(byte_start, i);
        if (bos_ws_seen || !out->empty()) is_bos_ws = false;
        byte_start = i + 1;
        bos_ws_seen = false;
        if (m.Type(id) == CONTROL) continue;
        if (m.Type(id) == UNKNOWN) {
            out->append(m.unk_surface);
            continue;
        }
        std::string_view piece = m.pieces[id].piece;
        if (is_bos_ws && (m.add_dummy_prefix || m.remove_extra_ws)) {
            bos_ws_seen = ConsumePrefix(&piece, kSpaceSymbol);
            // with extra-whitespace removal EVERY leading whitespace piece is dropped: the position stays "bos" until
            // something visible has been produced (observed upstream: ["▁", "▁", "a"] -> "a")
            if (m.remove_extra_ws) bos_ws_seen = false;
        }
        size_t pos = 0;
        while (pos < piece.size()) { // U+2581 -> ' '
            if (piece.compare(pos, 3, kSpaceSymbol) == 0) {
                out->push_back(' ');
                pos += 3;
            } else {
                out->push_back(piece[pos++]);
            }
        }
    }
    flush_bytes(byte_start, len);
    return util::Status();
}

int SentencePieceProcessor::GetPieceSize() const {
    return impl_->Size();
}
int SentencePieceProcessor::PieceToId(absl::string_view piece) const {
    if (impl_->byte_level) return piece.size() == 1 ? 3 + (unsigned char)piece[0] : 0;
    return impl_->PieceToId(piece);
}
const std::string& SentencePieceProcessor::IdToPiece(int id) const {
    static const std::string kByteLevel[3] = {"<unk>", "<s>", "</s>"};
    static const std::vector<std::string> kBytes = [] {
        std::vector<std::string> v;
        for (int b = 0; b < 256; ++b) v.emplace_back(1, (char)b);
        return v;
    }();
    if (impl_->byte_level) return id >= 3 && id < 3 + 256 ? kBytes[id - 3] : kByteLevel[id == 1 ? 1 : (id == 2 ? 2 : 0)];
    return id >= 0 && id < (int)impl_->pieces.size() ? impl_->pieces[id].piece : impl_->empty;
}
float SentencePieceProcessor::GetScore(int id) const {
    return id >= 0 && id < (int)impl_->pieces.size() ? impl_->pieces[id].score : 0.f;
}
bool SentencePieceProcessor::IsUnknown(int id) const {
    return impl_->byte_level ? id == 0 : impl_->Type(id) == UNKNOWN;
}
bool SentencePieceProcessor::IsControl(int id) const {
    return impl_->byte_level ? (id == 1 || id == 2) : impl_->Type(id) == CONTROL;
}
bool SentencePieceProcessor::IsUnused(int id) const {
    return impl_->Type(id) == UNUSED;
}
bool SentencePieceProcessor::IsByte(int id) const {
    return impl_->byte_level ? (id >= 3 && id < 3 + 256) : impl_->Type(id) == BYTE;
}
int SentencePieceProcessor::unk_id() const {
    if (impl_->byte_level) return 0;
    const int id = impl_->PieceToId(impl_->unk_piece);
    return IsUnknown(id) ? id : -1;
}
int SentencePieceProcessor::bos_id() const {
    if (impl_->byte_level) return 1;
    const int id = impl_->PieceToId(impl_->bos_piece);
    return IsControl(id) ? id : -1;
}
int SentencePieceProcessor::eos_id() const {
    if (impl_->byte_level) return 2;
    const int id = impl_->PieceToId(impl_->eos_piece);
    return IsControl(id) ? id : -1;
}
int SentencePieceProcessor::pad_id() const {
    if (impl_->byte_level) return -1;
    const int id = impl_->PieceToId(impl_->pad_piece);
    return IsControl(id) ? id : -1;
}

} // namespace sentencepiece

// C entry points for the CPU parity test (ctypes): encode -> ids, decode -> text
extern "C" {
__attribute__((visibility("default"))) void* b2pplnn_sp_load(const char* path, char* err, uint64_t cap) {
    auto* sp = new sentencepiece::SentencePieceProcessor();
    const auto st = sp->Load(path);
    if (!st.ok()) {
        if (err && cap) snprintf(err, cap, "%s", st.ToString().c_str());
        delete sp;
        return nullptr;
    }
    return sp;
}
__attribute__((visibility("default"))) void b2pplnn_sp_free(void* h) {
    delete (sentencepiece::SentencePieceProcessor*)h;
}
__attribute__((visibility("default"))) int32_t b2pplnn_sp_encode(void* h, const char* text, uint64_t len, int32_t* ids, int32_t cap) {
    std::vector<int> v;
    ((sentencepiece::SentencePieceProcessor*)h)->Encode(std::string_view(text, len), &v);
    for (size_t i = 0; i < v.size() && (int32_t)i < cap; ++i) ids[i] = v[i];
    return (int32_t)v.size();
}
__attribute__((visibility("default"))) int32_t b2pplnn_sp_decode(void* h, const int32_t* ids, int32_t n, char* out, uint64_t cap) {
    std::string s;
    const auto st = ((sentencepiece::SentencePieceProcessor*)h)->Decode(ids, (unsigned int)n, &s);
    if (!st.ok()) return -1;
    const uint64_t k = std::min<uint64_t>(s.size(), cap ? cap - 1 : 0);
    memcpy(out, s.data(), k);
    if (cap) out[k] = 0;
    return (int32_t)s.size();
}
__attribute__((visibility("default"))) int32_t b2pplnn_sp_info(void* h, int32_t what) {
    auto* sp = (sentencepiece::SentencePieceProcessor*)h;
    switch (what) {
        case 0: return sp->GetPieceSize();
        case 1: return sp->unk_id();
        case 2: return sp->bos_id();
        case 3: return sp->eos_id();
        case 4: return sp->pad_id();
        default: return -1;
    }
}
}
