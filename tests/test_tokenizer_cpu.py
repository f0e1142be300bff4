"""Tokenizer front end: host/src/sentencepiece.cc -- the SentencePiece inference implementation behind
the reference's TokenizerImplSP (src/tokenizer/tokenizer_impl_sp.h:22-68) -- against the OFFICIAL `sentencepiece` Python
package, on models trained here (seconds): LLaMA-style BPE with byte fallback and identity normalisation, BPE without
byte fallback, unigram with extra-whitespace removal, user-defined symbols.  Bar: identical ids for every input,
identical decoded text; models the implementation does not interpret (compiled character maps) are refused at Load.
Also: the reference's own tokenizer classes (TokenizerFactory -> TokenizerImplSP -> LlamaTokenizer), compiled in place
against this implementation, run on such a model (needs oracle/_ref, i.e. /root/reference at build time)."""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

spm = pytest.importorskip("sentencepiece")

ROOT = Path(__file__).resolve().parent.parent
HOSTLIB = ROOT / "ppl.llm.serving_b200" / "lib" / "libpplnn_b200.so"

WORDS = ("the quick brown fox jumps over lazy dog hello world large language model serving paged attention cache token "
         "decode prefill batch running step engine kernel tensor memory bandwidth roofline quantised weight scale "
         "Beijing Shanghai 北京 上海 東京 こんにちは 世界 naïve café résumé über straße Ελληνικά русский язык 12345 67890 "
         "a b c d e f g x y z A B C I You We They can't won't it's 3.14 2024-10-17 foo_bar baz.qux <tag> [x] {y} (z)").split()

TEXTS = [
    "", " ", "   ", "hello", "hello world", " hello world ", "hello   world", "  leading and trailing   ",
    "The quick brown fox jumps over the lazy dog.", "I believe the meaning of life is", "Simply put, the theory of relativity states that",
    "Building a website can be done in 10 simple steps:\n", "tab\tseparated\nnewline\r\n", "naïve café résumé",
    "北京是中国的首都", "こんにちは世界", "mixed 北京 and English words", "emoji 😀 and rare ☃ snow", "zzzzqqqq xxyyzz",
    "UPPER lower MiXeD 12345", "a" * 200, "hello" + " " * 20 + "world", "▁already escaped▁", "<s>literal control text</s>",
    "<0x41> literal byte piece text", "unknown \U0001F9EA\U0001F9EA\U0001F9EA run", "end with space ", "\n\n\n", "x",
]


def _corpus(tmp, n=3000, seed=0):
    rng = np.random.default_rng(seed)
    lines = [" ".join(rng.choice(WORDS, rng.integers(3, 14))) for _ in range(n)]
    p = tmp / "corpus.txt"
    p.write_text("\n".join(lines) + "\n", encoding="utf-8")
    return p


def _train(tmp, name, **kw):
    args = dict(input=str(_corpus(tmp)), model_prefix=str(tmp / name), vocab_size=kw.pop("vocab_size", 400),
                character_coverage=kw.pop("character_coverage", 0.995), num_threads=1, minloglevel=2,
                hard_vocab_limit=False)
    args.update(kw)
    spm.SentencePieceTrainer.train(**args)
    return tmp / f"{name}.model"


@pytest.fixture(scope="module")
def hostlib():
    assert HOSTLIB.exists(), "build with `python __graft_entry__.py build`"
    lib = C.CDLL(str(HOSTLIB))
    lib.b2pplnn_sp_load.restype = C.c_void_p
    lib.b2pplnn_sp_load.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64]
    lib.b2pplnn_sp_free.argtypes = [C.c_void_p]
    lib.b2pplnn_sp_encode.restype = C.c_int32
    lib.b2pplnn_sp_encode.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.POINTER(C.c_int32), C.c_int32]
    lib.b2pplnn_sp_decode.restype = C.c_int32
    lib.b2pplnn_sp_decode.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.c_char_p, C.c_uint64]
    lib.b2pplnn_sp_info.restype = C.c_int32
    lib.b2pplnn_sp_info.argtypes = [C.c_void_p, C.c_int32]
    return lib


class Mine:
    def __init__(self, lib, path):
        self.lib = lib
        err = C.create_string_buffer(1024)
        self.h = lib.b2pplnn_sp_load(str(path).encode(), err, len(err))
        self.err = err.value.decode()

    def encode(self, text: str):
        b = text.encode("utf-8")
        ids = (C.c_int32 * (4 * len(b) + 16))()
        n = self.lib.b2pplnn_sp_encode(self.h, b, len(b), ids, len(ids))
        return list(ids[:n])

    def encode_bytes(self, b: bytes):
        ids = (C.c_int32 * (4 * len(b) + 16))()
        n = self.lib.b2pplnn_sp_encode(self.h, b, len(b), ids, len(ids))
        return list(ids[:n])

    def decode(self, ids):
        arr = (C.c_int32 * max(1, len(ids)))(*ids)
        out = C.create_string_buffer(16 * len(ids) + 64)
        n = self.lib.b2pplnn_sp_decode(self.h, arr, len(ids), out, len(out))
        return None if n < 0 else out.raw[:n].decode("utf-8")

    def info(self):
        return [self.lib.b2pplnn_sp_info(self.h, i) for i in range(5)]

    def close(self):
        if self.h:
            self.lib.b2pplnn_sp_free(self.h)
            self.h = None


def _check_model(hostlib, path, extra_texts=()):
    ref = spm.SentencePieceProcessor(model_file=str(path))
    mine = Mine(hostlib, path)
    assert mine.h, mine.err
    try:
        assert mine.info() == [ref.get_piece_size(), ref.unk_id(), ref.bos_id(), ref.eos_id(), ref.pad_id()]
        rng = np.random.default_rng(1)
        texts = list(TEXTS) + list(extra_texts)
        texts += [" ".join(rng.choice(WORDS, rng.integers(1, 30))) for _ in range(200)]
        # random strings over a small alphabet incl. spaces: exercises merge order and tie-breaking
        alpha = list("abcdeht lowr ") + ["北", "京", "é", "  "]
        texts += ["".join(rng.choice(alpha, rng.integers(1, 40))) for _ in range(300)]
        for t in texts:
            want = ref.encode(t)
            got = mine.encode(t)
            assert got == want, f"encode({t!r}): got {got[:20]} want {want[:20]}"
            assert mine.decode(want) == ref.decode(want), f"decode of encode({t!r})"
        # decoding arbitrary id sequences (what a random-init model emits), incl. control / unknown / byte pieces
        n = ref.get_piece_size()
        for _ in range(200):
            ids = [int(x) for x in rng.integers(0, n, rng.integers(1, 24))]
            assert mine.decode(ids) == ref.decode(ids), ids
        for i in range(n):  # one id at a time: the reference decodes token by token (tokenizer_impl_sp.h:52-59)
            assert mine.decode([i]) == ref.decode([i]), i
        assert mine.decode([n]) is None and mine.decode([-1]) is None  # out of range: error status, as upstream
    finally:
        mine.close()


def test_llama_style_bpe_byte_fallback(hostlib, tmp_path):
    """the settings of LLaMA's tokenizer.model: BPE, byte_fallback, identity normaliser, dummy prefix, no extra-whitespace
    removal, digits split"""
    path = _train(tmp_path, "llama_like", model_type="bpe", byte_fallback=True, normalization_rule_name="identity",
                  remove_extra_whitespaces=False, add_dummy_prefix=True, split_digits=True, allow_whitespace_only_pieces=True,
                  vocab_size=600)
    _check_model(hostlib, path)
    # invalid UTF-8 in the prompt: each bad byte normalises to U+FFFD (then byte fallback), as upstream
    ref = spm.SentencePieceProcessor(model_file=str(path))
    mine = Mine(hostlib, path)
    for raw in [b"\xff", b"ok \xc3", b"\xe2\x96", b"a\x80b", b"\xf0\x9f\x98", b"\xc0\xaf", b"\xed\xa0\x80"]:
        assert mine.encode_bytes(raw) == ref.encode(raw), raw  # the official package takes raw bytes too
    mine.close()


def test_bpe_without_byte_fallback_merges_unknown_runs(hostlib, tmp_path):
    path = _train(tmp_path, "bpe_plain", model_type="bpe", byte_fallback=False, normalization_rule_name="identity",
                  character_coverage=0.98)
    _check_model(hostlib, path)


def test_unigram_identity_with_whitespace_removal(hostlib, tmp_path):
    path = _train(tmp_path, "uni", model_type="unigram", normalization_rule_name="identity", remove_extra_whitespaces=True)
    _check_model(hostlib, path)


def test_unigram_byte_fallback_and_user_defined_symbols(hostlib, tmp_path):
    path = _train(tmp_path, "uni_bf", model_type="unigram", byte_fallback=True, normalization_rule_name="identity",
                  user_defined_symbols=["<tag>", "[x]", "foo_bar"], remove_extra_whitespaces=False, vocab_size=700)
    _check_model(hostlib, path, extra_texts=["a <tag> b", "<tag><tag>[x]", "foo_barfoo_bar foo_ba", "x<tag"])


def test_bpe_user_defined_symbols_are_frozen(hostlib, tmp_path):
    path = _train(tmp_path, "bpe_ud", model_type="bpe", byte_fallback=True, normalization_rule_name="identity",
                  user_defined_symbols=["<tag>", "hello"], vocab_size=600)
    _check_model(hostlib, path, extra_texts=["hellohello world", "a<tag>b", "hell o hello"])


def test_char_model(hostlib, tmp_path):
    path = _train(tmp_path, "chars", model_type="char", normalization_rule_name="identity", vocab_size=200)
    _check_model(hostlib, path)


def test_models_with_a_compiled_character_map_are_refused(hostlib, tmp_path):
    path = _train(tmp_path, "nfkc", model_type="bpe")  # default normaliser: nmt_nfkc
    mine = Mine(hostlib, path)
    assert not mine.h and "character map" in mine.err
    junk = tmp_path / "junk.model"
    junk.write_bytes(b"not a model at all" * 5)
    mine = Mine(hostlib, junk)
    assert not mine.h and "not a sentencepiece model" in mine.err
    mine = Mine(hostlib, tmp_path / "missing.model")
    assert not mine.h and "cannot open" in mine.err


def test_byte_level_test_vocabulary(hostlib, tmp_path):
    p = tmp_path / "tokenizer.model"
    p.write_text("b2llm-byte-level-tokenizer\n")
    mine = Mine(hostlib, p)
    assert mine.h and mine.info() == [259, 0, 1, 2, -1]
    assert mine.encode("hé") == [3 + b for b in "hé".encode()]
    assert mine.decode([3 + b for b in "héllo".encode()]) == "héllo"
    assert mine.decode([1, 2, 31999, 3 + 65]) == "<31999>A"
    mine.close()


def test_reference_tokenizer_classes_over_this_implementation(tmp_path):
    """TokenizerFactory::Create("llama", "sentencepiece", ...) -> LlamaTokenizer(TokenizerImplSP): the reference's own
    headers compiled in place (oracle/_ref/tokenizer_check); Encode prepends BOS (llama_tokenizer.h:36-39), the
    token-by-token Decode restores the leading space of "▁word" pieces (tokenizer_impl_sp.h:52-59)"""
    exe = ROOT / "oracle" / "_ref" / "tokenizer_check"
    if not exe.exists():
        pytest.skip("oracle/_ref not built (no /root/reference at build time)")
    path = _train(tmp_path, "llama_like", model_type="bpe", byte_fallback=True, normalization_rule_name="identity",
                  remove_extra_whitespaces=False, vocab_size=600)
    ref = spm.SentencePieceProcessor(model_file=str(path))
    text = "hello world 北京 café"
    r = subprocess.run([str(exe), str(path), text], capture_output=True, text=True, timeout=60,
                       env=dict(os.environ, PPL_LOG_LEVEL="ERROR"))
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    ids = [int(x) for x in lines[0].split()[1:]]
    assert ids == [ref.bos_id()] + ref.encode(text)
    assert lines[1] == "bos %d eos %d" % (ref.bos_id(), ref.eos_id())
    assert lines[2] == "text " + ref.decode(ids)
    # streaming decode, one token at a time, re-assembles the prompt (the leading space comes from the dummy prefix)
    assert lines[3] == "stream  " + text
