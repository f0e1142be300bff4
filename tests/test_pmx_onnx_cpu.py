"""Model-slice loader: ``RuntimeBuilder::LoadModel`` on a ppl.pmx ONNX export
(resource_manager.cc:117-147, 280-290; docs/llama_guide.md:12-36), CPU part.

The export is written by ppl_llm_serving_b200.pmx_onnx_writer with the official protobuf runtime (an encoder
independent of host/src/onnx_wire.h); ``b2pplnn_inspect_model`` (libpplnn_b200.so) runs the SAME code path the
runtime uses -- onnx_model.cc -> pmx_llama.cc -> PmxLlama::ForEachWeight -- with a hashing sink in place of
``b2llm_engine_load_weight_shard``.  Checked: model dimensions and graph constants, the tensor-parallel shard layouts,
every payload encoding a real file can have (raw / typed / external data; fp16 / fp32 / bf16; packed and unpacked
repeated fields), and that graphs the fixed b2llm forward does not implement are refused, not mis-executed.
The GPU part (tokens through the reference's generator from such an export) is tests/test_host_cpp.py.
"""
import ctypes as C
import json
from pathlib import Path

import numpy as np
import pytest

from oracle.weights import ModelDesc, SynthWeights
from ppl_llm_serving_b200 import pmx_onnx_writer as W

ROOT = Path(__file__).resolve().parent.parent
HOSTLIB = ROOT / "ppl.llm.serving_b200" / "lib" / "libpplnn_b200.so"

KIND = dict(EMBEDDING=0, FINAL_NORM=1, LM_HEAD=2, ATTN_NORM=3, QKV=4, O=5, FFN_NORM=6, GATE=7, UP=8, DOWN=9)  # b2llm.h


@pytest.fixture(scope="module")
def hostlib():
    assert HOSTLIB.exists(), "build with `python __graft_entry__.py build`"
    lib = C.CDLL(str(HOSTLIB))
    lib.b2pplnn_inspect_model.restype = C.c_int32
    lib.b2pplnn_inspect_model.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64]
    return lib


def inspect(lib, path):
    buf = C.create_string_buffer(1 << 20)
    rc = lib.b2pplnn_inspect_model(str(path).encode(), buf, len(buf))
    return rc, json.loads(buf.value.decode())


def fnv1a(a: np.ndarray) -> str:
    h = 1469598103934665603
    for b in np.ascontiguousarray(a).tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return format(h, "x")


def small_desc(**kw):
    return ModelDesc(64, 96, 2, 4, kw.pop("num_kv_heads", 4), kw.pop("vocab_size", 80), cache_layout=kw.pop("cache_layout", 3),
                     cache_mode=kw.pop("cache_mode", 1), page_size=kw.pop("page_size", 16), max_position=256, **kw)


def expected_weights(desc, w, rank, tp):
    """(kind, layer, fp16 array) in the order the runtime uploads them; embedding / lm head are whole on every rank"""
    s = W.shard_weights(desc, w, rank, tp, fused_qkv=True)
    out = [(KIND["EMBEDDING"], 0, np.asarray(w.embedding(), np.float16)), (KIND["FINAL_NORM"], 0, s["norm.weight"]),
           (KIND["LM_HEAD"], 0, np.asarray(w.lm_head(), np.float16))]
    for l in range(desc.num_layers):
        p = f"layers.{l}."
        out += [(KIND["ATTN_NORM"], l, s[p + "attention_norm.weight"]), (KIND["QKV"], l, s[p + "attention.wqkv.weight"]),
                (KIND["O"], l, s[p + "attention.wo.weight"]), (KIND["FFN_NORM"], l, s[p + "ffn_norm.weight"]),
                (KIND["GATE"], l, s[p + "feed_forward.w1.weight"]), (KIND["UP"], l, s[p + "feed_forward.w3.weight"]),
                (KIND["DOWN"], l, s[p + "feed_forward.w2.weight"])]
    return out


def check_weights(info, expected, convert=lambda a: a):
    got = info["weights"]
    assert [(g["kind"], g["layer"]) for g in got] == [(k, l) for k, l, _ in expected]
    for g, (k, l, a) in zip(got, expected):
        assert g["elements"] == a.size, (g, a.shape)
        assert g["fnv1a"] == fnv1a(convert(a)), f"payload of kind {k} layer {l} ({g['name']}) differs"


def check_desc(info, desc, tp=1, rank=0):
    for k in ("hidden_dim", "intermediate_dim", "num_layers", "num_heads", "num_kv_heads", "vocab_size", "cache_quant_bit",
              "cache_quant_group", "cache_layout", "cache_mode", "page_size"):
        assert info[k] == getattr(desc, k), k
    assert info["norm_eps"] == pytest.approx(desc.norm_eps, rel=1e-6)
    assert info["rope_theta"] == pytest.approx(desc.rope_theta, rel=1e-6)
    assert info["tensor_parallel_size"] == tp and info["rank"] == rank


@pytest.mark.parametrize("syntax", ["proto3", "proto2"])
def test_single_slice_export(hostlib, tmp_path, syntax):
    desc = small_desc(norm_eps=1e-6, rope_theta=500000.0)
    w = SynthWeights(desc, 7)
    W.write_pmx_export(tmp_path, desc, w, syntax=syntax)
    rc, info = inspect(hostlib, tmp_path / "model_slice_0" / "model.onnx")
    assert rc == 0, info
    check_desc(info, desc)
    assert info["fused_qkv"] is True and info["producer"] == "pytorch" and info["warnings"] == []
    assert info["initializers"] == 3 + 7 * desc.num_layers
    check_weights(info, expected_weights(desc, w, 0, 1))


@pytest.mark.parametrize("emb_split,head_split", [("hidden", "vocab"), ("vocab", "vocab"), ("whole", "whole")])
def test_tensor_parallel_slices_gqa(hostlib, tmp_path, emb_split, head_split):
    """two slices, 4 q heads over 2 kv heads: shard shapes give tp = 2; the embedding / lm head pieces of BOTH slices are
    assembled into the whole tensors every rank holds"""
    desc = small_desc(num_kv_heads=2, cache_layout=1, cache_mode=0)
    w = SynthWeights(desc, 11)
    W.write_pmx_export(tmp_path, desc, w, tensor_parallel_size=2, emb_split=emb_split, head_split=head_split)
    for r in range(2):
        rc, info = inspect(hostlib, tmp_path / f"model_slice_{r}" / "model.onnx")
        assert rc == 0, info
        check_desc(info, desc, tp=2, rank=r)
        check_weights(info, expected_weights(desc, w, r, 2))


def test_vocab_parallel_lm_head_is_passed_through(hostlib, tmp_path):
    """vocab 128 at tp = 2: vocab / tp = 64 is a multiple of 32, so the engine's lm head is vocab-parallel and the loader hands
    every rank ITS [vocab / tp, hidden] slice of output.weight as the export stores it (no re-assembly); the embedding is
    still assembled whole.  (vocab 80 in the test above: 40 rows per rank -> the head stays whole and is assembled.)"""
    desc = small_desc(num_kv_heads=2, vocab_size=128)
    w = SynthWeights(desc, 13)
    W.write_pmx_export(tmp_path, desc, w, tensor_parallel_size=2, emb_split="hidden", head_split="vocab")
    for r in range(2):
        rc, info = inspect(hostlib, tmp_path / f"model_slice_{r}" / "model.onnx")
        assert rc == 0, info
        exp = expected_weights(desc, w, r, 2)
        head = np.asarray(w.lm_head(), np.float16)[r * 64:(r + 1) * 64]
        exp[2] = (KIND["LM_HEAD"], 0, head)
        check_weights(info, exp)


def test_split_qkv_external_data(hostlib, tmp_path):
    """--fused_qkv 0 (wq / wk / wv concatenated by the loader) with every initializer in its own external data file"""
    desc = small_desc()
    w = SynthWeights(desc, 3)
    W.write_pmx_export(tmp_path, desc, w, fused_qkv=False, external_data=True)
    assert (tmp_path / "model_slice_0" / "layers.0.attention.wq.weight").exists()
    rc, info = inspect(hostlib, tmp_path / "model_slice_0" / "model.onnx")
    assert rc == 0, info
    assert info["fused_qkv"] is False and info["initializers"] == 3 + 9 * desc.num_layers
    check_desc(info, desc)
    check_weights(info, expected_weights(desc, w, 0, 1))


@pytest.mark.parametrize("dtype,typed", [("fp32", False), ("fp32", True), ("bf16", False), ("fp16", True)])
def test_payload_encodings(hostlib, tmp_path, dtype, typed):
    desc = small_desc()
    w = SynthWeights(desc, 5)
    W.write_pmx_export(tmp_path, desc, w, dtype=dtype, typed_data=typed)
    rc, info = inspect(hostlib, tmp_path / "model_slice_0" / "model.onnx")
    assert rc == 0, info

    def via_bf16(a):
        bits = W._to_bf16_bits(a.astype(np.float32)).astype(np.uint32) << 16
        return bits.view(np.float32).astype(np.float16)

    check_weights(info, expected_weights(desc, w, 0, 1), convert=via_bf16 if dtype == "bf16" else (lambda a: a))


@pytest.mark.filterwarnings("ignore:overflow")
def test_float_to_half_matches_numpy(hostlib, tmp_path):
    """the loader's fp32 -> fp16 (round to nearest even, subnormals, overflow to inf) against numpy on awkward values"""
    desc = small_desc()
    rng = np.random.default_rng(0)

    class Awkward(SynthWeights):
        def embedding(self):
            n = desc.vocab_size * desc.hidden_dim
            v = np.concatenate([rng.standard_normal(n // 4) * 1e-6, rng.standard_normal(n // 4) * 6e-5,
                                rng.standard_normal(n // 4) * 7e4, rng.standard_normal(n - 3 * (n // 4))]).astype(np.float32)
            v[:8] = [0.0, -0.0, 65504.0, 65519.9, 65520.0, 5.9604645e-8, 2.9802322e-8, 2.98023224e-8 * 1.0000001]
            halves = (np.arange(256, dtype=np.uint16) + 0x3C00).view(np.float16).astype(np.float32)  # ties: exactly between halves
            v[8:8 + 255] = (halves[:-1] + halves[1:]) / 2
            return v.reshape(desc.vocab_size, desc.hidden_dim)

    w = Awkward(desc, 1)
    emb32 = w.embedding()
    w.embedding = lambda: emb32
    P = W.onnx_messages("proto3")
    W.write_pmx_export(tmp_path, desc, w)
    # replace the embedding initializer by the fp32 tensor itself
    f = tmp_path / "model_slice_0" / "model.onnx"
    m = P["ModelProto"]()
    m.ParseFromString(f.read_bytes())
    for t in m.graph.initializer:
        if t.name == "tok_embeddings.weight":
            t.data_type = W.DT_FLOAT
            t.raw_data = emb32.tobytes()
    f.write_bytes(m.SerializeToString())
    rc, info = inspect(hostlib, f)
    assert rc == 0, info
    with np.errstate(over="ignore"):
        want = emb32.astype(np.float16)
    assert info["weights"][0]["fnv1a"] == fnv1a(want)


def test_constants_fall_back_to_params_json(hostlib, tmp_path):
    desc = small_desc(cache_layout=2, cache_mode=0)
    w = SynthWeights(desc, 9)
    W.write_pmx_export(tmp_path, desc, w, attrs=False)
    rc, info = inspect(hostlib, tmp_path / "model_slice_0" / "model.onnx")
    assert rc == 0, info
    check_desc(info, desc)  # eps / theta: LLaMA-2 defaults == ModelDesc defaults
    assert len(info["warnings"]) == 3 and any("params.json" in x for x in info["warnings"])


def _rewrite(path, edit):
    P = W.onnx_messages("proto3")
    m = P["ModelProto"]()
    m.ParseFromString(path.read_bytes())
    edit(m, P)
    path.write_bytes(m.SerializeToString())


def _set_attr(m, op, name, value):
    for n in m.graph.node:
        if n.op_type == op:
            for a in n.attribute:
                if a.name == name:
                    if isinstance(value, str):
                        a.s = value.encode()
                    else:
                        a.i = value


@pytest.mark.parametrize("case,needle", [
    ("alibi", "is_alibi"), ("bias_term", "bias_term"), ("bias_tensor", "bias"), ("rotary_dim", "partial rotary"),
    ("rope_scaling", "scaling_type"), ("params_mismatch", "params.json says cache_layout"), ("missing_layer_weight", "incomplete weight set"),
    ("truncated", "malformed"), ("short_payload", "payload bytes"), ("missing_external", "cannot open"), ("not_onnx", "not an ONNX"),
    ("layer_shape", "weight shapes differ"), ("external_escape", "bad external data location")])
def test_unsupported_or_broken_exports_are_refused(hostlib, tmp_path, case, needle):
    desc = small_desc()
    w = SynthWeights(desc, 2)
    W.write_pmx_export(tmp_path, desc, w, external_data=case in ("missing_external", "external_escape"))
    f = tmp_path / "model_slice_0" / "model.onnx"
    if case == "alibi":
        _rewrite(f, lambda m, P: _set_attr(m, "MultiHeadCacheAttention", "is_alibi", 1))
    elif case == "bias_term":
        _rewrite(f, lambda m, P: _set_attr(m, "RowParallelLinear", "bias_term", 1))
    elif case == "bias_tensor":
        def add_bias(m, P):
            t = m.graph.initializer.add(name="layers.0.attention.wo.bias", data_type=W.DT_FLOAT16)
            t.dims.append(desc.hidden_dim)
            t.raw_data = np.zeros(desc.hidden_dim, np.float16).tobytes()
        _rewrite(f, add_bias)
    elif case == "rotary_dim":
        _rewrite(f, lambda m, P: _set_attr(m, "RotaryPositionEmbedding", "rotary_dim", 8))
    elif case == "rope_scaling":
        _rewrite(f, lambda m, P: _set_attr(m, "RotaryPositionEmbedding", "scaling_type", "linear"))
    elif case == "params_mismatch":
        p = json.loads((tmp_path / "params.json").read_text())
        p["cache_layout"] = 0
        (tmp_path / "params.json").write_text(json.dumps(p))
    elif case == "missing_layer_weight":
        def drop(m, P):
            keep = [t for t in m.graph.initializer if t.name != "layers.1.feed_forward.w2.weight"]
            del m.graph.initializer[:]
            m.graph.initializer.extend(keep)
        _rewrite(f, drop)
    elif case == "truncated":
        b = f.read_bytes()
        f.write_bytes(b[: len(b) // 2])
    elif case == "short_payload":
        def shorten(m, P):
            t = m.graph.initializer[3]
            t.raw_data = t.raw_data[:-2]
        _rewrite(f, shorten)
    elif case == "missing_external":
        (tmp_path / "model_slice_0" / "layers.1.attention.wo.weight").unlink()
    elif case == "not_onnx":
        f.write_bytes(b"this is not a protobuf at all, certainly not a model\n" * 10)
    elif case == "layer_shape":
        def reshape(m, P):
            for t in m.graph.initializer:
                if t.name == "layers.1.feed_forward.w1.weight":
                    d0, d1 = t.dims
                    del t.dims[:]
                    t.dims.extend([d0 // 2, d1 * 2])
        _rewrite(f, reshape)
    elif case == "external_escape":
        def escape(m, P):
            for e in m.graph.initializer[0].external_data:
                if e.key == "location":
                    e.value = "../params.json"
        _rewrite(f, escape)
    rc, info = inspect(hostlib, f)
    assert rc != 0 and needle in info["error"], info


def test_b2llm_descriptor_files_still_take_the_descriptor_path(hostlib, tmp_path):
    """a b2llm model-slice descriptor is not an ONNX file: the inspector (ONNX only) must say so, LoadModel sniffs the magic"""
    from ppl_llm_serving_b200.model_slice import write_model_dir
    desc = small_desc()
    write_model_dir(tmp_path, desc)
    rc, info = inspect(hostlib, tmp_path / "model_slice_0" / "model.onnx")
    assert rc != 0 and "not an ONNX" in info["error"]


def test_runtime_builder_load_model_recognises_both_file_kinds(tmp_path):
    """ppl::nn::onnx::RuntimeBuilder::LoadModel itself (the call at resource_manager.cc:124-131), on the CPU: the ONNX
    export and the b2llm descriptor both load; garbage is RC_INVALID_VALUE; Preprocess without an engine fails cleanly"""
    import os
    import subprocess
    from ppl_llm_serving_b200.model_slice import write_model_dir
    exe = ROOT / "ppl.llm.serving_b200" / "host" / "build" / "test_load_model"
    assert exe.exists(), "build with `python __graft_entry__.py build`"
    desc = small_desc()
    W.write_pmx_export(tmp_path / "onnx", desc, SynthWeights(desc, 4))
    write_model_dir(tmp_path / "desc", desc)
    (tmp_path / "junk.onnx").write_bytes(b"\x00\x01\x02junk" * 9)
    env = dict(os.environ, PPL_LOG_LEVEL="ERROR")
    good = subprocess.run([str(exe), str(tmp_path / "onnx/model_slice_0/model.onnx"), str(tmp_path / "desc/model_slice_0/model.onnx")],
                          capture_output=True, text=True, timeout=60, env=env)
    assert good.returncode == 0 and good.stdout.count("success") == 2, good.stdout + good.stderr
    bad = subprocess.run([str(exe), str(tmp_path / "junk.onnx")], capture_output=True, text=True, timeout=60, env=env)
    assert bad.returncode == 1 and "not an ONNX" in bad.stderr, bad.stdout + bad.stderr


def test_parsers_survive_damaged_files(tmp_path):
    """the ONNX model-slice loader and the sentencepiece model loader under AddressSanitizer + UBSan with truncated /
    bit-flipped / length-bombed copies of valid files (host/tests/fuzz_loaders.cc): refusing is fine, crashing is not"""
    import os
    import subprocess
    host = ROOT / "ppl.llm.serving_b200" / "host"
    b = subprocess.run(["make", "-C", str(host), "build/fuzz_loaders"], capture_output=True, text=True, timeout=600)
    if b.returncode != 0:
        pytest.skip("sanitizer build not available: " + b.stderr[-300:])
    desc = small_desc()
    W.write_pmx_export(tmp_path / "inline", desc, SynthWeights(desc, 1))
    W.write_pmx_export(tmp_path / "ext", desc, SynthWeights(desc, 1), external_data=True, fused_qkv=False, syntax="proto2")
    seeds = [tmp_path / "inline/model_slice_0/model.onnx", tmp_path / "ext/model_slice_0/model.onnx"]
    try:
        import sentencepiece as spm
        words = "the quick brown fox jumps over lazy dog hello world 北京 café".split()
        rng = np.random.default_rng(0)
        (tmp_path / "c.txt").write_text("\n".join(" ".join(rng.choice(words, 8)) for _ in range(2000)))
        for mt in ("bpe", "unigram"):
            spm.SentencePieceTrainer.train(input=str(tmp_path / "c.txt"), model_prefix=str(tmp_path / mt), vocab_size=300,
                                           model_type=mt, byte_fallback=True, normalization_rule_name="identity",
                                           minloglevel=2, hard_vocab_limit=False, num_threads=1)
            seeds.append(tmp_path / f"{mt}.model")
    except ImportError:
        pass
    r = subprocess.run([str(host / "build" / "fuzz_loaders"), "1200", "7"] + [str(s) for s in seeds], capture_output=True,
                       text=True, timeout=600, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=1", PPL_LOG_LEVEL="ERROR"))
    assert r.returncode == 0 and "fuzz ok" in r.stdout, (r.stdout + r.stderr)[-3000:]
