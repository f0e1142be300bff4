"""The oracle against its committed golden vectors and against itself (CPU only).

The reference has no golden vectors for this path: tests/golden/*.npz were produced by
scripts/make_golden.py from the oracle and FREEZE it -- any change of a numeric convention shows here.
Beyond the fixtures, size-independent properties of the algorithm are checked: incremental decode
equals re-running the prefix, KV layouts are permutations of one another, the integer GEMM has two
independent implementations, the sampler's invariants.
"""
from pathlib import Path

import numpy as np
import pytest

from oracle import _native, llama_ref as ref, sampler_ref
from oracle.weights import ModelDesc, SynthWeights, quantize_weight_per_channel, synth_tensor

GOLD = Path(__file__).resolve().parent / "golden"


def _desc_from(arr):
    h, inter, L, nh, nkv, V, layout, mode, ps, qm, mp, kvbit, kvgroup = (int(x) for x in arr)
    return ModelDesc(h, inter, L, nh, nkv, V, cache_layout=layout, cache_mode=mode, page_size=ps, quant_method=qm,
                     max_position=mp, cache_quant_bit=kvbit, cache_quant_group=kvgroup)


@pytest.mark.parametrize("name", ["step_w8a8_paged_l3", "step_fp16_contig_l1_gqa", "step_w8a8_fp16kv_paged_l2_gqa"])
def test_step_golden(name):
    g = np.load(GOLD / f"{name}.npz")
    desc = _desc_from(g["desc"])
    orc = ref.LlamaOracle(desc, SynthWeights(desc, 0xB200), 256)
    for it in range(3):
        kw = {}
        if desc.cache_mode == 1:
            kw = dict(page_list=g["page_list"], max_pages=int(g["max_pages"]))
        else:
            kw = dict(cache_indices=g["cache_indices"])
        step = ref.Step(g[f"s{it}_token_inputs"], g[f"s{it}_seq_starts"], g[f"s{it}_kv_starts"], g[f"s{it}_start_pos"],
                        int(g[f"s{it}_decoding_batches"]), **kw)
        logits = orc.forward(step)
        # BLAS summation order may differ between hosts: fp32 tolerance, not bit equality, on logits ...
        np.testing.assert_allclose(logits, g[f"s{it}_logits"], rtol=0, atol=2e-4 * np.abs(g[f"s{it}_logits"]).max())
        tok, lp = sampler_ref.sample_topk_topp(logits, None, None, None, desc.vocab_size, 1, 0.0)
        # ... but the greedy tokens are pinned exactly
        assert tok.tolist() == g[f"s{it}_tokens"].tolist()
        np.testing.assert_allclose(lp, g[f"s{it}_logprobs"], atol=1e-4)
    cache, scale = orc.cache.export()
    if desc.quant_method == 1 and desc.cache_quant_bit == 8:  # integer path end to end: the int8 cache is bit-exact
        mism = (cache != g["kv_cache_final"]).mean()
        assert mism < 1e-3, f"{mism:.2e} of the int8 KV codes differ"
    assert cache.shape == g["kv_cache_final"].shape and scale.shape == g["kv_scale_final"].shape


def test_ops_golden():
    g = np.load(GOLD / "ops.npz")
    assert np.array_equal(synth_tensor(0xB200, 5, (3, 64), 0.02), g["synth_t5"])
    q, s = ref.quant_rows(ref.rmsnorm_f32(g["rms_x"], g["rms_g"], 1e-5))
    assert np.array_equal(q, g["rms_q"]) and np.array_equal(s, g["rms_s"])
    assert np.array_equal(ref.gemm_i8_acc_numpy(g["gemm_a"], g["gemm_w"]), g["gemm_acc"])
    kq, ks = ref.kv_quant(g["kv_x"], 8)
    assert np.array_equal(kq, g["kv_q"]) and np.array_equal(ks, g["kv_s"])
    cos, sin = ref.rope_table(32, 128, 10000.0)
    assert np.array_equal(cos[31], g["rope_cos_31"]) and np.array_equal(sin[31], g["rope_sin_31"])
    assert np.array_equal(ref.apply_rope(g["kv_x"], np.array([0, 3, 17, 31]), cos, sin), g["rope_out"])
    tok, lp = sampler_ref.sample_topk_topp(g["samp_logits"], np.array([0.7, 1.0, 1.3, 0.5], np.float32),
                                           np.array([0.9, 0.5, 1.0, 0.0], np.float32), g["samp_rand"], 300, 8, 0.0)
    assert tok.tolist() == g["samp_tok"].tolist()
    np.testing.assert_allclose(lp, g["samp_lp"], atol=1e-5)


def test_int8_gemm_two_implementations_agree():
    """numpy (float64 BLAS, exact below 2^53) vs the C restatement (int32 accumulate)"""
    if _native.lib() is None:
        pytest.skip("no C compiler for oracle/csrc")
    rng = np.random.default_rng(0)
    for M, N, K in [(1, 8, 16), (7, 33, 129), (16, 64, 11008)]:
        a = rng.integers(-127, 128, (M, K), dtype=np.int8)
        w = rng.integers(-127, 128, (N, K), dtype=np.int8)
        assert np.array_equal(ref.gemm_i8_acc_numpy(a, w), _native.gemm_i8_i32(a, w))
    a = np.full((2, 11008), 127, np.int8)                # worst case magnitude: 127*127*11008 < 2^31
    assert int(_native.gemm_i8_i32(a, a)[0, 0]) == 127 * 127 * 11008


def test_quant_conventions():
    y = np.array([[0.0, 0.0, 0.0, 0.0], [1.0, -2.0, 0.5, 2.0], [127.0, 63.5, -126.5, 1e-3]], np.float32)
    q, s = ref.quant_rows(y)
    assert q[0].tolist() == [0, 0, 0, 0] and s[0] == 0          # all-zero row: scale 0, codes 0
    assert q[1].tolist() == [64, -127, 32, 127]                 # 63.5 -> 64 (half to even), 31.75 -> 32
    assert q[2].tolist() == [127, 64, -126, 0]                  # 63.5 -> 64 (even), -126.5 -> -126 (even)
    assert np.abs(q).max() <= 127                               # -128 never produced
    w = np.array([[0.5, -1.0], [0.0, 0.0]], np.float16)
    wq, ws = quantize_weight_per_channel(w)
    assert wq.tolist() == [[64, -127], [0, 0]] and ws[1] == 0
    x = np.zeros((1, 1, 16), np.float16)
    x[0, 0, :8] = [1, -1, 0.5, 0.25, 0, 0, 0, 0.004]
    kq, ks = ref.kv_quant(x, 8)
    assert ks.dtype == np.float16 and ks[0, 0, 1] == 0 and (kq[0, 0, 8:] == 0).all()
    deq = ref.kv_dequant(kq, ks, 8)
    assert np.abs(deq[0, 0, :8] - x[0, 0, :8].astype(np.float32)).max() <= float(ks[0, 0, 0]) * 0.5 + 1e-3


@pytest.mark.parametrize("layout", [0, 1, 2, 3])
def test_kv_layouts_are_permutations(layout):
    """llm_engine.cc:118-169: the four layouts hold the same elements; export -> load is the identity and the
    exported shape is the reference's."""
    desc = ModelDesc(64, 128, 3, 2, 2, 64, cache_layout=layout, cache_mode=0)
    c = ref.KVCache(desc, 5)
    rng = np.random.default_rng(layout)
    c.cache[:] = rng.integers(-127, 128, c.cache.shape, dtype=np.int8)
    c.scale[:] = rng.standard_normal(c.scale.shape).astype(np.float16)
    L, T, H, D, G = 3, 5, 2, 32, 4
    want = {0: (T, L, 2, H, D), 1: (L, T, 2, H, D), 2: (L, 2, T, H, D), 3: (L, 2, H, T, D)}[layout]
    ec, es = c.export()
    assert ec.shape == want and es.shape == want[:-1] + (G,)
    c2 = ref.KVCache(desc, 5)
    c2.load(ec.reshape(-1), es.reshape(-1))
    assert np.array_equal(c2.cache, c.cache) and np.array_equal(c2.scale, c.scale)
    # element (l=1, kv=1, t=3, h=1, d=7) sits where the layout formula says
    flat = ec.reshape(-1)
    strides = {0: lambda l, kv, t, h, d: (((t * L + l) * 2 + kv) * H + h) * D + d,
               1: lambda l, kv, t, h, d: (((l * T + t) * 2 + kv) * H + h) * D + d,
               2: lambda l, kv, t, h, d: (((l * 2 + kv) * T + t) * H + h) * D + d,
               3: lambda l, kv, t, h, d: (((l * 2 + kv) * H + h) * T + t) * D + d}[layout]
    assert flat[strides(1, 1, 3, 1, 7)] == c.cache[1, 1, 3, 1, 7]


def test_incremental_decode_matches_cached_prefill():
    """property: decoding token t with the cache == prefilling [0..t] when the prefix is read from the cache
    (cache_prefill), because both read the same quantised K/V for the prefix."""
    desc = ModelDesc(128, 256, 2, 2, 2, 128, cache_layout=2, cache_mode=0, quant_method=0, max_position=64)
    w = SynthWeights(desc, 1)
    rng = np.random.default_rng(5)
    toks = list(map(int, rng.integers(0, 128, 9)))
    a = ref.LlamaOracle(desc, w, 64)
    a.forward(ref.build_step(desc, [toks[:8]], [0], 0, cache_indices=[0]))
    la = a.forward(ref.build_step(desc, [toks[8:]], [8], 1, cache_indices=[0]))
    b = ref.LlamaOracle(desc, w, 64)
    b.forward(ref.build_step(desc, [toks[:8]], [0], 0, cache_indices=[0]))
    lb = b.forward(ref.build_step(desc, [toks[8:]], [8], 0, cache_indices=[0]))   # as a 1-token prefill with prefix
    # the decode path reads the current token's K/V quantised, the prefill path fresh: small difference only
    assert np.abs(la - lb).max() <= 2e-2 * np.abs(la).max()
    assert la.argmax() == lb.argmax()


def test_batch_independence_and_order():
    """rows of a ragged batch do not interact: permuting sequences permutes logits rows"""
    desc = ModelDesc(128, 256, 1, 2, 1, 128, cache_layout=3, cache_mode=1, page_size=4, quant_method=1, max_position=64)
    w = SynthWeights(desc, 2)
    rng = np.random.default_rng(9)
    p = [list(map(int, rng.integers(0, 128, n))) for n in (3, 6, 1)]
    pg = [[0, 4], [8, 12], [16]]
    o1 = ref.LlamaOracle(desc, w, 32)
    l1 = o1.forward(ref.build_step(desc, p, [0, 0, 0], 0, page_tables=pg))
    o2 = ref.LlamaOracle(desc, w, 32)
    l2 = o2.forward(ref.build_step(desc, p[::-1], [0, 0, 0], 0, page_tables=pg[::-1]))
    np.testing.assert_allclose(l1, l2[::-1], atol=1e-5 * np.abs(l1).max())


def test_sampler_properties():
    rng = np.random.default_rng(1)
    logits = rng.standard_normal((16, 1000)).astype(np.float32) * 4
    tok, lp = sampler_ref.sample_topk_topp(logits, None, None, None, 1000, 1, 0.0)
    assert tok.tolist() == logits.argmax(axis=1).tolist()
    lse = np.log(np.exp(logits.astype(np.float64)).sum(axis=1))
    np.testing.assert_allclose(lp, logits[np.arange(16), tok] - lse, atol=1e-4)
    # top_p <= 0 keeps one candidate whatever the random draw; rand = 0 picks the arg-max at any k
    r = rng.random(16).astype(np.float32)
    t2, _ = sampler_ref.sample_topk_topp(logits, None, np.zeros(16, np.float32), r, 1000, 50, 0.0)
    assert t2.tolist() == tok.tolist()
    t3, _ = sampler_ref.sample_topk_topp(logits, None, np.ones(16, np.float32), np.zeros(16, np.float32), 1000, 50, 1.0)
    assert t3.tolist() == tok.tolist()
    # every draw lies inside the top-k set; ties resolve to the lower index
    t4, _ = sampler_ref.sample_topk_topp(logits, np.full(16, 2.0, np.float32), np.ones(16, np.float32), r, 1000, 8, 1.0)
    topk = np.argsort(-logits, axis=1, kind="stable")[:, :8]
    assert all(t4[i] in topk[i] for i in range(16))
    tie = np.zeros((1, 10), np.float32)
    assert sampler_ref.sample_topk_topp(tie, None, None, None, 10, 1, 0.0)[0][0] == 0
    # padded row stride: entries beyond vocab are never selected
    padded = np.concatenate([logits, np.full((16, 24), 1e9, np.float32)], axis=1)
    assert sampler_ref.sample_topk_topp(padded, None, None, None, 1000, 1, 0.0)[0].tolist() == tok.tolist()


def test_penalty_reference_semantics():
    V = 12
    logits = np.arange(-6, 6, dtype=np.float32)[None, :].repeat(2, axis=0).copy()
    cm = np.zeros((4, V), np.uint16)
    cm[2, 5] = 9                                            # stale counts of a previous tenant of slot 2
    out = sampler_ref.apply_penalty(logits, [2.0, 1.0], [2.0, 1.5], [0.5, 0.0], [0.1, 0.0], [2, 0], [3, 3, 10, 7],
                                    [0, 3, 4], [0, 5], V, cm)
    assert cm[2, 5] == 0 and cm[2, 3] == 2 and cm[2, 10] == 1   # start_pos 0 clears the row first
    assert cm[0, 7] == 1                                         # start_pos > 0 keeps history
    # token 3: logit -3 -> *2 -> -6, -0.5 presence, -0.1*2 frequency -> -6.7, /T=2 -> -3.35
    assert out[0, 3] == pytest.approx(-3.35, abs=1e-6)
    # token 10: logit 4 -> /2 -> 2, -0.5, -0.1 -> 1.4 -> 0.7
    assert out[0, 10] == pytest.approx(0.7, abs=1e-6)
    assert out[0, 0] == pytest.approx(-3.0)                      # unseen: only the temperature
    assert out[1, 7] == pytest.approx(1.0 / 1.5, abs=1e-6)


def test_fp16_cache_decode_equals_prefill_of_the_longer_prompt():
    """cache_quant_bit 0 / group 1 (llm_generator.cc:131-136): the cache holds the rotated K and V as fp16, exactly the
    values prefill attention uses for fresh tokens -- so decoding token n after a prefill of n tokens must give the
    logits of a prefill over n + 1 tokens (a property the int8 cache does not have: it re-reads quantised values)."""
    desc = ModelDesc(256, 512, 2, 4, 2, 512, cache_layout=2, cache_mode=1, page_size=8, quant_method=0, max_position=64,
                     cache_quant_bit=0, cache_quant_group=1)
    assert desc.kv_bytes_per_token() == (2 * 2 * 2 * 64 * 2, 0)
    w = SynthWeights(desc, 7)
    rng = np.random.default_rng(3)
    toks = list(map(int, rng.integers(0, desc.vocab_size, 12)))
    pages = [[16, 0, 8]]
    a = ref.LlamaOracle(desc, w, 32)
    a.forward(ref.build_step(desc, [toks[:11]], [0], 0, page_tables=pages))
    dec = a.forward(ref.build_step(desc, [toks[11:]], [11], 1, page_tables=pages))
    b = ref.LlamaOracle(desc, w, 32)
    full = b.forward(ref.build_step(desc, [toks], [0], 0, page_tables=pages))
    assert a.cache.cache.dtype == np.float16 and a.cache.scale.size == 0
    assert np.array_equal(a.cache.cache, b.cache.cache)
    np.testing.assert_allclose(dec, full, rtol=0, atol=2e-3 * np.abs(full).max())
    # and the layouts round-trip through export / load
    c, s = a.cache.export()
    k = ref.KVCache(desc, 32)
    k.load(c, s)
    assert np.array_equal(k.cache, a.cache.cache)
