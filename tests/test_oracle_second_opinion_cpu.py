"""Independent second opinions on the oracle's arithmetic, from library code that is already in the image.

The oracle is builder-written and PARITY UNPINNED (the reference holds no kernels and no golden vectors).
These checks do not pin it to the reference -- nothing can -- but they remove "the oracle and the kernels share one
author's misunderstanding" for the pieces that have a public definition:

  * attention (``oracle.llama_ref._attend`` / ``attention_decode``)  vs  torch.nn.functional.scaled_dot_product_attention
    in float64 (causal, GQA by head repetition, ragged visibility);
  * exact int32 accumulation (``gemm_i8_acc``, numpy and C paths)     vs  torch._int_mm (where the build has it) and int64
    einsum;
  * the sampler (``oracle.sampler_ref.sample_topk_topp``)             vs  a torch restatement built from topk / softmax /
    cumsum / log_softmax;
  * RMSNorm and rotate-half RoPE                                      vs  their textbook float64 forms.
(tests/test_oracle_vs_hf.py checks the whole forward against Hugging Face's LlamaForCausalLM.)
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import llama_ref as ref
from oracle import sampler_ref
from oracle.weights import ModelDesc

F = torch.nn.functional


@pytest.mark.parametrize("hq,hkv", [(4, 4), (8, 2), (8, 1)])
def test_attend_matches_torch_sdpa_fp64(hq, hkv):
    rng = np.random.default_rng(hq * 10 + hkv)
    D, t, n, prefix = 64, 37, 9, 28          # 9 queries at absolute positions 28 .. 36 over 37 keys (prefix + fresh)
    q = rng.standard_normal((n, hq, D)).astype(np.float32)
    K = rng.standard_normal((t, hkv, D)).astype(np.float32)
    V = rng.standard_normal((t, hkv, D)).astype(np.float32)
    got = ref._attend(q, K, V, prefix + np.arange(n))
    rep = hq // hkv
    qt = torch.from_numpy(q).double().permute(1, 0, 2)[None]                                  # [1, hq, n, D]
    kt = torch.from_numpy(K).double().permute(1, 0, 2).repeat_interleave(rep, dim=0)[None]    # [1, hq, t, D]
    vt = torch.from_numpy(V).double().permute(1, 0, 2).repeat_interleave(rep, dim=0)[None]
    mask = torch.arange(t)[None, :] <= (prefix + torch.arange(n))[:, None]                    # True = visible
    exp = F.scaled_dot_product_attention(qt, kt, vt, attn_mask=mask)[0].permute(1, 0, 2).numpy()
    np.testing.assert_allclose(got, exp, rtol=0, atol=2e-6 * max(1.0, np.abs(exp).max()))


def test_attention_decode_over_int8_and_fp16_cache_matches_sdpa():
    """decode attention = one query over the cache: against SDPA over the values the cache defines (dequantised int8,
    or the stored fp16), paged slots in a shuffled order"""
    rng = np.random.default_rng(5)
    for bit, group in ((8, 8), (0, 1)):
        desc = ModelDesc(256, 512, 1, 4, 2, 64, cache_layout=3, cache_mode=1, page_size=8, max_position=64,
                         cache_quant_bit=bit, cache_quant_group=group)
        cache = ref.KVCache(desc, 64)
        slots = np.array([40, 41, 42, 43, 44, 45, 46, 47, 8, 9, 10, 11, 12], dtype=np.int64)
        k = rng.standard_normal((len(slots), 2, 64)).astype(np.float16)
        v = rng.standard_normal((len(slots), 2, 64)).astype(np.float16)
        cache.append(0, slots, k, v)
        q = rng.standard_normal((4, 64)).astype(np.float16)
        got = ref.attention_decode(q, cache, 0, slots, group)
        Kd, Vd = cache.read_values(0, 0, slots), cache.read_values(0, 1, slots)
        if bit == 0:
            assert np.array_equal(Kd, k.astype(np.float32)) and np.array_equal(Vd, v.astype(np.float32))
        else:  # int8 group-8: |dequantised - original| <= half a quantisation step of the group (+ fp16 rounding)
            step = np.abs(k.astype(np.float32)).reshape(len(slots), 2, 8, 8).max(-1, keepdims=True) / 127
            assert (np.abs(Kd - k.astype(np.float32)).reshape(len(slots), 2, 8, 8) <= 0.51 * step + 1e-3).all()
        qt = torch.from_numpy(q.astype(np.float64))[None, :, None, :]                       # [1, 4, 1, D]
        kt = torch.from_numpy(Kd.astype(np.float64)).permute(1, 0, 2).repeat_interleave(2, dim=0)[None]
        vt = torch.from_numpy(Vd.astype(np.float64)).permute(1, 0, 2).repeat_interleave(2, dim=0)[None]
        exp = F.scaled_dot_product_attention(qt, kt, vt)[0, :, 0, :].numpy()
        np.testing.assert_allclose(got, exp, rtol=0, atol=2e-6 * max(1.0, np.abs(exp).max()))


def test_int8_gemm_accumulation_is_exact():
    rng = np.random.default_rng(11)
    a = rng.integers(-127, 128, (33, 4096), dtype=np.int8)
    w = rng.integers(-127, 128, (40, 4096), dtype=np.int8)
    exp = np.einsum("mk,nk->mn", a.astype(np.int64), w.astype(np.int64))
    assert np.abs(exp).max() < 2 ** 31
    assert np.array_equal(ref.gemm_i8_acc_numpy(a, w).astype(np.int64), exp)
    assert np.array_equal(ref.gemm_i8_acc(a, w).astype(np.int64), exp)           # C restatement when built
    if hasattr(torch, "_int_mm"):
        try:  # CPU support depends on the build; shapes need M > 16 and K, N multiples of 8
            got = torch._int_mm(torch.from_numpy(a), torch.from_numpy(np.ascontiguousarray(w.T))).numpy()
        except RuntimeError:
            got = None
        if got is not None:
            assert np.array_equal(got.astype(np.int64), exp)
    # the dequant epilogue: two fp32 multiplies in a fixed order
    sa = rng.uniform(0.001, 0.02, 33).astype(np.float32)
    sw = rng.uniform(0.0001, 0.001, 40).astype(np.float32)
    deq = ref.dequant_acc(exp.astype(np.int32), sa, sw)
    t = (torch.from_numpy(exp.astype(np.int32)).float() * torch.from_numpy(sa)[:, None]) * torch.from_numpy(sw)[None, :]
    assert np.array_equal(deq, t.numpy())


def _torch_sampler(logits, temps, top_p, rand, vocab, top_k, default_top_p):
    """the sampler's semantics from torch primitives (float32 like the kernel; ties need care: topk's order among equal
    values is unspecified, so the test data has no ties among the top candidates)"""
    lg = torch.from_numpy(np.asarray(logits, np.float32))[:, :vocab]
    out, lps = [], []
    for b in range(lg.shape[0]):
        l = lg[b] / (1.0 if temps is None else float(temps[b]))
        val, idx = torch.topk(l, min(top_k, vocab))
        p = torch.softmax(val, dim=0)
        cum = torch.cumsum(p, dim=0)
        tp = default_top_p if top_p is None else float(top_p[b])
        if tp <= 0:
            keep = 1
        else:
            reach = torch.nonzero(cum >= tp)
            keep = int(reach[0]) + 1 if len(reach) else len(val)
        thr = float(rand[b]) * float(cum[keep - 1])
        sel = torch.nonzero(cum[:keep] > thr)
        i = int(sel[0]) if len(sel) else keep - 1
        out.append(int(idx[i]))
        lps.append(float(torch.log_softmax(l.double(), dim=0)[idx[i]]))
    return np.array(out), np.array(lps)


@pytest.mark.parametrize("top_k,top_p", [(1, 0.0), (5, 0.8), (40, 0.95), (8, 1.0)])
def test_sampler_matches_torch_restatement(top_k, top_p):
    rng = np.random.default_rng(top_k)
    B, V = 16, 1000
    logits = (3 * rng.standard_normal((B, V + 24))).astype(np.float32)   # row stride > vocab, as the runtime's output may be
    temps = rng.uniform(0.5, 1.5, B).astype(np.float32)
    tps = np.full(B, top_p, np.float32)
    rand = rng.uniform(0, 1, B).astype(np.float32)
    # keep the inverse-CDF draw away from a cumulative boundary (fp32 summation order is the only difference)
    got, glp = sampler_ref.sample_topk_topp(logits, temps, tps, rand, V, top_k, top_p)
    exp, elp = _torch_sampler(logits, temps, tps, rand, V, top_k, top_p)
    agree = got == exp
    assert agree.mean() >= 0.9, (got, exp)                    # disagreements only where rand sits on a boundary
    np.testing.assert_allclose(glp[agree], elp[agree], atol=2e-5)
    if top_k == 1:
        assert agree.all() and np.array_equal(got, (logits[:, :V] / temps[:, None]).argmax(1))


def test_rmsnorm_and_rope_match_textbook_float64():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((5, 256)).astype(np.float16)
    g = (1 + 0.02 * rng.standard_normal(256)).astype(np.float16)
    y = ref.rmsnorm_f32(x, g, 1e-5)
    x64 = x.astype(np.float64)
    exp = x64 / np.sqrt((x64 ** 2).mean(-1, keepdims=True) + 1e-5) * g.astype(np.float64)
    np.testing.assert_allclose(y, exp, rtol=3e-7, atol=1e-7)
    # rotate-half RoPE == complex rotation of the pairs (i, i + D/2) by pos * theta^(-2i/D)
    D, T, H = 64, 7, 3
    q = rng.standard_normal((T, H, D)).astype(np.float16)
    pos = np.array([0, 1, 2, 3, 10, 100, 255])
    cos, sin = ref.rope_table(256, D, 10000.0)
    got = ref.apply_rope(q, pos, cos, sin).astype(np.float64)
    ang = pos[:, None] * (10000.0 ** (-2.0 * np.arange(D // 2) / D))[None, :]
    z = (q[..., : D // 2].astype(np.float64) + 1j * q[..., D // 2:].astype(np.float64)) * np.exp(1j * ang)[:, None, :]
    exp = np.concatenate([z.real, z.imag], axis=-1)
    np.testing.assert_allclose(got, exp, rtol=0, atol=2e-3 * np.abs(exp).max())   # fp16 output rounding
