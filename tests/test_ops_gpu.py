"""Operator-level parity: every CUDA kernel, called through the C ABI, against the oracle.

Bars: bit-exact for integer / index / fp16-rounded deterministic outputs (quantised values, int32
GEMM results after the fixed-order fp32 dequant, rope, KV bytes); for kernels whose fp32 reduction
order legitimately differs (row norms, softmax, fp16 GEMM) the tolerance is stated at the assert.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import llama_ref as ref
from oracle import sampler_ref
from oracle.weights import ModelDesc, synth_tensor, quantize_weight_per_channel
from ppl_llm_serving_b200 import capi
from ppl_llm_serving_b200.engine import _ptr
from helpers import dev as _dev, stream_ptr, make_step_c, make_geom, random_pages

_KEEP = []


def dev(a):
    """device copy that stays alive until the test ends (a temporary passed as _ptr(dev(x)) would be
    freed -- and its block reused by the next allocation -- before the kernel runs)"""
    t = _dev(a)
    _KEEP.append(t)
    return t


@pytest.fixture(autouse=True)
def _release_keep():
    yield
    torch.cuda.synchronize()
    _KEEP.clear()

pytestmark = pytest.mark.gpu


def sync():
    torch.cuda.synchronize()


# ------------------------------------------------------------------ weights
def test_synth_matches_oracle(lib):
    n = 100003
    out = torch.empty(n, dtype=torch.float16, device="cuda")
    for tid, std, mean in ((1, 1.0, 0.0), (19, 0.02, 0.0), (2, 0.02, 1.0)):
        capi.check(lib.b2llm_op_synth_fp16(stream_ptr(), 0xB200, tid, n, std, mean, _ptr(out)))
        sync()
        exp = synth_tensor(0xB200, tid, (n,), std, mean)
        assert np.array_equal(out.cpu().numpy().view(np.uint16), exp.view(np.uint16))  # bit-exact


def test_quant_weight_bit_exact(lib):
    w = synth_tensor(3, 77, (384, 512), 0.02)
    w[5, :] = 0  # an all-zero channel
    q = torch.empty((384, 512), dtype=torch.int8, device="cuda")
    s = torch.empty(384, dtype=torch.float32, device="cuda")
    capi.check(lib.b2llm_op_quant_weight(stream_ptr(), _ptr(dev(w)), 384, 512, _ptr(q), _ptr(s)))
    sync()
    eq, es = quantize_weight_per_channel(w)
    assert np.array_equal(q.cpu().numpy(), eq)
    assert np.array_equal(s.cpu().numpy(), es)


# ------------------------------------------------------------------ norm / quant
def test_w4a16_weight_quant_bit_exact(lib):
    """int4 group-128 weight quantisation and its fp16 operand expansion (builder-defined W4A16)"""
    from oracle.weights import quantize_weight_w4
    rng = np.random.default_rng(4)
    N, K = 96, 512
    w = (0.02 * rng.standard_normal((N, K))).astype(np.float16)
    w[5, 128:256] = 0          # an all-zero group: scale 0, codes 0 (nibble 8)
    w[7, 3] = np.float16(0.5)  # an outlier dominating its group
    q, s16, deq = quantize_weight_w4(w)
    packed = torch.zeros((N, K // 2), dtype=torch.uint8, device="cuda")
    scale = torch.zeros((N, K // 128), dtype=torch.float16, device="cuda")
    capi.check(lib.b2llm_op_quant_weight_w4(stream_ptr(), _ptr(dev(w)), N, K, _ptr(packed), _ptr(scale)))
    sync()
    pk = packed.cpu().numpy()
    assert np.array_equal((pk & 0xF).astype(np.int8) - 8, q[:, 0::2]) and np.array_equal((pk >> 4).astype(np.int8) - 8, q[:, 1::2])
    assert np.array_equal(scale.cpu().numpy().view(np.uint16), s16.view(np.uint16))
    out = torch.zeros((N, K), dtype=torch.float16, device="cuda")
    capi.check(lib.b2llm_op_dequant_w4(stream_ptr(), _ptr(packed), _ptr(scale), N, K, _ptr(out)))
    sync()
    assert np.array_equal(out.cpu().numpy().view(np.uint16), deq.view(np.uint16))


# (256, 1280, 8192): one A tile per weight tile + 7 k-slices; (129, 384, 1024): 2 slices with a partial M tile;
# (256, 3584, 8192): two A tiles per weight tile + 5 k-slices of 25 k-blocks (the gate_up / down plan of 70B at TP = 8, half the N)
# two k-slices summed inside the kernel by clusters of two CTAs (50 .. 74 tiles): (200, 8192, 3584) = the down projection of
# 70B at TP = 8 with a ragged M, (256, 7168, 1024) = gate_up's width, (100, 8192, 1024) the same with the 128-row tile
@pytest.mark.parametrize("M,N,K", [(1, 128, 128), (100, 256, 512), (256, 1280, 8192), (129, 384, 1024), (300, 512, 256),
                                   (256, 3584, 8192), (200, 8192, 3584), (256, 7168, 1024), (100, 8192, 1024), (64, 256, 256),
                                   # more work items than SMs: CTAs run a second item over the same accumulator (160 tiles of the
                                   # 128-row tile; 2 row blocks x 100 = 200 tiles of the 256-row tile)
                                   (64, 20480, 256), (300, 12800, 256)])
def test_gemm_w4a16_fused(lib, M, N, K):
    """fused W4A16 tcgen05 GEMM (nibbles expanded to fp16(q * scale) by converter warps inside the kernel) against
    the oracle's definition; the operand values are bit-identical, only the fp32 accumulation order differs"""
    from oracle.weights import quantize_weight_w4
    rng = np.random.default_rng(M + N + K)
    w = (0.02 * rng.standard_normal((N, K))).astype(np.float16)
    a = rng.standard_normal((M, K)).astype(np.float16)
    q, s16, deq = quantize_weight_w4(w)
    packed = torch.zeros((N, K // 2), dtype=torch.uint8, device="cuda")
    scale = torch.zeros((N, K // 128), dtype=torch.float16, device="cuda")
    capi.check(lib.b2llm_op_quant_weight_w4(stream_ptr(), _ptr(dev(w)), N, K, _ptr(packed), _ptr(scale)))
    out = torch.zeros((M, N), dtype=torch.float16, device="cuda")
    capi.check(lib.b2llm_op_gemm_w4a16(stream_ptr(), _ptr(dev(a)), _ptr(packed), _ptr(scale), M, N, K, capi.EPI_F16, _ptr(out)))
    sync()
    exp = ref.gemm_f16_acc(a, deq)
    got = out.cpu().numpy().astype(np.float32)
    assert np.abs(got - exp).max() <= 1e-3 * np.abs(exp).max() + 1e-3   # fp16 output rounding (2^-11 relative)
    # residual and SwiGLU epilogues
    res = rng.standard_normal((M, N)).astype(np.float16)
    out2 = dev(res.copy())
    capi.check(lib.b2llm_op_gemm_w4a16(stream_ptr(), _ptr(dev(a)), _ptr(packed), _ptr(scale), M, N, K, capi.EPI_RESIDUAL, _ptr(out2)))
    out3 = torch.zeros((M, N // 2), dtype=torch.float16, device="cuda")
    capi.check(lib.b2llm_op_gemm_w4a16(stream_ptr(), _ptr(dev(a)), _ptr(packed), _ptr(scale), M, N, K, capi.EPI_SWIGLU, _ptr(out3)))
    sync()
    exp2 = res.astype(np.float32) + exp
    assert np.abs(out2.cpu().numpy().astype(np.float32) - exp2).max() <= 1e-3 * np.abs(exp2).max() + 2e-3
    exp3 = ref.silu_mul(exp[:, 0::2], exp[:, 1::2])
    np.testing.assert_allclose(out3.cpu().numpy().astype(np.float32), exp3, rtol=3e-3, atol=2e-3)


@pytest.mark.parametrize("rows,hidden", [(1, 256), (37, 4096), (5, 5120), (3, 11008), (300, 4096), (257, 8192), (1024, 512)])
def test_rmsnorm_quant(lib, rows, hidden):
    rng = np.random.default_rng(rows * 7 + hidden)
    x = rng.standard_normal((rows, hidden)).astype(np.float16)
    g = (1 + 0.02 * rng.standard_normal(hidden)).astype(np.float16)
    xd = dev(x)
    q = torch.empty((rows, hidden), dtype=torch.int8, device="cuda")
    s = torch.empty(rows, dtype=torch.float32, device="cuda")
    capi.check(lib.b2llm_op_rmsnorm_quant(stream_ptr(), _ptr(xd), None, _ptr(dev(g)), 1e-5, rows, hidden, _ptr(q), _ptr(s), None))
    sync()
    y = ref.rmsnorm_f32(x, g, 1e-5)
    eq, es = ref.quant_rows(y)
    # fp32 sum-of-squares order differs (tree vs float64): scale to 1e-6 relative, q within 1 step on
    # a vanishing fraction of entries
    np.testing.assert_allclose(s.cpu().numpy(), es, rtol=2e-6)
    diff = np.abs(q.cpu().numpy().astype(np.int32) - eq.astype(np.int32))
    assert diff.max() <= 1 and (diff != 0).mean() < 1e-3


@pytest.mark.parametrize("rows", [9, 300])  # CTA-per-row kernels / register-resident kernels
def test_rmsnorm_skip_and_fp16_out(lib, rows):
    rng = np.random.default_rng(5)
    hidden = 512
    x = rng.standard_normal((rows, hidden)).astype(np.float16)
    sk = rng.standard_normal((rows, hidden)).astype(np.float16)
    g = (1 + 0.02 * rng.standard_normal(hidden)).astype(np.float16)
    xd = dev(x)
    y = torch.empty((rows, hidden), dtype=torch.float16, device="cuda")
    capi.check(lib.b2llm_op_rmsnorm_quant(stream_ptr(), _ptr(xd), _ptr(dev(sk)), _ptr(dev(g)), 1e-5, rows, hidden, None, None, _ptr(y)))
    sync()
    xs = (x.astype(np.float32) + sk.astype(np.float32)).astype(np.float16)
    assert np.array_equal(xd.cpu().numpy().view(np.uint16), xs.view(np.uint16))  # residual join bit-exact
    ey = ref.rmsnorm_f32(xs, g, 1e-5)
    np.testing.assert_allclose(y.cpu().numpy().astype(np.float32), ey, rtol=2e-3, atol=1e-3)  # fp16 output rounding


@pytest.mark.parametrize("rows,cols", [(1, 128), (33, 4096), (4, 11008), (300, 11008), (1024, 4096)])  # >= 256 rows: register-resident kernels
def test_quant_rows_bit_exact(lib, rows, cols):
    rng = np.random.default_rng(cols)
    x = rng.standard_normal((rows, cols)).astype(np.float16)
    x[0, :] = 0
    q = torch.empty((rows, cols), dtype=torch.int8, device="cuda")
    s = torch.empty(rows, dtype=torch.float32, device="cuda")
    capi.check(lib.b2llm_op_quant_rows(stream_ptr(), _ptr(dev(x)), rows, cols, _ptr(q), _ptr(s)))
    sync()
    eq, es = ref.quant_rows(x.astype(np.float32))
    assert np.array_equal(q.cpu().numpy(), eq)
    assert np.array_equal(s.cpu().numpy(), es)


# ------------------------------------------------------------------ GEMM
def _gemm_inputs(M, N, K, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(-127, 128, (M, K), dtype=np.int8)
    w = rng.integers(-127, 128, (N, K), dtype=np.int8)
    sa = rng.uniform(0.001, 0.02, M).astype(np.float32)
    sw = rng.uniform(0.0001, 0.001, N).astype(np.float32)
    return a, w, sa, sw


@pytest.mark.parametrize("impl", [1, 2, 3])  # 3 = CTA-pair kernel (cta_group::2) where the shape allows it
@pytest.mark.parametrize("M,N,K", [(1, 128, 64), (17, 384, 256), (128, 256, 4096), (300, 1280, 1024), (1024, 512, 11008),
                                   (513, 768, 256), (1024, 4096, 4096), (2048, 22016, 512),
                                   # per-rank shapes of LLaMA-2-7B under tensor parallelism: down_proj at TP = 4 / 8 (K not a
                                   # multiple of the 128-byte k-block), gate_up at TP = 8 (N = 2752: partial last tile)
                                   (1024, 4096, 2752), (1024, 4096, 1376), (512, 2752, 4096)])
def test_gemm_w8a8_f16_bit_exact(lib, impl, M, N, K):
    if impl >= 2 and not _tc_ok(lib):
        pytest.skip("tcgen05 path not built")
    if impl == 1 and K % 64 != 0:
        pytest.skip("K outside the mma.sync baseline's envelope")
    a, w, sa, sw = _gemm_inputs(M, N, K, M + N + K)
    out = torch.zeros((M, N), dtype=torch.float16, device="cuda")
    rc = lib.b2llm_op_gemm_w8a8(stream_ptr(), _ptr(dev(a)), _ptr(dev(sa)), _ptr(dev(w)), _ptr(dev(sw)), M, N, K,
                                capi.EPI_F16, _ptr(out), impl)
    if impl >= 2 and rc == 4 and K % 128 != 0:
        pytest.skip("K outside the tcgen05 kernel's envelope (falls back to the mma.sync kernel when impl = 0)")
    capi.check(rc, "gemm")
    sync()
    exp = ref.dequant_acc(ref.gemm_i8_acc(a, w), sa, sw).astype(np.float16)
    assert np.array_equal(out.cpu().numpy().view(np.uint16), exp.view(np.uint16))


def _tc_ok(lib):
    a, w, sa, sw = _gemm_inputs(128, 128, 128, 0)
    out = torch.zeros((128, 128), dtype=torch.float16, device="cuda")
    rc = lib.b2llm_op_gemm_w8a8(stream_ptr(), _ptr(dev(a)), _ptr(dev(sa)), _ptr(dev(w)), _ptr(dev(sw)), 128, 128, 128,
                                capi.EPI_F16, _ptr(out), 2)
    torch.cuda.synchronize()
    return rc == 0


@pytest.mark.parametrize("impl,M", [(1, 67), (2, 67), (2, 400), (3, 400), (3, 1024)])
def test_gemm_w8a8_residual_and_swiglu(lib, impl, M):
    if impl >= 2 and not _tc_ok(lib):
        pytest.skip("tcgen05 path not built")
    N, K = 512, 512
    a, w, sa, sw = _gemm_inputs(M, N, K, 9)
    rng = np.random.default_rng(1)
    res = rng.standard_normal((M, N)).astype(np.float16)
    out = dev(res.copy())
    capi.check(lib.b2llm_op_gemm_w8a8(stream_ptr(), _ptr(dev(a)), _ptr(dev(sa)), _ptr(dev(w)), _ptr(dev(sw)), M, N, K,
                                      capi.EPI_RESIDUAL, _ptr(out), impl))
    sync()
    v = ref.dequant_acc(ref.gemm_i8_acc(a, w), sa, sw)
    exp = (res.astype(np.float32) + v).astype(np.float16)
    assert np.array_equal(out.cpu().numpy().view(np.uint16), exp.view(np.uint16))
    # SwiGLU over interleaved (gate, up) pairs
    out2 = torch.zeros((M, N // 2), dtype=torch.float16, device="cuda")
    capi.check(lib.b2llm_op_gemm_w8a8(stream_ptr(), _ptr(dev(a)), _ptr(dev(sa)), _ptr(dev(w)), _ptr(dev(sw)), M, N, K,
                                      capi.EPI_SWIGLU, _ptr(out2), impl))
    sync()
    exp2 = ref.silu_mul(v[:, 0::2], v[:, 1::2])
    # expf differs from numpy's exp by <= 2 ulp before the fp16 rounding
    np.testing.assert_allclose(out2.cpu().numpy().astype(np.float32), exp2, rtol=2e-3, atol=1e-4)


@pytest.mark.parametrize("impl", [1, 2, 3])
@pytest.mark.parametrize("M,N,K", [(3, 256, 128), (130, 1024, 512), (64, 32000, 4096), (1024, 32000, 512),
                                   (256, 4000, 1024)])  # vocab / 8: the vocab-parallel lm head at TP = 8
def test_gemm_f16_logits(lib, impl, M, N, K):
    rng = np.random.default_rng(N)
    a = rng.standard_normal((M, K)).astype(np.float16)
    w = (0.02 * rng.standard_normal((N, K))).astype(np.float16)
    out = torch.zeros((M, N), dtype=torch.float32, device="cuda")
    rc = lib.b2llm_op_gemm_f16(stream_ptr(), _ptr(dev(a)), _ptr(dev(w)), M, N, K, capi.EPI_F32, _ptr(out), N, impl)
    if impl >= 2 and rc == 4:
        pytest.skip("tcgen05 path not built")
    capi.check(rc)
    sync()
    exp = ref.gemm_f16_acc(a, w)
    # fp32 accumulation order differs: |err| <= 1e-3 * max|logit| is the north-star tolerance; observed ~1e-6
    assert np.abs(out.cpu().numpy() - exp).max() <= 1e-4 * np.abs(exp).max()


# ------------------------------------------------------------------ rope + KV append
def _mk_desc(layout, mode, nq=4, nkv=2, D=128, layers=2, page=16, kvbit=8):
    """kvbit 8: int8 cache, fp16 scale per 8 elements; kvbit 0: fp16 cache, no scale tensor (llm_generator.cc:131-136)"""
    return ModelDesc(nq * D, 256, layers, nq, nkv, 512, cache_layout=layout, cache_mode=mode, page_size=page,
                     max_position=512, cache_quant_bit=kvbit, cache_quant_group=8 if kvbit else 1)


def _ragged_step(desc, rng, seqlens, start_pos, decoding, T_cache):
    B = len(seqlens)
    toks = [list(rng.integers(0, desc.vocab_size, n)) for n in seqlens]
    if desc.cache_mode == 1:
        ps = desc.page_size
        need = max((sp + n + ps - 1) // ps for sp, n in zip(start_pos, seqlens))
        pages = random_pages(rng, B, need, ps, T_cache // ps)
        return ref.build_step(desc, toks, start_pos, decoding, page_tables=pages)
    stride = max(sp + n for sp, n in zip(start_pos, seqlens))
    return ref.build_step(desc, toks, start_pos, decoding, cache_indices=[i * stride for i in range(B)])


@pytest.mark.parametrize("kvbit", [8, 0])
@pytest.mark.parametrize("layout", [0, 1, 2, 3])
@pytest.mark.parametrize("mode", [0, 1])
def test_rope_kv_append_bit_exact(lib, layout, mode, kvbit):
    desc = _mk_desc(layout, mode, kvbit=kvbit)
    rng = np.random.default_rng(layout * 2 + mode)
    T_cache = 512
    step = _ragged_step(desc, rng, [1, 1, 7, 20], [30, 5, 0, 3], 2, T_cache)
    D, nq, nkv = desc.head_dim, desc.num_heads, desc.num_kv_heads
    T = len(step.token_inputs)
    qkv = rng.standard_normal((T, (nq + 2 * nkv) * D)).astype(np.float16)
    cos, sin = ref.rope_table(desc.max_position, D, desc.rope_theta)
    # the engine's own table must equal the oracle's
    c2 = np.empty_like(cos); s2 = np.empty_like(sin)
    capi.check(lib.b2llm_rope_table(desc.max_position, D, desc.rope_theta, _ptr(c2), _ptr(s2)))
    # libm vs numpy double cos/sin may differ in the last bit before the fp32 rounding: <= 1 ulp, rare
    for mine, theirs in ((c2, cos), (s2, sin)):
        assert (mine != theirs).mean() < 1e-4 and np.abs(mine.view(np.int32) - theirs.view(np.int32)).max() <= 1
    cos, sin = c2, s2  # both sides of the comparison below use the engine's table

    cache = ref.KVCache(desc, T_cache)
    c_np, s_np = cache.export()
    cd, sd = dev(c_np), dev(s_np)
    keep = []
    sc = make_step_c(step, keep)
    geom = make_geom(desc, T_cache)
    qd = dev(qkv)
    layer = 1
    capi.check(lib.b2llm_op_rope_kv_append(stream_ptr(), _ptr(qd), C.byref(sc), nq, C.byref(geom), layer, _ptr(dev(cos)),
                                           _ptr(dev(sin)), _ptr(cd), _ptr(sd)))
    sync()
    # oracle
    seqlens = np.diff(step.seq_starts)
    pos = np.concatenate([step.start_pos[b] + np.arange(seqlens[b]) for b in range(step.batch)])
    slots = np.concatenate([step.slots(desc, b, step.start_pos[b] + np.arange(seqlens[b])) for b in range(step.batch)])
    q = ref.apply_rope(qkv[:, :nq * D].reshape(T, nq, D), pos, cos, sin)
    k = ref.apply_rope(qkv[:, nq * D:(nq + nkv) * D].reshape(T, nkv, D), pos, cos, sin)
    v = qkv[:, (nq + nkv) * D:].reshape(T, nkv, D)
    cache.append(layer, slots, k, v)
    ec, es = cache.export()
    got = qd.cpu().numpy()
    assert np.array_equal(got[:, :nq * D].view(np.uint16), q.reshape(T, -1).view(np.uint16))
    assert np.array_equal(got[:, nq * D:(nq + nkv) * D].view(np.uint16), k.reshape(T, -1).view(np.uint16))
    assert np.array_equal(cd.cpu().numpy().view(np.uint8), ec.view(np.uint8))   # int8 values, or the fp16 rows bit for bit
    assert np.array_equal(sd.cpu().numpy().view(np.uint16), es.view(np.uint16))


# ------------------------------------------------------------------ attention
def _attention_case(lib, desc, seqlens, start_pos, decoding, impl, T_cache=2048, seed=0, cache_prefill=1):
    rng = np.random.default_rng(seed)
    step = _ragged_step(desc, rng, seqlens, start_pos, decoding, T_cache)
    D, nq, nkv = desc.head_dim, desc.num_heads, desc.num_kv_heads
    T = len(step.token_inputs)
    layer = desc.num_layers - 1
    cache = ref.KVCache(desc, T_cache)
    # history for every sequence
    for b in range(step.batch):
        sp = int(step.start_pos[b])
        if sp:
            sl = step.slots(desc, b, np.arange(sp))
            kh = rng.standard_normal((sp, nkv, D)).astype(np.float16)
            vh = rng.standard_normal((sp, nkv, D)).astype(np.float16)
            cache.append(layer, sl, kh, vh)
    qkv = rng.standard_normal((T, (nq + 2 * nkv) * D)).astype(np.float16)
    seqlens_a = np.diff(step.seq_starts)
    slots = np.concatenate([step.slots(desc, b, step.start_pos[b] + np.arange(seqlens_a[b])) for b in range(step.batch)])
    k = qkv[:, nq * D:(nq + nkv) * D].reshape(T, nkv, D)
    v = qkv[:, (nq + nkv) * D:].reshape(T, nkv, D)
    cache.append(layer, slots, k, v)
    q = qkv[:, :nq * D].reshape(T, nq, D)
    exp = np.empty((T, nq, D), dtype=np.float32)
    for b in range(step.batch):
        t0, t1 = step.seq_starts[b], step.seq_starts[b + 1]
        sp, n = int(step.start_pos[b]), int(t1 - t0)
        if b < decoding:
            exp[t0] = ref.attention_decode(q[t0], cache, layer, step.slots(desc, b, np.arange(sp + 1)), 8)
        else:
            Kf, Vf = k[t0:t1].astype(np.float32), v[t0:t1].astype(np.float32)
            if sp:
                sl = step.slots(desc, b, np.arange(sp))
                Kf = np.concatenate([cache.read_values(layer, 0, sl), Kf])
                Vf = np.concatenate([cache.read_values(layer, 1, sl), Vf])
            exp[t0:t1] = ref._attend(q[t0:t1].astype(np.float32), Kf, Vf, sp + np.arange(n))
    c_np, s_np = cache.export()
    keep = []
    sc = make_step_c(step, keep)
    sc.cache_prefill = cache_prefill
    geom = make_geom(desc, T_cache)
    out = torch.zeros((T, nq * D), dtype=torch.float16, device="cuda")
    ws = torch.empty(lib.b2llm_attention_workspace_size(step.batch, nq, D), dtype=torch.uint8, device="cuda")
    capi.check(lib.b2llm_op_attention(stream_ptr(), _ptr(dev(qkv)), C.byref(sc), nq, C.byref(geom), layer, _ptr(dev(c_np)),
                                      _ptr(dev(s_np)), _ptr(ws), _ptr(out), impl), "attention")
    sync()
    got = out.cpu().numpy().astype(np.float32).reshape(T, nq, D)
    # fp16 output (2^-11 relative) + fp16 P / dequantised K,V operands in the tensor-core path:
    # tolerance 2e-3 of the output scale, the north-star "1e-3 relative fp16" bar at logits level
    tol = 2e-3 * max(1.0, np.abs(exp).max())
    err = np.abs(got - exp).max()
    assert err <= tol, f"max err {err} > {tol}"
    return err


@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("layout,mode", [(3, 1), (0, 1), (1, 0), (2, 0), (3, 0)])
def test_attention_decode_mha(lib, impl, layout, mode):
    desc = _mk_desc(layout, mode, nq=4, nkv=4)
    _attention_case(lib, desc, [1] * 5, [0, 15, 16, 100, 333], 5, impl, seed=layout)


@pytest.mark.parametrize("impl", [1, 2])
def test_attention_decode_gqa_and_splits(lib, impl):
    desc = _mk_desc(3, 1, nq=8, nkv=1)  # 8 q heads share one kv head (70B TP=8 shape per rank)
    _attention_case(lib, desc, [1, 1], [1500, 700], 2, impl, T_cache=4096, seed=3)  # few CTAs -> split-KV + merge
    desc = _mk_desc(3, 1, nq=8, nkv=2)
    _attention_case(lib, desc, [1] * 3, [40, 1, 257], 3, impl, seed=4)


@pytest.mark.parametrize("impl", [1, 2, 8])
def test_attention_mixed_prefill_decode(lib, impl):
    desc = _mk_desc(3, 1, nq=4, nkv=2)
    # 2 decoding sequences first, then a fresh prompt and a prompt with a cached prefix
    _attention_case(lib, desc, [1, 1, 9, 6], [64, 7, 0, 32], 2, impl, seed=11)


@pytest.mark.parametrize("impl", [1, 2, 8])  # 8: the mma.sync prefill kernel forced (2 = default = tcgen05 where applicable)
@pytest.mark.parametrize("nq,nkv", [(4, 4), (8, 2)])
def test_attention_prefill_tiles_and_prefixes(lib, impl, nq, nkv):
    """prefill sequences spanning several 64-query tiles / 64-key blocks, tile-boundary lengths, cached prefixes that
    are not multiples of the block, mixed with decode sequences (tensor-core flash-attention kernel for impl 2)"""
    desc = _mk_desc(3, 1, nq=nq, nkv=nkv)
    _attention_case(lib, desc, [1, 200, 64, 65, 1, 130], [77, 0, 0, 48, 0, 100], 1, impl, T_cache=4096, seed=21)
    desc = _mk_desc(1, 0, nq=nq, nkv=nkv)
    _attention_case(lib, desc, [300, 17], [0, 250], 0, impl, T_cache=4096, seed=22)


def test_attention_long_ragged_batch(lib):
    desc = _mk_desc(3, 1, nq=4, nkv=4, page=16)
    rng = np.random.default_rng(0)
    lens = [int(x) for x in rng.integers(1, 400, 48)]
    _attention_case(lib, desc, [1] * 48, lens, 48, 2, T_cache=48 * 416, seed=5)


@pytest.mark.parametrize("impl", [1, 2, 5, 7])  # simple kernel; tensor-core kernel: merged (default) / slim TMA loaders, cp.async
@pytest.mark.parametrize("layout,mode,page", [(3, 1, 16), (2, 0, 16), (1, 1, 64), (0, 0, 16), (3, 1, 8)])
def test_attention_fp16_cache(lib, impl, layout, mode, page):
    """cache_quant_bit 0 / group 1 (llm_generator.cc:131-136): K and V are cached as fp16, no scale tensor.  Decode (MHA,
    GQA 8 with split-KV + merge, ragged batch) and prefill with a cached prefix, all layouts; page size 8 forces the
    cp.async loader (units are not 16 contiguous slots)"""
    if impl == 7 and page != 8:
        pytest.skip("impl 7 = default dispatch on a page size the TMA loaders cannot serve")
    if impl != 7 and page == 8 and impl != 1:
        pytest.skip("TMA loaders need page_size % 16 == 0")
    impl = 2 if impl == 7 else impl
    kw = dict(page=page, kvbit=0)
    _attention_case(lib, _mk_desc(layout, mode, nq=4, nkv=4, **kw), [1] * 5, [0, 15, 16, 100, 333], 5, impl, seed=layout)
    _attention_case(lib, _mk_desc(layout, mode, nq=8, nkv=1, **kw), [1, 1], [1500, 700], 2, impl, T_cache=4096, seed=3)
    rng = np.random.default_rng(0)
    lens = [int(x) for x in rng.integers(1, 400, 48)]
    _attention_case(lib, _mk_desc(layout, mode, nq=4, nkv=4, **kw), [1] * 48, lens, 48, impl, T_cache=48 * 512, seed=5)
    _attention_case(lib, _mk_desc(layout, mode, nq=4, nkv=2, **kw), [1, 1, 9, 70], [64, 7, 0, 32], 2, impl, seed=11)


@pytest.mark.parametrize("impl", [4, 5])  # 4: dividing loader (default until run 17), 5: slim loader (default since)
@pytest.mark.parametrize("layout,mode", [(3, 1), (1, 0)])
def test_attention_decode_both_validated_loaders(lib, impl, layout, mode):
    """the decode kernel's two device-validated TMA loaders, selected explicitly (b2llm_op_attention impl 4 / 5), on
    the shapes of test_attention_decode_mha / _gqa_and_splits / _long_ragged_batch"""
    _attention_case(lib, _mk_desc(layout, mode, nq=4, nkv=4), [1] * 5, [0, 15, 16, 100, 333], 5, impl, seed=layout)
    if (layout, mode) != (3, 1):
        return  # the GQA / split / ragged shapes below ran on the device with layout 3 + paging only (runs 13 and 17)
    _attention_case(lib, _mk_desc(layout, mode, nq=8, nkv=1), [1, 1], [1500, 700], 2, impl, T_cache=4096, seed=3)
    rng = np.random.default_rng(0)
    lens = [int(x) for x in rng.integers(1, 400, 48)]
    _attention_case(lib, _mk_desc(layout, mode, nq=4, nkv=4), [1] * 48, lens, 48, impl, T_cache=48 * 416, seed=5)


@pytest.mark.parametrize("layout,mode,page", [(3, 1, 16), (2, 0, 16), (3, 0, 16), (2, 1, 64), (3, 1, 128), (1, 0, 16)])
def test_attention_decode_merged_loader(lib, layout, mode, page):
    """impl 3: slim loader + K and V of a unit in one 4-D TMA box (LOADER 3, layouts 2 / 3; layout 1 falls back to slim).
    Also the only GPU coverage of the incremental page walk at page sizes other than 16."""
    _attention_case(lib, _mk_desc(layout, mode, nq=4, nkv=4, page=page), [1] * 5, [0, 15, 16, 100, 333], 5, 3, seed=layout)
    _attention_case(lib, _mk_desc(layout, mode, nq=8, nkv=1, page=page), [1, 1], [1500, 700], 2, 3, T_cache=4096, seed=3)
    rng = np.random.default_rng(0)
    lens = [int(x) for x in rng.integers(1, 400, 48)]
    _attention_case(lib, _mk_desc(layout, mode, nq=4, nkv=4, page=page), [1] * 48, lens, 48, 3, T_cache=48 * 512, seed=5)
    _attention_case(lib, _mk_desc(layout, mode, nq=4, nkv=4, page=page), [1] * 5, [0, 15, 16, 100, 333], 5, 5, seed=layout)


@pytest.mark.parametrize("nq,nkv", [(4, 4), (8, 2)])
def test_attention_prefill_tcgen05(lib, nq, nkv):
    """impl 6: the tcgen05 / TMEM prefill kernel (attention_prefill_tc.cu) on fresh prompts whose lengths straddle the
    128-query tiles and 128-key blocks, alone and behind decode sequences; steps with cached prefixes must fall back to
    the mma.sync kernel and still be right"""
    desc = _mk_desc(3, 1, nq=nq, nkv=nkv)
    for lens in ([1], [17], [127], [128], [129], [300, 64, 513], [700]):
        _attention_case(lib, desc, lens, [0] * len(lens), 0, 6, T_cache=4096, seed=31 + len(lens), cache_prefill=0)
    _attention_case(lib, desc, [1, 1, 200, 130], [77, 5, 0, 0], 2, 6, T_cache=4096, seed=41, cache_prefill=0)
    _attention_case(lib, desc, [1, 200, 65], [77, 0, 48], 1, 6, T_cache=4096, seed=42)  # cached prefix -> fallback


# ------------------------------------------------------------------ sampler / penalty
def _sample(lib, logits, temps, top_p, rand, top_k, default_top_p, stride=None):
    B, V = logits.shape
    stride = stride or V
    buf = np.zeros((B, stride), dtype=np.float32)
    buf[:, :V] = logits
    ws_bytes = lib.b2llm_sample_topk_topp_get_workspace_size(B, V, top_k)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device="cuda")
    out = torch.empty(B, dtype=torch.int32, device="cuda")
    lp = torch.empty(B, dtype=torch.float32, device="cuda")
    capi.check(lib.b2llm_sample_topk_topp(stream_ptr(), _ptr(dev(buf)), _ptr(dev(temps)) if temps is not None else None,
                                          _ptr(dev(top_p)) if top_p is not None else None,
                                          _ptr(dev(rand)) if rand is not None else None, B, V, stride, top_k,
                                          default_top_p, 0.25, _ptr(ws), _ptr(out), _ptr(lp)))
    sync()
    return out.cpu().numpy(), lp.cpu().numpy()


@pytest.mark.parametrize("V,stride", [(512, 512), (32000, 32000), (32000, 32064), (1000, 1024)])
def test_sampler_greedy(lib, V, stride):
    rng = np.random.default_rng(V)
    logits = rng.standard_normal((13, V)).astype(np.float32) * 2
    logits[3, 10] = logits[3, 400] = 50.0  # tie -> lowest index
    got, lp = _sample(lib, logits, None, None, None, 1, 0.0, stride)
    exp, elp = sampler_ref.sample_topk_topp(logits, None, None, None, V, 1, 0.0)
    assert np.array_equal(got, exp)  # token-for-token
    np.testing.assert_allclose(lp, elp, atol=2e-5)  # fp32 log-sum-exp order


@pytest.mark.parametrize("top_k", [2, 8, 50])
def test_sampler_topk_topp(lib, top_k):
    rng = np.random.default_rng(top_k)
    B, V = 16, 32000
    logits = rng.standard_normal((B, V)).astype(np.float32) * 3
    temps = rng.uniform(0.5, 1.5, B).astype(np.float32)
    top_p = rng.uniform(0.3, 1.0, B).astype(np.float32)
    top_p[0] = 0.0
    rand = rng.uniform(0, 1, B).astype(np.float32)
    rand[1] = 1.0
    got, lp = _sample(lib, logits, temps, top_p, rand, top_k, 0.9)
    exp, elp = sampler_ref.sample_topk_topp(logits, temps, top_p, rand, V, top_k, 0.9)
    assert np.array_equal(got, exp)
    np.testing.assert_allclose(lp, elp, atol=3e-5)
    # null per-request arrays -> defaults (the reference's non-changed-step quirk)
    got, lp = _sample(lib, logits, None, None, rand, top_k, 0.8)
    exp, elp = sampler_ref.sample_topk_topp(logits, None, None, rand, V, top_k, 0.8)
    assert np.array_equal(got, exp)


def test_apply_penalty(lib):
    """three consecutive steps over one count map: (1) four requests enter, two of them on a prefix-cache hit
    (start_pos > 0 on their first step, llm_generator.cc:229-242) into slots whose rows hold stale counts; (2) they all
    continue with one decode token; (3) two of them are replaced by new requests that REUSE their batch slots, one of
    which again enters at start_pos > 0 -- its row must be cleared although start_pos != 0."""
    rng = np.random.default_rng(2)
    B, V, slots = 4, 1000, 6
    temps = rng.uniform(0.5, 1.5, B).astype(np.float32)
    rep = rng.uniform(1.0, 1.5, B).astype(np.float32)
    pres = rng.uniform(0, 0.5, B).astype(np.float32)
    freq = rng.uniform(0, 0.5, B).astype(np.float32)
    batch_slots = np.array([3, 0, 5, 1], dtype=np.int64)
    cm = rng.integers(0, 3, (slots, V)).astype(np.uint16)   # stale counts: cudaMalloc'ed, never cleared (post_processor.cc:94-117)
    cmd = dev(cm.view(np.int16))
    ecm, next_pos = cm.copy(), {}
    steps = [
        (np.array([0, 1, 2, 9, 12]), np.array([17, 4, 0, 0])),    # prefix hits (1 token at 17, 1 at 4) + two fresh prompts
        (np.array([0, 1, 2, 3, 4]), np.array([18, 5, 7, 3])),     # everybody decodes
        (np.array([0, 1, 4, 5, 6]), np.array([19, 16, 8, 4])),    # slot 0 reused by a request entering at 16 (3 tokens)
    ]
    for seqstarts, start_pos in steps:
        seqstarts, start_pos = seqstarts.astype(np.int64), start_pos.astype(np.int64)
        tokens = rng.integers(0, V, int(seqstarts[-1])).astype(np.int64)
        if len(tokens) > 4:
            tokens[3] = tokens[4]  # duplicate inside a prompt
        logits = rng.standard_normal((B, V)).astype(np.float32)
        ld = dev(logits)
        capi.check(lib.b2llm_apply_penalty(stream_ptr(), _ptr(ld), _ptr(dev(temps)), _ptr(dev(rep)), _ptr(dev(pres)),
                                           _ptr(dev(freq)), _ptr(dev(batch_slots)), _ptr(dev(tokens)), _ptr(dev(seqstarts)),
                                           _ptr(dev(start_pos)), B, V, _ptr(cmd), _ptr(ld)))
        sync()
        el = logits.copy()
        sampler_ref.apply_penalty(el, temps, rep, pres, freq, batch_slots, tokens, seqstarts, start_pos, V, ecm, next_pos=next_pos)
        assert np.array_equal(cmd.cpu().numpy().view(np.uint16), ecm)  # counts bit-exact
        assert np.array_equal(ld.cpu().numpy(), el)  # every fp32 op individually rounded on both sides
    assert ecm[0].sum() == 3 and ecm[2].sum() == cm[2].sum()  # slot 0 restarted with 3 tokens; an unused slot is untouched
