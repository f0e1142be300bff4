"""Integer and layout logic of the transposed W4A16 GEMM (csrc/gemm_tcgen05.cu: gemm_w4t_kernel), restated on the CPU.

1. `w4_expand8`: four `lop3` of the form (w >> 4 j) & 0x000F000F | 0x64006400 give the half2 pairs (1024 + e_j, 1024 + e_{j+4});
   subtracting 1032 is exact, the product with the fp16 scale rounds once, and four PRMTs (0x5410 / 0x7632) put the pairs
   back in k order.  The result must be the oracle's dequantised operand fp16(q * s) (oracle/weights.py: quantize_weight_w4),
   bit for bit, for every nibble value and a spread of scales.
2. Tensor-memory operand addressing: converter thread (channel r, half hf) of a k-block reads bytes [16 hf, 16 hf + 16) of the
   row's 32 packed bytes and stores 16 columns [32 stage + 16 hf, + 16) of its own lane; the MMA of k-step k reads columns
   [32 stage + 8 k, + 8) of every lane as its A operand (two fp16 per 32-bit column, even k in the low half: the layout of
   cute's tmem_frg for a 16-bit A with M = 128).  What the MMA sees must be W[r, 64 kb + 16 k + kk].
3. The epilogue's lane-pair exchange: accumulator lanes are output channels, columns are activation rows; even lanes end up
   with row jj of channels (n, n + 1), odd lanes with row jj + 1 of (n - 1, n).  Every (row, channel) of a 32 x 32 chunk must
   be stored exactly once with its own value; the SwiGLU variant must pair (gate, up) = channels (2 i, 2 i + 1).
(The tcgen05 semantics themselves -- that an MMA reads its A operand from those columns -- are confirmed on the device by
tests/test_ops_gpu.py::test_gemm_w4a16_fused; a slip in an offset or a selector shows up here first.)"""
import numpy as np

from oracle.weights import quantize_weight_w4

U32 = np.uint32


def lop3_and_or(x, m, k):
    return U32((int(x) & m) | k)


def prmt(a, b, sel):
    """PRMT.b32 in its default mode: result byte i = byte (sel >> 4 i) & 7 of the 8-byte pool {a (bytes 0-3), b (bytes 4-7)}"""
    pool = [(int(a) >> (8 * i)) & 0xFF for i in range(4)] + [(int(b) >> (8 * i)) & 0xFF for i in range(4)]
    return U32(sum(pool[(sel >> (4 * i)) & 7] << (8 * i) for i in range(4)))


def half2_of(word):
    return np.array([int(word) & 0xFFFF, int(word) >> 16], dtype=np.uint16).view(np.float16)


def word_of(h2):
    u = np.asarray(h2, dtype=np.float16).view(np.uint16)
    return U32(int(u[0]) | (int(u[1]) << 16))


def w4_expand8(w, scale16):
    """the device function, instruction for instruction: 8 nibbles -> 4 words of (e0, e1) (e2, e3) (e4, e5) (e6, e7)"""
    bias = np.float16(1032.0)
    p = []
    for j in range(4):
        pk = lop3_and_or(int(w) >> (4 * j), 0x000F000F, 0x64006400)
        h = half2_of(pk)
        h = ((h - bias).astype(np.float16) * np.float16(scale16)).astype(np.float16)   # hsub2 (exact), hmul2 (one rounding)
        p.append(word_of(h))
    return [prmt(p[0], p[1], 0x5410), prmt(p[2], p[3], 0x5410), prmt(p[0], p[1], 0x7632), prmt(p[2], p[3], 0x7632)]


def pack_rows(q):
    """b2llm_op_quant_weight_w4's format: byte j of a row = (q[2 j] + 8) | (q[2 j + 1] + 8) << 4"""
    u = (q.astype(np.int16) + 8).astype(np.uint8)
    return (u[:, 0::2] | (u[:, 1::2] << 4)).astype(np.uint8)


def test_expand8_equals_oracle_dequantisation_bit_for_bit():
    rng = np.random.default_rng(0)
    scales = np.concatenate([rng.uniform(1e-4, 0.05, 60), [0.0, 6.1e-5, 1.0, 65504 / 8]]).astype(np.float16)
    for s16 in scales:
        q = rng.integers(-7, 8, (1, 128)).astype(np.int8)
        q[0, :15] = np.arange(-7, 8)                                  # every code at least once
        deq = (q.astype(np.float32) * np.float32(s16)).astype(np.float16)   # the oracle's operand definition
        words = pack_rows(q).view(np.uint32)[0]
        got = np.concatenate([np.concatenate([half2_of(x) for x in w4_expand8(w, s16)]) for w in words])
        assert np.array_equal(got.view(np.uint16), deq[0].view(np.uint16)), float(s16)


def test_oracle_quantiser_and_expansion_agree_on_real_weights():
    rng = np.random.default_rng(1)
    w = (0.02 * rng.standard_normal((8, 256))).astype(np.float16)
    q, s16, deq = quantize_weight_w4(w)
    packed = pack_rows(q)
    for r in range(8):
        words = packed[r].view(np.uint32)
        got = np.concatenate([np.concatenate([half2_of(x) for x in w4_expand8(wd, s16[r, (8 * i) // 128])])
                              for i, wd in enumerate(words)])
        assert np.array_equal(got.view(np.uint16), deq[r].view(np.uint16))


def test_tmem_operand_columns_seen_by_each_mma_step():
    NA, STAGES, A_COLS = 256, 6, 32
    rng = np.random.default_rng(2)
    K = 64 * 7
    q = rng.integers(-7, 8, (128, K)).astype(np.int8)
    s16 = rng.uniform(1e-3, 0.02, (128, (K + 127) // 128)).astype(np.float16)
    packed = pack_rows(q)                                              # [128, K / 2]
    deq = (q.astype(np.float32) * np.repeat(s16.astype(np.float32), 128, axis=1)[:, :K]).astype(np.float16)
    tmem = np.zeros((128, 512), dtype=np.uint32)
    for kb in range(K // 64):
        stage = kb % STAGES
        for r in range(128):
            for hf in range(2):                                        # converter thread (r, hf)
                raw = packed[r, kb * 32 + 16 * hf: kb * 32 + 16 * hf + 16].view(np.uint32)
                v = [x for wd in raw for x in w4_expand8(wd, s16[r, kb >> 1])]
                col0 = NA + stage * A_COLS + hf * 16                   # tcgen05.st.32x32b.x16 at the thread's own lane
                tmem[r, col0:col0 + 16] = v
        for k in range(4):                                             # tcgen05.mma [d], [tmem_w + 32 stage + 8 k], ...
            cols = tmem[:, NA + stage * A_COLS + 8 * k: NA + stage * A_COLS + 8 * k + 8]
            seen = np.ascontiguousarray(cols).view(np.uint16).reshape(128, 16)           # element kk = half kk % 2 of column kk / 2
            want = deq[:, 64 * kb + 16 * k: 64 * kb + 16 * k + 16].view(np.uint16)
            assert np.array_equal(seen, want), (kb, k)


def _exchange(rr, odd):
    """per lane: a[p], b[p] = values of channels (n & ~1), (n & ~1) + 1 in row 2 p + odd, after one shuffle with lane ^ 1"""
    lanes = rr.shape[0]
    a, b = np.zeros((lanes, 16)), np.zeros((lanes, 16))
    for p in range(16):
        send = np.where(odd, rr[:, 2 * p], rr[:, 2 * p + 1])
        got = send[np.arange(lanes) ^ 1]
        a[:, p] = np.where(odd, got, rr[:, 2 * p])
        b[:, p] = np.where(odd, rr[:, 2 * p + 1], got)
    return a, b


def test_epilogue_lane_pair_exchange_covers_every_element_once():
    rng = np.random.default_rng(3)
    D = rng.standard_normal((32, 32))                                  # D[row jj, channel lane] of one 32 x 32 chunk
    rr = D.T.copy()                                                    # lane's registers: rr[lane][jj] = accumulator column jj
    lane = np.arange(32)
    odd = (lane & 1).astype(bool)
    a, b = _exchange(rr, odd)
    out = np.full((32, 32), np.nan)
    writes = np.zeros((32, 32), dtype=int)
    for ln in range(32):
        for p in range(16):
            row, ch = 2 * p + (ln & 1), ln & ~1                        # dst = out + (m0 + odd) * ldc + (n & ~1), half2 store, p * 2 rows on
            out[row, ch], out[row, ch + 1] = a[ln, p], b[ln, p]
            writes[row, ch] += 1
            writes[row, ch + 1] += 1
    assert (writes == 1).all()
    assert np.array_equal(out, D)
    # SwiGLU: one output column per (gate, up) channel pair, row 2 p + odd of the lane
    sw = np.full((32, 16), np.nan)
    wcount = np.zeros((32, 16), dtype=int)
    for ln in range(32):
        for p in range(16):
            g, u = a[ln, p], b[ln, p]
            sw[2 * p + (ln & 1), ln >> 1] = g / (1.0 + np.exp(-g)) * u
            wcount[2 * p + (ln & 1), ln >> 1] += 1
    assert (wcount == 1).all()
    np.testing.assert_allclose(sw, D[:, 0::2] / (1.0 + np.exp(-D[:, 0::2])) * D[:, 1::2], rtol=1e-12)
