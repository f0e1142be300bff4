"""numpy restatement of one forward step of ppl.llm.serving's engine over a ragged batch.

Test infrastructure (see ``oracle/__init__.py``); PARITY UNPINNED for the arithmetic.

What the reference pins (and this file follows):
  * the step descriptor ``ModelInput``                    src/engine/llm_engine.h:40-60
  * runtime inputs by index (token_ids, seq_starts, kv_starts, cache_indices, decoding_batches,
    start_pos, max_seq_len, max_kv_len, kv_cache, kv_scale) and the fp32 logits output
                                                           src/engine/llm_engine.h:124-138
  * the four KV-cache layouts                              src/engine/llm_engine.cc:118-169
  * cache_mode 0 (contiguous index) / 1 (page table, INT64_MAX padded)
                                                           src/engine/llm_engine.cc:60-72,
                                                           src/generator/llm_generator.cc:263-298
  * int8 KV with group 8 is the only quantised cache; cache_quant_bit 0 / group 1 = fp16 cache
    without a scale tensor                                 src/generator/llm_generator.cc:131-136,
                                                           src/backends/cuda/resource_manager.cc:381-388
  * sequences [0, decoding_batches) are in decode phase (one token each), the rest prefill
                                                           src/generator/llm_generator.cc:229-242,706-714
  * logits are fp32 [B, stride>=vocab], one row per sequence (last token)
                                                           src/engine/llm_engine.cc:200,219-224
What it restates from the published LLaMA-2 / ppl.pmx op definitions (Runtime::Run() is
EXTERNAL, src/engine/llm_engine.cc:113-116):
  pre-norm RMSNorm, rotate-half RoPE, online per-token int8 activation quantisation,
  per-output-channel int8 weights with exact int32 accumulation, SwiGLU, int8 group-8 KV.

Numeric conventions (the CUDA kernels implement the same cast points):
  residual stream and all inter-op activations are fp16; every reduction and every epilogue
  is fp32; int8 rounding is round-half-to-even, clamp [-127, 127]; activation scale =
  rowmax/127 (fp32); KV scale = fp16(groupmax/127); decode attention reads the *quantised*
  cache for every position including the current token, dequantised values being fp16
  (fp16(int8 * scale), the tensor-core operand type); prefill attention uses the fresh
  fp16 K/V for the new tokens (and the dequantised cache for a cached prefix).
"""
from __future__ import annotations

import numpy as np

from .weights import ModelDesc, SynthWeights

F32 = np.float32
INT64_MAX = np.iinfo(np.int64).max


# --------------------------------------------------------------------------- element ops
def rmsnorm_f32(x16: np.ndarray, gamma16: np.ndarray, eps: float) -> np.ndarray:
    """y = x * rsqrt(mean(x^2) + eps) * gamma, fp32 result (rows of [T, H])."""
    x = x16.astype(F32)
    var = (x.astype(np.float64) ** 2).mean(axis=-1).astype(F32)
    inv = (F32(1.0) / np.sqrt(var + F32(eps))).astype(F32)
    return (x * inv[:, None]) * gamma16.astype(F32)[None, :]


def quant_rows(y: np.ndarray):
    """per-token symmetric int8: scale = max|y|/127 (fp32), q = rint(y * (127/max|y|))."""
    y = y.astype(F32)
    amax = np.abs(y).max(axis=-1)
    scale = (amax / F32(127.0)).astype(F32)
    safe = np.where(amax > 0, amax, F32(1.0)).astype(F32)
    inv = np.where(amax > 0, F32(127.0) / safe, F32(0.0)).astype(F32)
    q = np.clip(np.rint(y * inv[:, None]), -127, 127).astype(np.int8)
    return q, scale


def gemm_i8_acc_numpy(a8: np.ndarray, w8: np.ndarray) -> np.ndarray:
    """exact int32 accumulation of int8 [M,K] x int8 [N,K]^T (float64 BLAS is exact: |acc| < 2^53)."""
    return (a8.astype(np.float64) @ w8.astype(np.float64).T).astype(np.int64).astype(np.int32)


def gemm_i8_acc(a8: np.ndarray, w8: np.ndarray) -> np.ndarray:
    """same result through the C restatement (oracle/csrc/oracle_kernels.c) when it is available --
    integer arithmetic, so the two are identical bit for bit (tests/test_oracle_cpu.py checks)."""
    from . import _native
    out = _native.gemm_i8_i32(a8, w8)
    return out if out is not None else gemm_i8_acc_numpy(a8, w8)


def dequant_acc(acc: np.ndarray, a_scale: np.ndarray, w_scale: np.ndarray) -> np.ndarray:
    """(float(acc) * a_scale[m]) * w_scale[n], two fp32 multiplies in this order."""
    return (acc.astype(F32) * a_scale.astype(F32)[:, None]) * w_scale.astype(F32)[None, :]


def gemm_f16_acc(a16: np.ndarray, w16: np.ndarray) -> np.ndarray:
    return a16.astype(F32) @ w16.astype(F32).T


def silu_mul(g: np.ndarray, u: np.ndarray) -> np.ndarray:
    g = g.astype(F32)
    return (g / (F32(1.0) + np.exp(-g).astype(F32))).astype(F32) * u.astype(F32)


def rope_table(max_pos: int, head_dim: int, theta: float):
    """cos/sin fp32 [max_pos, head_dim/2], angles computed in float64."""
    i = np.arange(head_dim // 2, dtype=np.float64)
    inv_freq = np.power(np.float64(theta), -2.0 * i / head_dim)
    ang = np.arange(max_pos, dtype=np.float64)[:, None] * inv_freq[None, :]
    return np.cos(ang).astype(F32), np.sin(ang).astype(F32)


def apply_rope(x16: np.ndarray, pos: np.ndarray, cos: np.ndarray, sin: np.ndarray) -> np.ndarray:
    """rotate-half pairing (i, i + D/2) on [T, H, D] fp16; each product rounded to fp32 (no fma)."""
    D = x16.shape[-1]
    h = D // 2
    x = x16.astype(F32)
    x1, x2 = x[..., :h], x[..., h:]
    c = cos[pos][:, None, :]
    s = sin[pos][:, None, :]
    o1 = (x1 * c).astype(F32) - (x2 * s).astype(F32)
    o2 = (x2 * c).astype(F32) + (x1 * s).astype(F32)
    return np.concatenate([o1, o2], axis=-1).astype(np.float16)


def kv_quant(x16: np.ndarray, group: int = 8):
    """int8 group quantisation along the last axis with an fp16 scale per group.

    scale16 = fp16(max|x|/127); q = clamp(rint(x / fp32(scale16)), -127, 127) (0 if scale16 == 0).
    """
    shp = x16.shape
    x = x16.astype(F32).reshape(shp[:-1] + (shp[-1] // group, group))
    amax = np.abs(x).max(axis=-1)
    s16 = (amax / F32(127.0)).astype(F32).astype(np.float16)
    s = s16.astype(F32)
    safe = np.where(s > 0, s, F32(1.0))
    q = np.where(s[..., None] > 0, np.rint(x / safe[..., None]), 0.0)
    q = np.clip(q, -127, 127).astype(np.int8).reshape(shp)
    return q, s16


def kv_dequant(q8: np.ndarray, s16: np.ndarray, group: int = 8) -> np.ndarray:
    """dequantised cache values are fp16 (the tensor-core operand type): fp16(int8 * scale), as fp32."""
    shp = q8.shape
    x = q8.astype(F32).reshape(shp[:-1] + (shp[-1] // group, group))
    return (x * s16.astype(F32)[..., None]).astype(np.float16).astype(F32).reshape(shp)


# --------------------------------------------------------------------------- KV cache
class KVCache:
    """int8 cache + fp16 scale (``cache_quant_bit`` 8, group 8) or a plain fp16 cache with no scale tensor
    (``cache_quant_bit`` 0, group 1: llm_generator.cc:131-136; runtime input 10 is then not bound,
    llm_engine.h:134-136).  Held canonically as [L, 2, T, H, D]; ``export()`` / ``load()`` convert to and from the
    reference's four layouts (llm_engine.cc:118-169):

    layout 0: [T, L, 2, H, D]   1: [L, T, 2, H, D]   2: [L, 2, T, H, D]   3: [L, 2, H, T, D]
    """

    _AXES = {0: (2, 0, 1, 3, 4), 1: (0, 2, 1, 3, 4), 2: (0, 1, 2, 3, 4), 3: (0, 1, 3, 2, 4)}

    def __init__(self, desc: ModelDesc, max_tokens: int):
        self.desc = desc
        self.T = max_tokens
        L, H, D = desc.num_layers, desc.num_kv_heads, desc.head_dim
        self.fp16 = desc.cache_quant_bit == 0
        G = 0 if self.fp16 else D // desc.cache_quant_group
        self.cache = np.zeros((L, 2, max_tokens, H, D), dtype=np.float16 if self.fp16 else np.int8)
        self.scale = np.zeros((L, 2, max_tokens, H, G), dtype=np.float16)

    def write(self, layer, slots, k8, ks, v8, vs):
        slots = np.asarray(slots, dtype=np.int64)
        self.cache[layer, 0, slots] = k8
        self.cache[layer, 1, slots] = v8
        if not self.fp16:
            self.scale[layer, 0, slots] = ks
            self.scale[layer, 1, slots] = vs

    def append(self, layer, slots, k16, v16):
        """store this step's (rotated) K and V rows: quantised per group of 8, or as they are (fp16 cache)"""
        if self.fp16:
            self.write(layer, slots, k16, None, v16, None)
            return None
        g = self.desc.cache_quant_group
        k8, ks = kv_quant(k16, g)
        v8, vs = kv_quant(v16, g)
        self.write(layer, slots, k8, ks, v8, vs)
        return k8, ks, v8, vs

    def read_values(self, layer, kv, slots):
        """fp32 values attention sees: fp16(int8 * scale) for the int8 cache, the stored fp16 otherwise"""
        slots = np.asarray(slots, dtype=np.int64)
        if self.fp16:
            return self.cache[layer, kv, slots].astype(F32)
        return kv_dequant(self.cache[layer, kv, slots], self.scale[layer, kv, slots], self.desc.cache_quant_group)

    def read(self, layer, kv, slots):
        slots = np.asarray(slots, dtype=np.int64)
        return self.cache[layer, kv, slots], self.scale[layer, kv, slots]

    def export(self):
        """(cache, scale) as contiguous arrays in ``desc.cache_layout``."""
        ax = self._AXES[self.desc.cache_layout]
        return np.ascontiguousarray(self.cache.transpose(ax)), np.ascontiguousarray(self.scale.transpose(ax))

    def load(self, cache_bytes: np.ndarray, scale_bytes: np.ndarray):
        ax = self._AXES[self.desc.cache_layout]
        inv = np.argsort(ax)
        shp_c = tuple(np.array(self.cache.shape)[list(ax)])
        shp_s = tuple(np.array(self.scale.shape)[list(ax)])
        self.cache = np.ascontiguousarray(cache_bytes.reshape(shp_c).transpose(inv))
        self.scale = np.ascontiguousarray(scale_bytes.reshape(shp_s).transpose(inv))


class Step:
    """Mirror of ``ppl::llm::ModelInput`` (llm_engine.h:40-60), numpy int64 arrays."""

    def __init__(self, token_inputs, seq_starts, kv_starts, start_pos, decoding_batches,
                 cache_indices=None, page_list=None, max_pages=0):
        self.token_inputs = np.asarray(token_inputs, dtype=np.int64)
        self.seq_starts = np.asarray(seq_starts, dtype=np.int64)
        self.kv_starts = np.asarray(kv_starts, dtype=np.int64)
        self.start_pos = np.asarray(start_pos, dtype=np.int64)
        self.decoding_batches = int(decoding_batches)
        self.cache_indices = None if cache_indices is None else np.asarray(cache_indices, dtype=np.int64)
        self.page_list = None if page_list is None else np.asarray(page_list, dtype=np.int64)
        self.max_pages = int(max_pages)
        seqlens = np.diff(self.seq_starts)
        self.max_seq_len = int(seqlens.max()) if len(seqlens) else 0
        self.max_kv_len = int((self.start_pos + seqlens).max()) if len(seqlens) else 0

    @property
    def batch(self):
        return len(self.start_pos)

    def slots(self, desc: ModelDesc, b: int, positions: np.ndarray) -> np.ndarray:
        """token slot in the cache for positions of sequence b.

        cache_mode 0: cache_indices[b] + p (llm_engine.cc:61-63);
        cache_mode 1: page_list[b, p // page_size] + p % page_size, each entry being the first token
        slot of a page (PageManager is EXTERNAL; entries are page *begin indices* as in ppl.pmx's
        ``cache_starts``; llm_generator.cc:278-296 only copies them).
        """
        positions = np.asarray(positions, dtype=np.int64)
        if desc.cache_mode == 0:
            return self.cache_indices[b] + positions
        ps = desc.page_size
        pages = self.page_list[b * self.max_pages + positions // ps]
        assert (pages != INT64_MAX).all(), "page table padding hit"
        return pages + positions % ps


# --------------------------------------------------------------------------- attention
def attention_decode(q16, cache: KVCache, layer, slots, group):
    """one query token (all heads) against kv_len cached tokens. q16 [Hq, D] -> fp32 [Hq, D]."""
    Hq, D = q16.shape
    K = cache.read_values(layer, 0, slots)  # [t, Hkv, D]
    V = cache.read_values(layer, 1, slots)
    return _attend(q16.astype(F32)[None], K, V, np.array([K.shape[0] - 1]))[0]


def _attend(q, K, V, last_visible):
    """q [n, Hq, D] fp32, K/V [t, Hkv, D] fp32; query i sees keys [0, last_visible[i]]."""
    n, Hq, D = q.shape
    t, Hkv, _ = K.shape
    G = Hq // Hkv
    qg = q.reshape(n, Hkv, G, D)
    s = np.einsum("nhgd,thd->nhgt", qg, K, optimize=True).astype(F32) * F32(1.0 / np.sqrt(D))
    mask = np.arange(t)[None, :] > np.asarray(last_visible)[:, None]
    s = np.where(mask[:, None, None, :], F32(-np.inf), s)
    m = s.max(axis=-1, keepdims=True)
    e = np.exp(s - m).astype(F32)
    p = e / e.sum(axis=-1, keepdims=True)
    o = np.einsum("nhgt,thd->nhgd", p, V, optimize=True).astype(F32)
    return o.reshape(n, Hq, D)


# --------------------------------------------------------------------------- the step
class LlamaOracle:
    """``tp`` > 1 restates the reference's tensor parallelism (``--tensor-parallel-size``; one model slice per rank,
    ``resource_manager.cc:280-286``; ``num_kv_heads / tp`` per rank, ``llm_engine.cc:124``): q/k/v/gate/up are split
    by output channel (results identical to tp = 1: per-channel weight scales, same quantised input), o_proj and
    down_proj by input channel -- each rank quantises ITS slice of the activation row and ITS slice of the weight
    (scale = slice max / 127), stores its partial product as fp16 and the partials are summed (the all-reduce at the
    "two residual join points").  ``allreduce``: None = simulate all ranks in this process; otherwise a callable
    ``fp16 [T, h] partial of `rank` -> fp16 sum over ranks`` (tests/test_tp_gloo.py runs one process per rank)."""

    def __init__(self, desc: ModelDesc, weights: SynthWeights, kv_max_tokens: int, tp: int = 1, rank: int | None = None,
                 allreduce=None):
        self.desc = desc
        self.w = weights
        self.cache = KVCache(desc, kv_max_tokens)
        self.cos, self.sin = rope_table(desc.max_position, desc.head_dim, desc.rope_theta)
        self.tp, self.rank, self.allreduce = tp, rank, allreduce
        self._slice_cache = {}

    def _row_parallel_partial(self, r: int, x16: np.ndarray, lw, name: str, layer: int) -> np.ndarray:
        """rank r's fp16 partial of a row-parallel projection: columns [r*K/tp, (r+1)*K/tp) of x and W"""
        d = self.desc
        K = x16.shape[1]
        lo, hi = r * K // self.tp, (r + 1) * K // self.tp
        xs = x16[:, lo:hi]
        if d.quant_method == 1:
            key = (layer, name, r)
            if key not in self._slice_cache:
                from .weights import quantize_weight_per_channel
                self._slice_cache[key] = quantize_weight_per_channel(lw[name][:, lo:hi])
            wq, ws = self._slice_cache[key]
            q, s = quant_rows(xs.astype(F32))
            return dequant_acc(gemm_i8_acc(q, wq), s, ws).astype(np.float16)
        if d.quant_method == 2:  # groups of 128 run along the rank's own K slice
            key = (layer, name, r)
            if key not in self._slice_cache:
                from .weights import quantize_weight_w4
                self._slice_cache[key] = quantize_weight_w4(np.ascontiguousarray(lw[name][:, lo:hi]))[2]
            return gemm_f16_acc(xs, self._slice_cache[key]).astype(np.float16)
        return gemm_f16_acc(xs, lw[name][:, lo:hi]).astype(np.float16)

    def _row_parallel(self, x16: np.ndarray, lw, name: str, layer: int) -> np.ndarray:
        """fp32 [T, h]: the projection output that is added to the residual stream"""
        if self.tp == 1:
            return self._linear(x16.astype(F32) if self.desc.quant_method == 1 else x16, lw, name)
        if self.allreduce is not None:
            return self.allreduce(self._row_parallel_partial(self.rank, x16, lw, name, layer)).astype(F32)
        total = np.zeros((x16.shape[0], lw[name].shape[0]), dtype=F32)
        for r in range(self.tp):
            total += self._row_parallel_partial(r, x16, lw, name, layer).astype(F32)
        return total.astype(np.float16).astype(F32)

    # linear layer in the model's quant mode: returns fp32 pre-rounding result
    def _linear(self, x_f32_or_16, lw, name, prequant=None):
        d = self.desc
        if d.quant_method == 1:
            q, s = prequant if prequant is not None else quant_rows(x_f32_or_16)
            acc = gemm_i8_acc(q, lw[name + "_q"])
            return dequant_acc(acc, s, lw[name + "_s"])
        if d.quant_method == 2:  # W4A16: fp16 activations x fp16(q * scale) weights, fp32 accumulate
            return gemm_f16_acc(x_f32_or_16.astype(np.float16), lw[name + "_w4"])
        return gemm_f16_acc(x_f32_or_16.astype(np.float16), lw[name])

    def forward(self, step: Step, trace: dict | None = None, ulp_nudge: bool = False) -> np.ndarray:
        """returns fp32 logits [B, vocab] (row b = last token of sequence b).

        ``ulp_nudge``: move the largest-magnitude element of every attention-output row by one fp16
        ulp before it is re-quantised.  Used by the tests to measure how far the logits of this
        algorithm move under the smallest representable perturbation (the noise floor any two
        implementations with different fp32 summation orders are subject to)."""
        d = self.desc
        D, Hq, Hkv = d.head_dim, d.num_heads, d.num_kv_heads
        T = len(step.token_inputs)
        B = step.batch
        seqlens = np.diff(step.seq_starts)
        # position of every token
        pos = np.concatenate([step.start_pos[b] + np.arange(seqlens[b]) for b in range(B)]) if T else np.zeros(0, np.int64)
        slots = np.concatenate([step.slots(d, b, step.start_pos[b] + np.arange(seqlens[b])) for b in range(B)])

        x = self.w.embedding()[step.token_inputs]  # fp16 [T, h]
        for l in range(d.num_layers):
            lw = self.w.layer(l)
            # ---- attention block
            y = rmsnorm_f32(x, lw["attn_norm"], d.norm_eps)
            if d.quant_method == 1:
                pq = quant_rows(y)
                if trace is not None and l == 0:
                    trace["l0_attn_in_q"], trace["l0_attn_in_s"] = pq
                qkv = self._linear(None, lw, "wqkv", prequant=pq).astype(np.float16)
            else:
                qkv = self._linear(y.astype(np.float16), lw, "wqkv").astype(np.float16)
            if trace is not None and l == 0:
                trace["l0_qkv"] = qkv.copy()
            q = qkv[:, : Hq * D].reshape(T, Hq, D)
            k = qkv[:, Hq * D: (Hq + Hkv) * D].reshape(T, Hkv, D)
            v = qkv[:, (Hq + Hkv) * D:].reshape(T, Hkv, D)
            q = apply_rope(q, pos, self.cos, self.sin)
            k = apply_rope(k, pos, self.cos, self.sin)
            appended = self.cache.append(l, slots, k, v)
            if trace is not None and l == 0:
                trace["l0_q_rot"], trace["l0_k_rot"] = q.copy(), k.copy()
                if appended is not None:
                    trace["l0_k8"], trace["l0_ks"], trace["l0_v8"], trace["l0_vs"] = appended

            attn = np.empty((T, Hq, D), dtype=F32)
            for b in range(B):
                t0, t1 = step.seq_starts[b], step.seq_starts[b + 1]
                sp, n = int(step.start_pos[b]), int(t1 - t0)
                if b < step.decoding_batches:
                    assert n == 1
                    sl = step.slots(d, b, np.arange(sp + 1))
                    attn[t0] = attention_decode(q[t0], self.cache, l, sl, d.cache_quant_group)
                else:
                    Kf, Vf = k[t0:t1].astype(F32), v[t0:t1].astype(F32)
                    if sp > 0:  # cached prefix (prefix-cache hit): dequantised cache for [0, sp)
                        sl = step.slots(d, b, np.arange(sp))
                        Kf = np.concatenate([self.cache.read_values(l, 0, sl), Kf])
                        Vf = np.concatenate([self.cache.read_values(l, 1, sl), Vf])
                    attn[t0:t1] = _attend(q[t0:t1].astype(F32), Kf, Vf, sp + np.arange(n))
            attn16 = attn.reshape(T, Hq * D).astype(np.float16)
            if ulp_nudge:
                j = np.abs(attn16.astype(F32)).argmax(axis=1)
                r = np.arange(T)
                attn16[r, j] = np.nextafter(attn16[r, j], np.float16(0))
            if trace is not None and l == 0:
                trace["l0_attn"] = attn16.copy()
            o = self._row_parallel(attn16, lw, "wo", l)
            x = (x.astype(F32) + o).astype(np.float16)
            if trace is not None and l == 0:
                trace["l0_x_mid"] = x.copy()

            # ---- feed-forward block
            y = rmsnorm_f32(x, lw["ffn_norm"], d.norm_eps)
            if d.quant_method == 1:
                pq = quant_rows(y)
                g = self._linear(None, lw, "wgate", prequant=pq)
                u = self._linear(None, lw, "wup", prequant=pq)
            else:
                y16 = y.astype(np.float16)
                g = self._linear(y16, lw, "wgate")
                u = self._linear(y16, lw, "wup")
            act = silu_mul(g, u).astype(np.float16)
            if trace is not None and l == 0:
                trace["l0_act"] = act.copy()
            dn = self._row_parallel(act, lw, "wdown", l)
            x = (x.astype(F32) + dn).astype(np.float16)
            if trace is not None and l == 0:
                trace["l0_x_out"] = x.copy()

        last = step.seq_starts[1:] - 1
        xl = x[last]
        yl = rmsnorm_f32(xl, self.w.final_norm(), d.norm_eps).astype(np.float16)
        logits = gemm_f16_acc(yl, self.w.lm_head())
        if trace is not None:
            trace["x_final"] = x.copy()
            trace["y_last"] = yl.copy()
        return logits.astype(F32)


# --------------------------------------------------------------------------- helpers for tests / bench
def build_step(desc: ModelDesc, seqs_tokens, start_pos, decoding_batches, page_tables=None, cache_indices=None):
    """assemble a Step the way UpdateInput does (llm_generator.cc:263-298)."""
    token_inputs, seq_starts, kv_starts = [], [0], [0]
    for toks, sp in zip(seqs_tokens, start_pos):
        token_inputs.extend(toks)
        seq_starts.append(seq_starts[-1] + len(toks))
        kv_starts.append(kv_starts[-1] + sp + len(toks))
    page_list, max_pages = None, 0
    if desc.cache_mode == 1:
        max_pages = max(len(p) for p in page_tables)
        page_list = np.full(len(page_tables) * max_pages, INT64_MAX, dtype=np.int64)
        for i, p in enumerate(page_tables):
            page_list[i * max_pages: i * max_pages + len(p)] = p
    return Step(token_inputs, seq_starts, kv_starts, start_pos, decoding_batches,
                cache_indices=cache_indices, page_list=page_list, max_pages=max_pages)
