"""Seeded synthetic LLaMA weights, reproducible bit-for-bit on CPU (numpy) and on
the device (``csrc/weights.cu`` runs the same integer hash).

Test infrastructure (see ``oracle/__init__.py``).  The reference ships no weights;
its model slices come from a ppl.pmx export (``docs/llama_guide.md:14-36``,
``src/backends/cuda/resource_manager.cc:280-290``).  Synthetic weights are what
``BASELINE.json`` asks for.

Generator: element ``i`` of tensor ``tid`` is
    z  = splitmix64(seed * 0x9E3779B97F4A7C15 + tid * 0xD1B54A32D192ED03 + i)
    s  = sum of the four 16-bit limbs of z            (Irwin-Hall, ~normal)
    w  = fp16( fp32(s - 131070) * fp32(std / 37836.6...) + mean )
Everything up to the single fp32 multiply-add is integer arithmetic, so numpy
and CUDA agree exactly.
"""
from __future__ import annotations

import numpy as np

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
IH_STD = 65535.0 / np.sqrt(3.0)  # std of a sum of four U{0..65535}

# tensor ids (shared with csrc/weights.cu)
TID_EMBED = 1
TID_FINAL_NORM = 2
TID_LM_HEAD = 3
TID_LAYER_BASE = 16  # + layer * 8 + kind
K_ATTN_NORM, K_WQKV, K_WO, K_FFN_NORM, K_WGATE, K_WUP, K_WDOWN = range(7)


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & M64
        return z ^ (z >> np.uint64(31))


def synth_tensor(seed: int, tid: int, shape, std: float, mean: float = 0.0) -> np.ndarray:
    """fp16 tensor; must match ``synth_fp16_kernel`` in csrc/weights.cu."""
    n = int(np.prod(shape))
    with np.errstate(over="ignore"):
        base = (np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
                + np.uint64(tid) * np.uint64(0xD1B54A32D192ED03)) & M64
        out = np.empty(n, dtype=np.float16)
        step = 1 << 22
        mul = np.float32(np.float64(np.float32(std)) / IH_STD)  # std crosses the C ABI as a float
        for lo in range(0, n, step):
            hi = min(n, lo + step)
            idx = np.arange(lo, hi, dtype=np.uint64)
            z = splitmix64((base + idx) & M64)
            s = ((z & np.uint64(0xFFFF)) + ((z >> np.uint64(16)) & np.uint64(0xFFFF))
                 + ((z >> np.uint64(32)) & np.uint64(0xFFFF)) + (z >> np.uint64(48))).astype(np.int64)
            f = (s - 131070).astype(np.float32)
            # one fp32 multiply, one fp32 add (no fma on either side: mean is 0 or 1 and the
            # CUDA side uses __fmul_rn/__fadd_rn)
            out[lo:hi] = ((f * mul) + np.float32(mean)).astype(np.float16)
    return out.reshape(shape)


def quantize_weight_per_channel(w16: np.ndarray):
    """online_i8i8 weight quantisation: one scale per output channel (row of [N, K]).

    ``quant_method == "online_i8i8"`` at ``resource_manager.cc:51-52``; the pass itself is
    inside ppl.nn [EXTERNAL].  Convention fixed here: scale = max|w| / 127 in fp32,
    q = clamp(rint(w * (127 / max|w|)), -127, 127).
    """
    w = w16.astype(np.float32)
    amax = np.abs(w).max(axis=1)
    scale = (amax / np.float32(127.0)).astype(np.float32)
    inv = np.where(amax > 0, np.float32(127.0) / np.where(amax > 0, amax, 1), 0).astype(np.float32)
    q = np.clip(np.rint(w * inv[:, None]), -127, 127).astype(np.int8)
    return q, scale


def quantize_weight_w4(w16: np.ndarray, group: int = 128):
    """W4A16 (builder-defined -- the reference rejects every quant method but "none" / "online_i8i8",
    ``resource_manager.cc:49-56``): symmetric int4, one fp16 scale per ``group`` consecutive K elements of an output
    channel.  scale16 = fp16(max|w| / 7); q = clamp(rint(w / fp32(scale16)), -7, 7) (0 where scale16 == 0).
    Returns (q int8 [N, K], scale fp16 [N, K/group], dequantised operand fp16 [N, K] = fp16(q * scale))."""
    N, K = w16.shape
    w = w16.astype(np.float32).reshape(N, K // group, group)
    amax = np.abs(w).max(axis=-1)
    s16 = (amax / np.float32(7.0)).astype(np.float32).astype(np.float16)
    s = s16.astype(np.float32)
    safe = np.where(s > 0, s, np.float32(1.0))
    q = np.where(s[..., None] > 0, np.rint(w / safe[..., None]), 0.0)
    q = np.clip(q, -7, 7).astype(np.int8)
    deq = (q.astype(np.float32) * s[..., None]).astype(np.float16).reshape(N, K)
    return q.reshape(N, K), s16, deq


class ModelDesc:
    """Mirror of ``ppl::llm::ModelConfig`` (``src/common/config.h:64-84``) plus the knobs that live
    in the exported graph rather than params.json (eps, rope theta; ``config.h:74``)."""

    def __init__(self, hidden_dim, intermediate_dim, num_layers, num_heads, num_kv_heads, vocab_size,
                 norm_eps=1e-5, rope_theta=10000.0, cache_quant_bit=8, cache_quant_group=8,
                 cache_layout=3, cache_mode=1, page_size=16, quant_method=1, max_position=4096):
        self.hidden_dim = hidden_dim
        self.intermediate_dim = intermediate_dim
        self.num_layers = num_layers
        self.num_heads = num_heads
        self.num_kv_heads = num_kv_heads
        self.vocab_size = vocab_size
        self.norm_eps = norm_eps
        self.rope_theta = rope_theta
        self.cache_quant_bit = cache_quant_bit
        self.cache_quant_group = cache_quant_group
        self.cache_layout = cache_layout
        self.cache_mode = cache_mode
        self.page_size = page_size
        self.quant_method = quant_method  # 0 none (fp16), 1 online_i8i8, 2 w4a16 (builder-defined)
        self.max_position = max_position

    @property
    def head_dim(self):
        return self.hidden_dim // self.num_heads

    def kv_bytes_per_token(self):
        """cb + sb of ``resource_manager.cc:381-388`` (tp = 1)."""
        d = self.head_dim
        cb = self.num_layers * 2 * self.num_kv_heads * d * (1 if self.cache_quant_bit == 8 else 2)
        sb = self.num_layers * 2 * self.num_kv_heads * d // self.cache_quant_group * 2 \
            if self.cache_quant_bit > 0 else 0
        return cb, sb


# std per tensor kind (mean 1 for the norm gains)
STD_EMBED = 1.0
STD_W = 0.02
STD_NORM = 0.02


class SynthWeights:
    """All weights of a synthetic LLaMA, generated lazily per layer."""

    def __init__(self, desc: ModelDesc, seed: int = 0xB200):
        self.desc = desc
        self.seed = seed
        self._cache = {}

    def _get(self, key, fn):
        if key not in self._cache:
            self._cache[key] = fn()
        return self._cache[key]

    def embedding(self):
        d = self.desc
        return self._get("emb", lambda: synth_tensor(self.seed, TID_EMBED, (d.vocab_size, d.hidden_dim), STD_EMBED))

    def final_norm(self):
        d = self.desc
        return self._get("fn", lambda: synth_tensor(self.seed, TID_FINAL_NORM, (d.hidden_dim,), STD_NORM, 1.0))

    def lm_head(self):
        d = self.desc
        return self._get("lm", lambda: synth_tensor(self.seed, TID_LM_HEAD, (d.vocab_size, d.hidden_dim), STD_W))

    def layer(self, l: int):
        d = self.desc
        D = d.head_dim
        nqkv = (d.num_heads + 2 * d.num_kv_heads) * D

        def make():
            t = lambda k: TID_LAYER_BASE + l * 8 + k
            w = {
                "attn_norm": synth_tensor(self.seed, t(K_ATTN_NORM), (d.hidden_dim,), STD_NORM, 1.0),
                "wqkv": synth_tensor(self.seed, t(K_WQKV), (nqkv, d.hidden_dim), STD_W),
                "wo": synth_tensor(self.seed, t(K_WO), (d.hidden_dim, d.num_heads * D), STD_W),
                "ffn_norm": synth_tensor(self.seed, t(K_FFN_NORM), (d.hidden_dim,), STD_NORM, 1.0),
                "wgate": synth_tensor(self.seed, t(K_WGATE), (d.intermediate_dim, d.hidden_dim), STD_W),
                "wup": synth_tensor(self.seed, t(K_WUP), (d.intermediate_dim, d.hidden_dim), STD_W),
                "wdown": synth_tensor(self.seed, t(K_WDOWN), (d.hidden_dim, d.intermediate_dim), STD_W),
            }
            if d.quant_method == 1:
                for name in ("wqkv", "wo", "wgate", "wup", "wdown"):
                    q, s = quantize_weight_per_channel(w[name])
                    w[name + "_q"] = q
                    w[name + "_s"] = s
            elif d.quant_method == 2:
                for name in ("wqkv", "wo", "wgate", "wup", "wdown"):
                    w[name + "_w4"] = quantize_weight_w4(w[name])[2]  # the fp16 operand fp16(q * scale)
            return w

        return self._get(("layer", l), make)
