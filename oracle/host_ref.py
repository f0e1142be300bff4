"""Restatement of the host-side integer logic on the hot path (pure Python; small cases only).

Test infrastructure (see ``oracle/__init__.py``).  Unlike the arithmetic of the forward, these
ARE pinnable from reference source, and ``tests/test_host_kat.py`` pins them.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

M64 = (1 << 64) - 1
M32 = (1 << 32) - 1
INT64_MAX = (1 << 63) - 1


def hash_combine(prev: int, vec, length: int | None = None) -> int:
    """``utils::HashCombine`` (src/utils/utils.cc:87-94), including C's integer promotions:
    ``vec[i] + 0x9e3779b9`` is int32 + unsigned int -> 32-bit unsigned wrap, then widened."""
    n = len(vec) if length is None else length
    seed = n & M64
    seed ^= ((prev + 0x9E3779B9) + ((seed << 6) & M64) + (seed >> 2)) & M64
    for i in range(n):
        t = (int(vec[i]) + 0x9E3779B9) & M32
        seed ^= (t + ((seed << 6) & M64) + (seed >> 2)) & M64
    return seed


def page_count(first_fill_len: int, rest_iters: int, page_size: int) -> int:
    """pages reserved at admission (src/generator/llm_generator.cc:484,553): the whole lifetime,
    ``total_len = first_fill_len + rest_iters - 1``."""
    total_len = first_fill_len + rest_iters - 1
    return (total_len + page_size - 1) // page_size


def build_model_input(seq_next_tokens, start_pos):
    """seq_starts / kv_starts / max_seq_len / max_kv_len as ``UpdateInput`` builds them
    (src/generator/llm_generator.cc:263-298)."""
    token_inputs, seq_starts, kv_starts = [], [0], [0]
    max_seq_len = max_kv_len = 0
    for toks, sp in zip(seq_next_tokens, start_pos):
        token_inputs.extend(toks)
        seq_starts.append(seq_starts[-1] + len(toks))
        kv_starts.append(kv_starts[-1] + sp + len(toks))
        max_seq_len = max(max_seq_len, len(toks))
        max_kv_len = max(max_kv_len, sp + len(toks))
    return token_inputs, seq_starts, kv_starts, max_seq_len, max_kv_len


def kv_cache_max_tokens(scale: float, avail_bytes: int, num_layers, num_kv_heads, tp, hidden_dim, num_heads,
                        cache_quant_bit, cache_quant_group):
    """KV budget of ``CudaResourceManager`` (src/backends/cuda/resource_manager.cc:329-342,381-388).

    The reference evaluates ``scale * avail * cb / (cb + sb)`` left to right with ``scale`` a float,
    i.e. in fp32, then truncates to uint64; reproduced with numpy float32.
    """
    size_kv = {0: 2, 8: 1}[cache_quant_bit]
    cb = num_layers * 2 * num_kv_heads // tp * hidden_dim // num_heads * size_kv
    sb = 0
    if cache_quant_bit > 0:
        sb = num_layers * 2 * num_kv_heads // tp * hidden_dim // num_heads // cache_quant_group * 2
    f = np.float32(scale) * np.float32(avail_bytes)
    f = np.float32(f * np.float32(cb))
    f = np.float32(f / np.float32(cb + sb))
    return int(np.uint64(f)) // cb, cb, sb


def finished(rest_iters: int, early_stopping: bool, token: int, stop_tokens, request_stop_tokens) -> bool:
    """finish rule of the step loop (src/generator/llm_generator.cc:720-726)."""
    if rest_iters == 0:
        return True
    return bool(early_stopping and (token in stop_tokens or token in request_stop_tokens))


class PrefixCacheModel:
    """``utils::PrefixCacheManager`` (src/utils/prefix_cache_manager.h:111-186): hash -> page with
    refcounts; pages whose refcount reaches 0 enter an LRU (most recent at the head, eviction from
    the tail)."""

    def __init__(self):
        self.map = {}            # hash -> [page, refcount]
        self.lru = OrderedDict()  # insertion order: oldest first == tail of the reference's list

    def find(self, h):
        return self.map[h][0] if h in self.map else -1

    def insert(self, h, page):
        if h not in self.map:
            self.map[h] = [page, 1]

    def inc_ref(self, hashes):
        for h in hashes:
            if h not in self.map:
                break
            self.map[h][1] += 1
            self.lru.pop(h, None)

    def dec_ref(self, hashes):
        for h in hashes:
            if h not in self.map:
                break
            self.map[h][1] -= 1
            if self.map[h][1] == 0 and h not in self.lru:
                self.lru[h] = self.map[h][0]

    def evict(self, n):
        pages = []
        for _ in range(min(n, len(self.lru))):
            h, page = self.lru.popitem(last=False)
            pages.append(page)
            self.map.pop(h, None)
        return pages

    def size(self):
        return len(self.map)
