"""shared helpers for the GPU parity tests (torch = device memory only)."""
import ctypes as C

import numpy as np
import torch

from ppl_llm_serving_b200 import capi
from ppl_llm_serving_b200.engine import _ptr

INT64_MAX = np.iinfo(np.int64).max


def dev(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def make_step_c(step, keep: list) -> capi.StepC:
    """oracle.llama_ref.Step -> device-resident b2llm_step (tensors appended to `keep`)."""
    s = capi.StepC()
    tok, ss, ks, sp = dev(step.token_inputs), dev(step.seq_starts), dev(step.kv_starts), dev(step.start_pos)
    idx_np = step.cache_indices if step.page_list is None else step.page_list
    idx = dev(idx_np)
    keep += [tok, ss, ks, sp, idx]
    s.token_ids, s.seq_starts, s.kv_starts = tok.data_ptr(), ss.data_ptr(), ks.data_ptr()
    s.start_pos, s.cache_indices = sp.data_ptr(), idx.data_ptr()
    s.num_tokens, s.batch = len(step.token_inputs), step.batch
    s.decoding_batches = step.decoding_batches
    s.max_seq_len, s.max_kv_len, s.max_pages = step.max_seq_len, step.max_kv_len, step.max_pages
    s.cache_prefill = 0
    return s


def make_geom(desc, max_tokens) -> capi.KvGeomC:
    g = capi.KvGeomC()
    g.num_layers, g.num_kv_heads, g.head_dim = desc.num_layers, desc.num_kv_heads, desc.head_dim
    g.quant_group, g.cache_layout, g.cache_mode = desc.cache_quant_group, desc.cache_layout, desc.cache_mode
    g.page_size, g.max_tokens = desc.page_size, max_tokens
    return g


def random_pages(rng, num_seqs, pages_per_seq, page_size, total_pages):
    """non-contiguous page tables: entries are page begin-token indices"""
    perm = rng.permutation(total_pages)[: num_seqs * pages_per_seq]
    return [list((perm[i * pages_per_seq:(i + 1) * pages_per_seq] * page_size).astype(np.int64)) for i in range(num_seqs)]
