"""ctypes binding of the C ABI in ``include/b2llm.h`` (libb2llm.so).

This is plumbing only: torch supplies device memory and streams, every computation happens in
the hand-written sm_100a kernels behind the C ABI.  There is no CPU or PyTorch fallback -- if the
library is missing or no B200 is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
# B2LLM_LIB: another build of the same library (A/B measurements of kernel variants); default = the in-tree build
LIB_PATH = Path(os.environ["B2LLM_LIB"]).resolve() if os.environ.get("B2LLM_LIB") else _HERE / "lib" / "libb2llm.so"

B2LLM_OK = 0
QUANT_NONE, QUANT_ONLINE_I8I8, QUANT_W4A16 = 0, 1, 2
W_EMBEDDING, W_FINAL_NORM, W_LM_HEAD, W_ATTN_NORM, W_QKV, W_O, W_FFN_NORM, W_GATE, W_UP, W_DOWN = range(10)
EPI_F16, EPI_RESIDUAL, EPI_SWIGLU, EPI_F32 = range(4)


class ModelDescC(C.Structure):
    _fields_ = [
        ("hidden_dim", C.c_int32), ("intermediate_dim", C.c_int32), ("num_layers", C.c_int32),
        ("num_heads", C.c_int32), ("num_kv_heads", C.c_int32), ("vocab_size", C.c_int32),
        ("norm_eps", C.c_float), ("rope_theta", C.c_float),
        ("cache_quant_bit", C.c_int32), ("cache_quant_group", C.c_int32), ("cache_layout", C.c_int32),
        ("cache_mode", C.c_int32), ("page_size", C.c_int32), ("quant_method", C.c_int32),
        ("max_position", C.c_int32), ("max_tokens_per_step", C.c_int32), ("max_running_batch", C.c_int32),
        ("reserved", C.c_int32 * 3),
    ]


class StepC(C.Structure):
    _fields_ = [
        ("token_ids", C.c_void_p), ("seq_starts", C.c_void_p), ("kv_starts", C.c_void_p),
        ("cache_indices", C.c_void_p), ("start_pos", C.c_void_p),
        ("num_tokens", C.c_int64), ("batch", C.c_int64), ("decoding_batches", C.c_int64),
        ("max_seq_len", C.c_int64), ("max_kv_len", C.c_int64), ("max_pages", C.c_int64),
        ("cache_prefill", C.c_int32), ("reserved", C.c_int32),
    ]


class KvGeomC(C.Structure):
    _fields_ = [
        ("num_layers", C.c_int32), ("num_kv_heads", C.c_int32), ("head_dim", C.c_int32), ("quant_group", C.c_int32),
        ("cache_layout", C.c_int32), ("cache_mode", C.c_int32), ("page_size", C.c_int32), ("reserved", C.c_int32),
        ("max_tokens", C.c_uint64),
    ]


# every symbol include/b2llm.h declares: name -> (restype, argtypes)
_P, _I32, _I64, _U64, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float
SIGNATURES = {
    "b2llm_version": (C.c_char_p, []),
    "b2llm_last_error": (C.c_char_p, []),
    "b2llm_engine_create": (_I32, [C.POINTER(ModelDescC), _I32, _I32, _P, _P, C.POINTER(_P)]),
    "b2llm_engine_destroy": (_I32, [_P]),
    "b2llm_engine_reserve": (_I32, [_P, _I64, _I64]),
    "b2llm_engine_configure": (_I32, [_P, _I32, _I64]),
    "b2llm_engine_load_weight": (_I32, [_P, _I32, _I32, _P, _U64]),
    "b2llm_engine_load_weight_shard": (_I32, [_P, _I32, _I32, _P, _U64]),
    "b2llm_engine_random_init": (_I32, [_P, _U64]),
    "b2llm_engine_bind_kv": (_I32, [_P, _P, _P, _U64]),
    "b2llm_engine_kv_bytes_per_token": (_I32, [_P, C.POINTER(_U64), C.POINTER(_U64)]),
    "b2llm_engine_set_inputs": (_I32, [_P, _P, _I64, _P, _P, _P, _I64, _P, _I64, _I64, _I64, _I64, _I32]),
    "b2llm_engine_run": (_I32, [_P, _I32, C.POINTER(_P), C.POINTER(_I64)]),
    "b2llm_engine_forward": (_I32, [_P, C.POINTER(StepC), C.POINTER(_P), C.POINTER(_I64)]),
    "b2llm_engine_staged_inputs": (_I32, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    "b2llm_engine_last_launch_count": (_I64, [_P]),
    "b2llm_engine_tp_join_stats": (_I32, [_P, C.POINTER(C.c_double)]),
    "b2llm_attention_decode_plan": (_I32, [_I64, _I32, _I32, _I64, C.POINTER(_I32), C.POINTER(_I32)]),
    "b2llm_debug_attention_trace": (_I32, [C.c_void_p, _I64]),
    "b2llm_engine_profile": (_I32, [_P, _I32]),
    "b2llm_engine_profile_read": (_I32, [_P, C.POINTER(C.c_double), C.POINTER(_I64), _I32]),
    "b2llm_engine_debug_read": (_I32, [_P, _I32, _P, _U64]),
    "b2llm_sample_topk_topp_get_workspace_size": (_I64, [_I32, _I32, _I32]),
    "b2llm_sample_topk_topp": (_I32, [_P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _F, _F, _P, _P, _P]),
    "b2llm_apply_penalty": (_I32, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I32, _I32, _P, _P]),
    "b2llm_op_rmsnorm_quant": (_I32, [_P, _P, _P, _P, _F, _I64, _I32, _P, _P, _P]),
    "b2llm_op_quant_rows": (_I32, [_P, _P, _I64, _I32, _P, _P]),
    "b2llm_op_gemm_w8a8": (_I32, [_P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _P, _I32]),
    "b2llm_op_gemm_f16": (_I32, [_P, _P, _P, _I64, _I32, _I32, _I32, _P, _I64, _I32]),
    "b2llm_op_rope_kv_append": (_I32, [_P, _P, C.POINTER(StepC), _I32, C.POINTER(KvGeomC), _I32, _P, _P, _P, _P]),
    "b2llm_attention_workspace_size": (_I64, [_I64, _I32, _I32]),
    "b2llm_op_attention": (_I32, [_P, _P, C.POINTER(StepC), _I32, C.POINTER(KvGeomC), _I32, _P, _P, _P, _P, _I32]),
    "b2llm_rope_table": (_I32, [_I32, _I32, _F, _P, _P]),
    "b2llm_op_synth_fp16": (_I32, [_P, _U64, _U64, _U64, _F, _F, _P]),
    "b2llm_op_quant_weight": (_I32, [_P, _P, _I32, _I32, _P, _P]),
    "b2llm_op_quant_weight_w4": (_I32, [_P, _P, _I32, _I32, _P, _P]),
    "b2llm_op_dequant_w4": (_I32, [_P, _P, _P, _I32, _I32, _P]),
    "b2llm_op_gemm_w4a16": (_I32, [_P, _P, _P, _P, _I64, _I32, _I32, _I32, _P]),
}

_lib = None


class B2llmError(RuntimeError):
    pass


def load_library(path: os.PathLike | None = None) -> C.CDLL:
    """dlopen libb2llm.so and type every entry point.  Raises if it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise B2llmError(f"{p} not found: build it with `python __graft_entry__.py build` "
                         f"(make -C ppl.llm.serving_b200/csrc); there is no fallback path")
    lib = C.CDLL(str(p), mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != B2LLM_OK:
        msg = load_library().b2llm_last_error().decode(errors="replace")
        raise B2llmError(f"{what} failed with RetCode {rc}: {msg}")


def desc_to_c(d, max_tokens_per_step: int, max_running_batch: int) -> ModelDescC:
    """``d``: any object with ModelConfig-like attributes (e.g. oracle.weights.ModelDesc)."""
    c = ModelDescC()
    for f in ("hidden_dim", "intermediate_dim", "num_layers", "num_heads", "num_kv_heads", "vocab_size",
              "cache_quant_bit", "cache_quant_group", "cache_layout", "cache_mode", "page_size", "quant_method",
              "max_position"):
        setattr(c, f, int(getattr(d, f)))
    c.norm_eps = float(d.norm_eps)
    c.rope_theta = float(d.rope_theta)
    c.max_tokens_per_step = int(max_tokens_per_step)
    c.max_running_batch = int(max_running_batch)
    return c
