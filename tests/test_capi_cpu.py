"""The C-ABI library without a GPU: it loads, exports every symbol include/b2llm.h declares, the ctypes
binding covers the same set, pure-host entry points work, and device entry points fail LOUDLY with a
RetCode (no CPU fallback) when there is no B200."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from oracle import llama_ref as ref
from ppl_llm_serving_b200 import capi

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "b2llm.h").read_text()


def declared_symbols():
    return sorted(set(re.findall(r"B2LLM_API\s+[\w\s\*]+?\b(b2llm_\w+)\s*\(", HEADER)))


def test_header_declares_what_the_binding_binds():
    names = declared_symbols()
    assert len(names) >= 25
    assert set(names) == set(capi.SIGNATURES), (set(names) ^ set(capi.SIGNATURES))


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), f"libb2llm.so does not export {name}"
    assert lib.b2llm_version().decode().startswith("b2llm")


def test_no_torch_or_cxx_types_in_the_abi():
    body = HEADER.split('extern "C" {', 1)[1]
    assert "std::" not in body and "at::" not in body and "torch" not in body
    assert "cudaStream_t" not in re.sub(r"/\*.*?\*/", "", body, flags=re.S)   # streams cross as void*


def test_struct_layouts_match_header():
    assert C.sizeof(capi.ModelDescC) == 4 * 20
    assert C.sizeof(capi.StepC) == 8 * 5 + 8 * 6 + 8
    assert C.sizeof(capi.KvGeomC) == 4 * 8 + 8


def test_rope_table_host_entry_point(lib):
    cos = np.empty((64, 64), np.float32)
    sin = np.empty((64, 64), np.float32)
    assert lib.b2llm_rope_table(64, 128, 10000.0, cos.ctypes.data, sin.ctypes.data) == 0
    ecos, esin = ref.rope_table(64, 128, 10000.0)
    assert np.array_equal(cos, ecos) and np.array_equal(sin, esin)
    assert lib.b2llm_rope_table(0, 128, 10000.0, cos.ctypes.data, sin.ctypes.data) == 2  # RC_INVALID_VALUE
    assert b"rope_table" in lib.b2llm_last_error()


def test_workspace_sizes_are_host_only(lib):
    assert lib.b2llm_sample_topk_topp_get_workspace_size(1024, 32000, 1) >= 0
    assert lib.b2llm_sample_topk_topp_get_workspace_size(1024, 32000, 50) > 0
    assert lib.b2llm_attention_workspace_size(1024, 32, 128) > 0


def test_engine_create_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from oracle.weights import ModelDesc
    d = capi.desc_to_c(ModelDesc(256, 512, 1, 2, 2, 128), 16, 4)
    eng = C.c_void_p()
    rc = lib.b2llm_engine_create(C.byref(d), 0, 1, None, None, C.byref(eng))
    assert rc == 5 and not eng.value                      # RC_DEVICE_RUNTIME_ERROR, nothing created
    assert b"no CPU fallback" in lib.b2llm_last_error()
    # argument validation happens before any device work and is reported as RC_INVALID_VALUE
    d.num_heads = 3
    assert lib.b2llm_engine_create(C.byref(d), 0, 1, None, None, C.byref(eng)) == 2


def test_missing_library_raises():
    with pytest.raises(capi.B2llmError, match="no fallback"):
        capi.load_library("/nonexistent/libb2llm.so")


def test_argument_validation_precedes_device_work(lib):
    """error behaviour of the boundary: RetCode values, never exceptions or crashes; null / malformed
    arguments are rejected before any CUDA call, so this runs without a GPU"""
    P = C.c_void_p
    assert lib.b2llm_op_rmsnorm_quant(None, None, None, None, 1e-5, 4, 256, None, None, None) == 2
    assert lib.b2llm_op_quant_rows(None, None, 4, 256, None, None) == 2
    assert lib.b2llm_op_gemm_w8a8(None, None, None, None, None, 4, 256, 256, 0, None, 0) == 2
    one = (C.c_byte * 16)()
    p = C.cast(one, P)
    assert lib.b2llm_op_gemm_w8a8(None, p, p, p, p, 4, 256, 256, 7, p, 0) == 2          # unknown epilogue
    assert lib.b2llm_op_gemm_f16(None, None, None, 4, 256, 256, 0, None, 0, 0) == 2
    assert lib.b2llm_op_gemm_w4a16(None, None, None, None, 4, 256, 256, 0, None) == 2
    assert lib.b2llm_op_attention(None, None, None, 4, None, 0, None, None, None, None, 0) == 2
    assert lib.b2llm_op_rope_kv_append(None, None, None, 4, None, 0, None, None, None, None) == 2
    assert lib.b2llm_engine_set_inputs(None, None, 0, None, None, None, 0, None, 0, 0, 0, 0, 0) == 2
    assert lib.b2llm_engine_forward(None, None, None, None) == 2
    assert lib.b2llm_engine_bind_kv(None, None, None, 0) == 2
    assert lib.b2llm_engine_reserve(None, 1, 1) == 2
    assert lib.b2llm_engine_configure(None, 3, 1) == 2
    assert lib.b2llm_engine_destroy(None) == 0                                          # destroying nothing is fine
    assert lib.b2llm_engine_last_launch_count(None) == 0
    assert lib.b2llm_last_error()                                                        # text of the last failure
