"""Python host mirror of the reference's engine objects over the C ABI.

Names, argument meaning and error behaviour follow the reference so that tests read like the
reference's own call sites:

  ModelInput / ModelOutput          src/engine/llm_engine.h:40-73
  LLMEngine.Execute                 src/engine/llm_engine.cc:171-236
  CudaPostProcessor                 src/backends/cuda/post_processor.{h,cc}
  CudaResourceManager (KV budget)   src/backends/cuda/resource_manager.cc:329-362,381-388

torch is used for device memory and streams only.  The C++ mirror (same classes, same signatures,
for linking the reference's tools) lives in ``host/``; this module exists so that pytest and
bench.py can drive the same C ABI.
"""
from __future__ import annotations

import ctypes as C
import time
from dataclasses import dataclass, field

import numpy as np
import torch

from . import capi

RC_SUCCESS, RC_OTHER_ERROR, RC_INVALID_VALUE = 0, 1, 2
INT64_MAX = np.iinfo(np.int64).max


@dataclass
class ModelConfig:
    """ppl::llm::ModelConfig (src/common/config.h:64-84) + the graph-level constants and quant method
    the reference keeps elsewhere (resource_manager.cc:49-56; config.h:74)."""
    hidden_dim: int = 0
    intermediate_dim: int = 0
    num_layers: int = 0
    num_heads: int = 0
    num_kv_heads: int = 0
    vocab_size: int = 0
    norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    cache_quant_bit: int = 8
    cache_quant_group: int = 8
    cache_layout: int = 3
    cache_mode: int = 1
    page_size: int = 16
    dynamic_batching: bool = True
    auto_causal: bool = True
    quant_method: int = 1       # 0 "none", 1 "online_i8i8"
    max_position: int = 4096

    @property
    def head_dim(self):
        return self.hidden_dim // self.num_heads


def ParseModelConfig(model_param_path: str, model_config: ModelConfig) -> bool:
    """params.json reader with the reference's required keys and defaults (src/common/config.cc:31-148)."""
    import json
    try:
        with open(model_param_path) as f:
            doc = json.load(f)
    except (OSError, ValueError):
        return False
    required = ["num_heads", "num_layers", "hidden_dim", "intermediate_dim", "vocab_size", "cache_quant_bit",
                "cache_quant_group", "cache_layout", "cache_mode", "dynamic_batching", "auto_causal"]
    for k in required:
        if k not in doc:
            return False
    for k in required:
        setattr(model_config, k, type(getattr(model_config, k))(doc[k]))
    model_config.num_kv_heads = int(doc.get("num_kv_heads", model_config.num_heads))
    if model_config.cache_mode == 1:
        if "page_size" not in doc:
            return False
        model_config.page_size = int(doc["page_size"])
    return True


LLAMA2_7B = dict(hidden_dim=4096, intermediate_dim=11008, num_layers=32, num_heads=32, num_kv_heads=32, vocab_size=32000)
LLAMA2_13B = dict(hidden_dim=5120, intermediate_dim=13824, num_layers=40, num_heads=40, num_kv_heads=40, vocab_size=32000)
LLAMA2_70B = dict(hidden_dim=8192, intermediate_dim=28672, num_layers=80, num_heads=64, num_kv_heads=8, vocab_size=32000)


@dataclass
class ModelInput:
    """src/engine/llm_engine.h:40-60 (host vectors, rebuilt every step by the generator)."""
    decoding_batches: int = 0
    max_seq_len: int = 0
    max_kv_len: int = 0
    max_pages: int = 0
    token_inputs: list = field(default_factory=list)
    seq_starts: list = field(default_factory=list)
    start_pos: list = field(default_factory=list)
    cache_indices: list = field(default_factory=list)
    page_list: list = field(default_factory=list)
    kv_starts: list = field(default_factory=list)
    temperatures: list = field(default_factory=list)
    top_p_list: list = field(default_factory=list)
    top_k_list: list = field(default_factory=list)
    repetition_penalty_list: list = field(default_factory=list)
    presence_penalty_list: list = field(default_factory=list)
    frequency_penalty_list: list = field(default_factory=list)
    batch_slots: list = field(default_factory=list)


@dataclass
class ModelOutput:
    """src/engine/llm_engine.h:62-73"""
    output_token: np.ndarray = None
    logprobs: np.ndarray = None

    def Resize(self, n: int):
        self.output_token = np.zeros(n, dtype=np.int32)
        self.logprobs = np.zeros(n, dtype=np.float32)


@dataclass
class StepCounter:
    """the `current` half of WorkerPerStepCounter (src/common/profiler.h:60-73), microseconds"""
    set_input_cost: int = 0
    model_forward_cost: int = 0
    choose_token_cost: int = 0
    output_token_cnt: int = 0


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        return C.c_void_p(t.data_ptr())
    if isinstance(t, np.ndarray):
        return C.c_void_p(t.ctypes.data)
    raise TypeError(type(t))


def _i64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64))


class CudaPostProcessor:
    """Mirror of ppl::llm::cuda::CudaPostProcessor (post_processor.cc:71-281)."""

    def __init__(self, stream: torch.cuda.Stream):
        self.lib = capi.load_library()
        self.stream = stream
        self.max_batch = 0
        self.workspace = None
        self.count_map = None
        self._rand = np.random.RandomState(1)  # stands in for the reference's unseeded host rand()

    def InitPostProcessorMem(self, max_running_batch: int, vocab_size: int, enable_penalty: bool) -> int:
        dev = torch.device("cuda", torch.cuda.current_device())
        self.max_batch = max_running_batch
        self.temperatures = torch.empty(max_running_batch, dtype=torch.float32, device=dev)
        self.top_p = torch.empty(max_running_batch, dtype=torch.float32, device=dev)
        self.rand = torch.empty(max_running_batch, dtype=torch.float32, device=dev)
        self.output = torch.empty(max_running_batch, dtype=torch.int32, device=dev)
        self.logprobs = torch.empty(max_running_batch, dtype=torch.float32, device=dev)
        self.out_host = torch.empty(max_running_batch, dtype=torch.int32).pin_memory()
        self.lp_host = torch.empty(max_running_batch, dtype=torch.float32).pin_memory()
        if enable_penalty:
            self.count_map = torch.zeros(max_running_batch * vocab_size, dtype=torch.int16, device=dev)
            self.batch_slots = torch.empty(max_running_batch, dtype=torch.int64, device=dev)
            self.rep = torch.empty(max_running_batch, dtype=torch.float32, device=dev)
            self.presence = torch.empty(max_running_batch, dtype=torch.float32, device=dev)
            self.freq = torch.empty(max_running_batch, dtype=torch.float32, device=dev)
        return RC_SUCCESS

    def SampleTopKTopP(self, logits_device: int, temperatures_host, top_k_host, top_p_host, batch, vocab_size,
                       batch_stride, default_top_k, default_top_p, req_list_changed, output_host, logprobs_host,
                       enable_penalty, rand_host=None) -> int:
        """logits_device: raw device address (the runtime's output tensor buffer)."""
        if enable_penalty:
            temperatures_host = None  # already applied by ApplyPenalty (post_processor.cc:126)
        ws_bytes = 0
        if default_top_k > 0:
            ws_bytes = self.lib.b2llm_sample_topk_topp_get_workspace_size(batch, vocab_size, default_top_k)
        if ws_bytes and (self.workspace is None or self.workspace.numel() < ws_bytes):
            self.workspace = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
        temps_opt = top_p_opt = None
        with torch.cuda.stream(self.stream):
            if req_list_changed:  # the reference's quirk: only then are per-request values passed
                if temperatures_host is not None and len(temperatures_host):
                    self.temperatures[:batch].copy_(torch.as_tensor(np.asarray(temperatures_host, dtype=np.float32)),
                                                    non_blocking=True)
                    temps_opt = self.temperatures
                if top_p_host is not None and len(top_p_host):
                    self.top_p[:batch].copy_(torch.as_tensor(np.asarray(top_p_host, dtype=np.float32)), non_blocking=True)
                    top_p_opt = self.top_p
            default_rand = float(self._rand.random_sample()) if rand_host is None else 0.0
            rnd = np.asarray(self._rand.random_sample(batch) if rand_host is None else rand_host, dtype=np.float32)
            self.rand[:batch].copy_(torch.as_tensor(rnd), non_blocking=True)
            rc = self.lib.b2llm_sample_topk_topp(
                C.c_void_p(self.stream.cuda_stream), C.c_void_p(logits_device), _ptr(temps_opt), _ptr(top_p_opt),
                _ptr(self.rand), batch, vocab_size, batch_stride, default_top_k, default_top_p, default_rand,
                _ptr(self.workspace), _ptr(self.output), _ptr(self.logprobs))
            if rc != RC_SUCCESS:
                return rc
            self.out_host[:batch].copy_(self.output[:batch], non_blocking=True)
            self.lp_host[:batch].copy_(self.logprobs[:batch], non_blocking=True)
        self.stream.synchronize()  # the step's only host/device join (post_processor.cc:212)
        output_host[:batch] = self.out_host[:batch].numpy()
        logprobs_host[:batch] = self.lp_host[:batch].numpy()
        return RC_SUCCESS

    def ApplyPenalty(self, temperatures_host, repetition_penalties_host, presence_penalties_host,
                     frequency_penalties_host, batch_slots_host, token_inputs_dev: int, seqstarts_dev: int,
                     start_pos_dev: int, batch, vocab_size, req_list_changed, logits_dev: int) -> int:
        pres_opt = freq_opt = None
        with torch.cuda.stream(self.stream):
            if req_list_changed:
                f32 = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32))
                self.temperatures[:batch].copy_(f32(temperatures_host), non_blocking=True)
                self.batch_slots[:batch].copy_(torch.as_tensor(_i64(batch_slots_host)), non_blocking=True)
                self.rep[:batch].copy_(f32(repetition_penalties_host), non_blocking=True)
                if presence_penalties_host is not None:
                    self.presence[:batch].copy_(f32(presence_penalties_host), non_blocking=True)
                    pres_opt = self.presence
                if frequency_penalties_host is not None:
                    self.freq[:batch].copy_(f32(frequency_penalties_host), non_blocking=True)
                    freq_opt = self.freq
            return self.lib.b2llm_apply_penalty(
                C.c_void_p(self.stream.cuda_stream), C.c_void_p(logits_dev), _ptr(self.temperatures), _ptr(self.rep),
                _ptr(pres_opt), _ptr(freq_opt), _ptr(self.batch_slots), C.c_void_p(token_inputs_dev),
                C.c_void_p(seqstarts_dev), C.c_void_p(start_pos_dev), batch, vocab_size, _ptr(self.count_map),
                C.c_void_p(logits_dev))


class CudaResourceManager:
    """The parts of ppl::llm::cuda::CudaResourceManager on the path: engine bring-up and the KV budget
    ``max_tokens = floor(scale * free * cb / (cb + sb)) / cb`` (resource_manager.cc:329-342), evaluated
    in fp32 like the reference."""

    def __init__(self):
        self.lib = capi.load_library()
        self.engine = None
        self.kv_cache_max_tokens = 0

    def Init(self, model_desc, max_tokens_scale: float, max_running_batch: int, max_tokens_per_step: int,
             enable_penalty: bool = False, kv_cache_max_tokens: int | None = None, seed: int | None = 0xB200,
             device: int = 0, tensor_parallel_size: int = 1, rank: int = 0, nccl_comm=None) -> int:
        """``tensor_parallel_size`` > 1: this process is rank ``rank`` of a TP group (one process per GPU) and
        ``nccl_comm`` its raw ncclComm_t (nccl.create_comm); every rank must use the same ``kv_cache_max_tokens``
        (the reference takes rank 0's number for all, resource_manager.cc:329-344)."""
        torch.cuda.set_device(device)
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        self.desc = model_desc
        self.tp, self.rank = tensor_parallel_size, rank
        dc = capi.desc_to_c(model_desc, max_tokens_per_step, max_running_batch)
        eng = C.c_void_p()
        rc = self.lib.b2llm_engine_create(C.byref(dc), rank, tensor_parallel_size, nccl_comm,
                                          C.c_void_p(self.stream.cuda_stream), C.byref(eng))
        if rc != RC_SUCCESS:
            return rc
        self.engine = eng
        if seed is not None:
            rc = self.lib.b2llm_engine_random_init(eng, seed)
            if rc != RC_SUCCESS:
                return rc
        self.post_processor = CudaPostProcessor(self.stream)
        rc = self.post_processor.InitPostProcessorMem(max_running_batch, model_desc.vocab_size, enable_penalty)
        if rc != RC_SUCCESS:
            return rc
        cb, sb = C.c_uint64(), C.c_uint64()
        self.lib.b2llm_engine_kv_bytes_per_token(eng, C.byref(cb), C.byref(sb))
        cb, sb = cb.value, sb.value
        if kv_cache_max_tokens is None:
            free, _total = torch.cuda.mem_get_info(device)
            f = np.float32(max_tokens_scale) * np.float32(free)
            f = np.float32(np.float32(f * np.float32(cb)) / np.float32(cb + sb))
            kv_cache_max_tokens = int(np.uint64(f)) // cb
            if tensor_parallel_size > 1:  # all ranks use rank 0's budget (resource_manager.cc:329-344)
                import torch.distributed as dist
                t = torch.tensor([kv_cache_max_tokens], dtype=torch.int64, device="cuda")
                dist.broadcast(t, src=0)
                kv_cache_max_tokens = int(t.item())
        self.kv_cache_max_tokens = int(kv_cache_max_tokens)
        self.kv_cache_mem = torch.empty(self.kv_cache_max_tokens * cb, dtype=torch.int8, device="cuda")
        # cache_quant_bit 0: fp16 cache, no scale memory (resource_manager.cc:353-361 allocates it only when sb > 0)
        self.kv_scale_mem = torch.empty(self.kv_cache_max_tokens * sb // 2, dtype=torch.float16, device="cuda") if sb else None
        return self.lib.b2llm_engine_bind_kv(eng, _ptr(self.kv_cache_mem), _ptr(self.kv_scale_mem),
                                             self.kv_cache_max_tokens)

    def load_weights(self, weights) -> None:
        """upload an ``oracle.weights.SynthWeights``-like object (fp16 host tensors) through the ABI"""
        L = self.lib
        e = self.engine

        def up(kind, layer, a):
            a = np.ascontiguousarray(a, dtype=np.float16)
            capi.check(L.b2llm_engine_load_weight(e, kind, layer, _ptr(a), a.size), f"load_weight({kind},{layer})")

        up(capi.W_EMBEDDING, 0, weights.embedding())
        up(capi.W_FINAL_NORM, 0, weights.final_norm())
        up(capi.W_LM_HEAD, 0, weights.lm_head())
        for l in range(self.desc.num_layers):
            w = weights.layer(l)
            up(capi.W_ATTN_NORM, l, w["attn_norm"])
            up(capi.W_QKV, l, w["wqkv"])
            up(capi.W_O, l, w["wo"])
            up(capi.W_FFN_NORM, l, w["ffn_norm"])
            up(capi.W_GATE, l, w["wgate"])
            up(capi.W_UP, l, w["wup"])
            up(capi.W_DOWN, l, w["wdown"])

    def close(self):
        if self.engine is not None:
            self.lib.b2llm_engine_destroy(self.engine)
            self.engine = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LLMEngine:
    """Mirror of ppl::llm::LLMEngine (llm_engine.h:105-164, llm_engine.cc:171-236)."""

    def __init__(self, resource: CudaResourceManager, enable_penalty: bool, top_k: int, top_p: float):
        self.res = resource
        self.lib = resource.lib
        self.enable_penalty = enable_penalty
        self.top_k = top_k
        self.top_p = top_p
        self.step_counter = StepCounter()
        self.logits_ptr = 0
        self.logits_stride = 0

    def SetInput(self, mi: ModelInput, req_list_changed: bool) -> int:
        d = self.res.desc
        tok, ss, ks, sp = _i64(mi.token_inputs), _i64(mi.seq_starts), _i64(mi.kv_starts), _i64(mi.start_pos)
        idx = _i64(mi.cache_indices if d.cache_mode == 0 else mi.page_list)
        self._keep = (tok, ss, ks, sp, idx)
        return self.lib.b2llm_engine_set_inputs(
            self.res.engine, _ptr(tok), len(tok), _ptr(ss), _ptr(ks), _ptr(sp), len(sp),
            _ptr(idx) if idx.size else None, int(mi.max_pages), int(mi.decoding_batches), int(mi.max_seq_len),
            int(mi.max_kv_len), 1 if req_list_changed else 0)

    def RunModel(self, is_prefix_cache_hit: bool) -> int:
        lp, ls = C.c_void_p(), C.c_int64()
        rc = self.lib.b2llm_engine_run(self.res.engine, 1 if is_prefix_cache_hit else 0, C.byref(lp), C.byref(ls))
        self.logits_ptr, self.logits_stride = lp.value or 0, ls.value
        return rc

    def Execute(self, model_input: ModelInput, req_list_changed: bool, is_prefix_cache_hit: bool,
                model_output: ModelOutput, rand_host=None):
        """returns (RetCode, error_msg); fills model_output (pre-Resize()d by the caller)."""
        running_batch = len(model_input.start_pos)
        t0 = time.perf_counter()
        rc = self.SetInput(model_input, req_list_changed)
        t1 = time.perf_counter()
        self.step_counter.set_input_cost = int((t1 - t0) * 1e6)
        if rc != RC_SUCCESS:
            return RC_OTHER_ERROR, "SetInputTask failed: " + self.lib.b2llm_last_error().decode()
        rc = self.RunModel(is_prefix_cache_hit)
        t2 = time.perf_counter()
        self.step_counter.model_forward_cost = int((t2 - t1) * 1e6)
        if rc != RC_SUCCESS:
            return RC_OTHER_ERROR, "RunModelTask failed: " + self.lib.b2llm_last_error().decode()
        pp = self.res.post_processor
        vocab = self.res.desc.vocab_size
        if self.enable_penalty:
            tp, sp_, st = C.c_void_p(), C.c_void_p(), C.c_void_p()
            self.lib.b2llm_engine_staged_inputs(self.res.engine, C.byref(tp), C.byref(sp_), C.byref(st))
            rc = pp.ApplyPenalty(model_input.temperatures, model_input.repetition_penalty_list, None, None,
                                 model_input.batch_slots, tp.value, sp_.value, st.value, running_batch, vocab,
                                 req_list_changed, self.logits_ptr)
            if rc != RC_SUCCESS:
                return RC_OTHER_ERROR, "Apply Penalty failed: " + self.lib.b2llm_last_error().decode()
        default_top_k = self.top_k if not model_input.top_k_list else model_input.top_k_list[0]
        rc = pp.SampleTopKTopP(self.logits_ptr, model_input.temperatures, model_input.top_k_list,
                               model_input.top_p_list, running_batch, vocab, self.logits_stride, default_top_k,
                               self.top_p, req_list_changed, model_output.output_token, model_output.logprobs,
                               self.enable_penalty, rand_host=rand_host)
        t3 = time.perf_counter()
        self.step_counter.choose_token_cost = int((t3 - t2) * 1e6)
        if rc != RC_SUCCESS:
            return RC_OTHER_ERROR, "SampleTopKTopP failed: " + self.lib.b2llm_last_error().decode()
        self.step_counter.output_token_cnt = running_batch
        return RC_SUCCESS, ""

    def logits(self, batch: int) -> np.ndarray:
        """fp32 [batch, vocab] host copy of the last forward's logits (parity tests)."""
        out = np.empty((batch, self.res.desc.vocab_size), dtype=np.float32)
        capi.check(self.lib.b2llm_engine_debug_read(self.res.engine, 3, _ptr(out), out.nbytes), "debug_read")
        return out

    def debug_read(self, what: int, shape, dtype) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        capi.check(self.lib.b2llm_engine_debug_read(self.res.engine, what, _ptr(out), out.nbytes), "debug_read")
        return out
