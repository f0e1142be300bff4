#!/usr/bin/env python
"""bench.py -- decode tokens/sec of the B200-native ppl.llm.serving hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the CPU arm (oracle port; the reference
                                                           # has neither kernels nor a CPU backend)

Workload (config.workload): BASELINE.json configs[1] "LLaMA-2-7B W8A8, running-batch 1024, seq 2048,
TP=1" at the largest uniform KV length that fits ONE B200 with the reference's own KV budget formula
(resource_manager.cc:329-342, --max-tokens-scale 0.94): the literal 1024 x 2048 needs 687 GB of int8
KV.  A step = one decode forward of 1024 running sequences (each attending to kv_len
cached tokens through a shuffled page table) + greedy sampling.

N > 1 reports two things in the one JSON line:
  * `value` / `e2e`: N independent replicas (requests are independent) -> "scaling": "weak";
  * `config.tp`: the reference's ONLY sharding, tensor parallelism over the N GPUs
    (--tensor-parallel-size N, resource_manager.cc:392-422, llm_engine.cc:124): a token-for-token parity
    gate against the oracle's TP restatement, then the same 7B step sharded TP = N (strong scaling), and
    BASELINE configs[2] (13B W8A8 TP=4) at N = 4 / configs[3] (70B GQA W4A16 TP=8) at N = 8, each with
    its all-reduce time per step and its fraction of the per-GPU HBM roofline.

One JSON line on stdout (rank 0).  `value` = device-timed steps with inputs resident in HBM (per-class
event profiling OFF; the class breakdown comes from a separate profiled pass); `e2e` = the same steps
through LLMEngine.Execute with host ModelInput vectors (H2D of the step inputs + D2H of tokens/logprobs
inside the timed region).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

if "reference" in sys.argv:
    # the CPU arm uses every host thread it can get: torchrun exports OMP_NUM_THREADS=1 to its workers, which halved
    # (and worse) the arm at N > 1 in round 1 -- set the thread pools explicitly, before numpy / BLAS load
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = str(os.cpu_count() or 1)

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))


def _baseline_metric():
    try:
        return json.loads((ROOT / "BASELINE.json").read_text())["metric"]
    except Exception:
        return "decode tokens/sec (whole box) LLaMA-7B W8A8 batch1024 seq2048 @1/2/4/8 B200"


METRIC = _baseline_metric()
UNIT = "tokens/s"
BATCH = 1024
PAGE = 16
MAX_TOKENS_SCALE = 0.94          # README.md:66 of the reference
CPU_SAMPLE_BATCH = 32
N_CLASSES = 4                    # b2llm_engine_profile classes: attention, layer GEMMs, lm_head, collectives


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.t_begin = index, [], False, None

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f + [time.perf_counter()])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        # samples taken inside the timed region; a region shorter than one nvidia-smi call falls back to the samples
        # taken under the identical warm-up load just before it
        inside = [s for s in self.samples if self.t_begin is not None and s[-1] >= self.t_begin]
        if inside:
            self.samples = inside
        sm = sorted(int(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(self.samples[0][1]), "reasons": reasons}


# ---------------------------------------------------------------------------------------------- CPU arm
def cpu_workload(batch, kv_len, layers, seed=0):
    """the oracle (numpy + its C int8 GEMM) set up for `layers` transformer layers + head of a LLaMA-2-7B W8A8
    decode step of `batch` sequences with `kv_len` cached tokens each.
    The ONLY place bench.py executes oracle/ for timing (cpu_baseline leg and --impl reference)."""
    from oracle import llama_ref as ref
    from oracle.weights import ModelDesc, SynthWeights
    desc = ModelDesc(4096, 11008, layers, 32, 32, 32000, cache_layout=3, cache_mode=1, page_size=PAGE,
                     quant_method=1, max_position=max(4096, kv_len + 1))
    w = SynthWeights(desc, 0xB200)
    rng = np.random.default_rng(seed)
    # weight VALUES do not affect timing: fill the oracle's weight cache directly instead of running the
    # (slow, single-threaded) reproducible hash generator over 0.7 G elements
    h, inter = desc.hidden_dim, desc.intermediate_dim
    for l in range(layers):
        lw = {"attn_norm": np.ones(h, np.float16), "ffn_norm": np.ones(h, np.float16)}
        for name, (n, k) in {"wqkv": (3 * h, h), "wo": (h, h), "wgate": (inter, h), "wup": (inter, h), "wdown": (h, inter)}.items():
            lw[name + "_q"] = rng.integers(-127, 128, (n, k), dtype=np.int8)
            lw[name + "_s"] = np.full(n, 2e-4, np.float32)
        w._cache[("layer", l)] = lw
    w._cache["emb"] = rng.standard_normal((desc.vocab_size, h), dtype=np.float32).astype(np.float16)
    w._cache["fn"] = np.ones(h, np.float16)
    w._cache["lm"] = (0.02 * rng.standard_normal((desc.vocab_size, h), dtype=np.float32)).astype(np.float16)
    pages_per = (kv_len + PAGE - 1) // PAGE
    orc = ref.LlamaOracle(desc, w, batch * pages_per * PAGE)
    orc.cache.cache[:] = rng.integers(-127, 128, orc.cache.cache.shape, dtype=np.int8)
    orc.cache.scale[:] = np.float16(0.01)
    perm = rng.permutation(batch * pages_per)
    tables = [list((perm[b * pages_per:(b + 1) * pages_per] * PAGE).astype(np.int64)) for b in range(batch)]
    toks = [[int(t)] for t in rng.integers(0, 32000, batch)]
    step = ref.build_step(desc, toks, [kv_len - 1] * batch, batch, page_tables=tables)
    return orc, step


def cpu_threads():
    """thread counts actually in effect for the BLAS / OpenMP pools of this process"""
    try:
        from threadpoolctl import threadpool_info
        return {f"{p.get('user_api')}:{p.get('internal_api')}": p.get("num_threads") for p in threadpool_info()}
    except Exception:
        return {}


def cpu_arm(kv_len, steps, warmup):
    """One CPU 'step' = a bounded SAMPLE of the workload's step that is actually executed and timed: the oracle's
    forward + greedy sampler for batch 32 of the 1024 sequences through a 1-layer and through a 2-layer LLaMA-2-7B
    (same kv_len, same paged int8 KV).  From the medians: per-layer = t2 - t1, head (embedding + final norm +
    lm_head + sampler) = 2 t1 - t2, and the 32-layer step of that batch = 32 per-layer + head (the extrapolation is
    linear in identical layers; it is reported as such, never as a measured step)."""
    from oracle import sampler_ref
    orc1, step1 = cpu_workload(CPU_SAMPLE_BATCH, kv_len, 1)
    orc2, step2 = cpu_workload(CPU_SAMPLE_BATCH, kv_len, 2)
    t1s, t2s = [], []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        sampler_ref.sample_topk_topp(orc1.forward(step1), None, None, None, 32000, 1, 0.0)
        ta = time.perf_counter()
        sampler_ref.sample_topk_topp(orc2.forward(step2), None, None, None, 32000, 1, 0.0)
        tb = time.perf_counter()
        if i >= warmup:
            t1s.append(ta - t0)
            t2s.append(tb - ta)
    t1, t2 = float(np.median(t1s)), float(np.median(t2s))
    per_layer = max(t2 - t1, 1e-9)
    head = max(t1 - per_layer, 0.0)
    full = 32 * per_layer + head
    cores = os.cpu_count()
    threads = cpu_threads()
    return {
        "tokens_per_s": CPU_SAMPLE_BATCH / full, "sample_step_ms": (t1 + t2) * 1e3, "full_step_ms_extrapolated": full * 1e3,
        "per_layer_ms": per_layer * 1e3, "head_ms": head * 1e3, "cores": cores, "threads": threads,
        "sample": (f"per step, actually run and timed: oracle (numpy/BLAS + C int8 GEMM) forward + greedy sampler of batch "
                   f"{CPU_SAMPLE_BATCH} of {BATCH} at kv_len {kv_len} through 1 and through 2 layers ({(t1 + t2) * 1e3:.0f} ms "
                   f"median over {steps} steps); per-layer {per_layer * 1e3:.1f} ms, head {head * 1e3:.1f} ms -> 32-layer step "
                   f"of that batch {full * 1e3:.0f} ms (linear extrapolation) -> tokens/s = {CPU_SAMPLE_BATCH} / that; one host "
                   f"({cores} logical cores, thread pools {threads}), NOT multiplied by the number of GPUs; builder-written port "
                   f"(the reference has no CPU backend and no kernels in-tree, SURVEY.md F1/F3)"),
    }


def reference_arm(args, kv_len):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_arm(kv_len, max(1, args.steps), max(0, min(args.warmup, 1)))
    line = {
        "impl": "reference", "metric": METRIC, "value": r["tokens_per_s"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["sample_step_ms"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int8", "data": "synthetic",
        "config": {"workload": f"LLaMA-2-7B W8A8 TP=1, running batch {BATCH}, uniform kv_len {kv_len} (the length the b200 arm fits on "
                               f"one 180 GB B200 at max_tokens_scale {MAX_TOKENS_SCALE}), int8 group-8 paged KV page_size {PAGE} layout 3, "
                               f"greedy; CPU arm: ms_per_step is the bounded sample that was run (see cpu_baseline.sample), value the "
                               f"tokens/s of the 32-layer step extrapolated from it; one host, not scaled by n_gpus",
                   "full_step_ms_extrapolated": r["full_step_ms_extrapolated"], "per_layer_ms": r["per_layer_ms"],
                   "head_ms": r["head_ms"]},
        "cpu_baseline": {"value": r["tokens_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["tokens_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- B200 arm
class DecodeRun:
    """One engine (rank `rank` of `tp`) + a synthetic steady-state decode step: `batch` running sequences, each with
    `kv_len` cached tokens behind a shuffled page table.  Mirrors what the generator hands LLMEngine::Execute."""

    def __init__(self, torch, cfg, batch, kv_len_target, device, tp=1, rank=0, comm=None, kv_budget_tokens=None, dist=None):
        from ppl_llm_serving_b200 import capi
        from ppl_llm_serving_b200.engine import CudaResourceManager, LLMEngine, ModelInput, ModelOutput, RC_SUCCESS
        self.torch, self.cfg, self.batch, self.tp = torch, cfg, batch, tp
        self.res = CudaResourceManager()
        rc = self.res.Init(cfg, MAX_TOKENS_SCALE, max_running_batch=batch, max_tokens_per_step=batch, enable_penalty=False,
                           kv_cache_max_tokens=kv_budget_tokens, seed=0xB200, device=device, tensor_parallel_size=tp,
                           rank=rank, nccl_comm=comm)
        self.ok = rc == RC_SUCCESS
        self.err = "" if self.ok else capi.load_library().b2llm_last_error().decode()
        if dist is not None and tp > 1:  # every rank must have come up before anyone enters a collective
            flag = torch.tensor([1 if self.ok else 0], dtype=torch.int32, device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0 and self.ok:
                self.ok, self.err = False, "engine init failed on another rank"
        if not self.ok:
            return
        self.lib = self.res.lib
        self.max_tokens = self.res.kv_cache_max_tokens
        self.kv_len = min(kv_len_target, (self.max_tokens // batch) // PAGE * PAGE)
        pages_per = self.kv_len // PAGE
        assert pages_per > 0 and pages_per * PAGE * batch <= self.max_tokens
        # synthetic cache contents (values do not affect timing; scales finite)
        self.res.kv_cache_mem.random_(-127, 128)
        self.res.kv_scale_mem.fill_(0.01)
        rng = np.random.default_rng(1002)
        perm = rng.permutation(batch * pages_per)
        mi = ModelInput()
        mi.token_inputs = rng.integers(0, cfg.vocab_size, batch).astype(np.int64)
        mi.seq_starts = np.arange(batch + 1, dtype=np.int64)
        mi.start_pos = np.full(batch, self.kv_len - 1, dtype=np.int64)
        mi.kv_starts = np.arange(batch + 1, dtype=np.int64) * self.kv_len
        mi.page_list = (perm.reshape(batch, pages_per) * PAGE).astype(np.int64).reshape(-1)
        mi.max_pages, mi.decoding_batches, mi.max_seq_len, mi.max_kv_len = pages_per, batch, 1, self.kv_len
        mi.temperatures = np.ones(batch, np.float32)
        mi.top_p_list = np.zeros(batch, np.float32)
        mi.top_k_list = [1] * batch
        # the step's host inputs live in pinned memory (what the e2e leg copies from every step); pageable as a fallback
        self._pinned_keep, self.pinned = [], True
        for name in ("token_inputs", "seq_starts", "start_pos", "kv_starts", "page_list", "temperatures", "top_p_list"):
            try:
                t = torch.from_numpy(np.ascontiguousarray(getattr(mi, name))).pin_memory()
                self._pinned_keep.append(t)
                setattr(mi, name, t.numpy())
            except Exception:
                self.pinned = False
        self.mi = mi
        self.out = ModelOutput()
        self.out.Resize(batch)
        self.engine = LLMEngine(self.res, False, 1, 0.0)
        self.stream = self.res.stream
        self._sptr = C.c_void_p(self.stream.cuda_stream)
        self._dev_tok = torch.empty(batch, dtype=torch.int32, device="cuda")
        self._dev_lp = torch.empty(batch, dtype=torch.float32, device="cuda")
        assert self.engine.SetInput(mi, True) == RC_SUCCESS, self.lib.b2llm_last_error()

    def device_step(self):
        """inputs already staged in HBM: forward + greedy sampler, nothing crosses PCIe"""
        from ppl_llm_serving_b200.engine import _ptr
        rc = self.engine.RunModel(False)
        assert rc == 0, self.lib.b2llm_last_error()
        rc = self.lib.b2llm_sample_topk_topp(self._sptr, C.c_void_p(self.engine.logits_ptr), None, None, None, self.batch,
                                             self.cfg.vocab_size, self.engine.logits_stride, 1, 0.0, 0.0, None,
                                             _ptr(self._dev_tok), _ptr(self._dev_lp))
        assert rc == 0, self.lib.b2llm_last_error()

    def time_device(self, steps, barrier):
        """K steps, CUDA events on the engine's stream, profiling off -> ms for the K steps"""
        torch = self.torch
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(self.stream)
        for _ in range(steps):
            self.device_step()
        ev1.record(self.stream)
        barrier()
        return ev0.elapsed_time(ev1)

    def profile_classes(self, steps):
        """separate pass with per-class CUDA events (they add ~1 % to the step): ms per step and launches per step by class"""
        torch = self.torch
        self.lib.b2llm_engine_profile(self.res.engine, 1)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(self.stream)
        for _ in range(steps):
            self.device_step()
        ev1.record(self.stream)
        ms = (C.c_double * N_CLASSES)()
        n = (C.c_int64 * N_CLASSES)()
        self.lib.b2llm_engine_profile_read(self.res.engine, ms, n, N_CLASSES)
        self.lib.b2llm_engine_profile(self.res.engine, 0)
        self.profiled_ms_per_step = ev0.elapsed_time(ev1) / steps   # the class shares refer to THIS pass (events cost ~1-3 %)
        return [ms[i] / steps for i in range(N_CLASSES)], [n[i] / steps for i in range(N_CLASSES)]

    def time_e2e(self, steps, barrier):
        """the public call a user makes: host vectors in, host tokens out; wall clock around K Execute() calls"""
        for _ in range(2):
            rc, err = self.engine.Execute(self.mi, False, False, self.out)
            assert rc == 0, err
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            rc, err = self.engine.Execute(self.mi, False, False, self.out)
            assert rc == 0, err
        self.torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3

    def launches_per_step(self):
        return int(self.lib.b2llm_engine_last_launch_count(self.res.engine)) + 1  # + sampler

    def bytes_per_gpu(self):
        """ALGORITHMIC HBM bytes of one step on ONE GPU of the TP group: this rank's weight slice
        read once, its lm_head slice, and its kv heads of every cached token (int8 values + fp16 group scales)."""
        c, tp = self.cfg, self.tp
        q = c.quant_method
        wbytes = {0: 2.0, 1: 1.0, 2: 0.5 + 2.0 / 128}[q]
        D = c.head_dim
        lin = (c.num_heads + 2 * c.num_kv_heads) * D * c.hidden_dim + c.hidden_dim * c.hidden_dim + 3 * c.hidden_dim * c.intermediate_dim
        weights = c.num_layers * lin * wbytes / tp + c.vocab_size * c.hidden_dim * 2 / tp
        kv_tok_layer = 2 * c.num_kv_heads * D * (1 + 2 / c.cache_quant_group) / tp
        attn_launch = self.batch * self.kv_len * kv_tok_layer
        return weights + c.num_layers * attn_launch, attn_launch

    def close(self):
        if getattr(self, "res", None) is not None:
            self.res.close()
            for a in ("kv_cache_mem", "kv_scale_mem", "post_processor"):
                if hasattr(self.res, a):
                    delattr(self.res, a)
        self.res = None
        self.engine = None
        self.torch.cuda.empty_cache()


def tp_parity_gate(torch, dist, world, rank, local_rank, comm, quant):
    """2-layer model on the TP group: prefill + 3 greedy decode steps, every rank's tokens must equal the tokens of the
    oracle's TP restatement LlamaOracle(tp = world) and the logits agree within 1e-3 of the row's max |logit| (W8A8 rows
    may sit on a one-ulp re-quantisation flip: bound 5e-2, median 1e-3 -- tests/test_engine_gpu.py).  The oracle is the
    CHECKER here, run on rank 0 only."""
    from oracle import llama_ref as ref
    from oracle import sampler_ref
    from oracle.weights import ModelDesc, SynthWeights
    from ppl_llm_serving_b200.engine import CudaResourceManager, LLMEngine, ModelInput, ModelOutput, RC_SUCCESS
    heads = max(8, world)
    desc = ModelDesc(heads * 128, 256 * max(8, world), 2, heads, max(world, heads // 2), 2048, cache_layout=3, cache_mode=1,
                     page_size=16, quant_method=quant, max_position=256)
    res = CudaResourceManager()
    rc = res.Init(desc, 0.9, 8, 128, kv_cache_max_tokens=512, seed=0xB200, device=local_rank, tensor_parallel_size=world,
                  rank=rank, nccl_comm=comm)
    assert rc == RC_SUCCESS, res.lib.b2llm_last_error()
    eng = LLMEngine(res, False, 1, 0.0)
    rng = np.random.default_rng(0)
    prompts = [list(map(int, rng.integers(0, desc.vocab_size, n))) for n in (7, 19, 33)]
    pages = [[0, 16, 32], [48, 64, 80], [96, 112, 128]]
    step = ref.build_step(desc, prompts, [0, 0, 0], 0, page_tables=pages)
    pos = [len(p) for p in prompts]
    toks, logits = [], []
    for it in range(4):
        mi = ModelInput(token_inputs=step.token_inputs.tolist(), seq_starts=step.seq_starts.tolist(),
                        kv_starts=step.kv_starts.tolist(), start_pos=step.start_pos.tolist(),
                        page_list=step.page_list.tolist(), max_pages=step.max_pages,
                        decoding_batches=step.decoding_batches, max_seq_len=step.max_seq_len, max_kv_len=step.max_kv_len,
                        temperatures=[1.0] * 3, top_p_list=[0.0] * 3, top_k_list=[1] * 3)
        out = ModelOutput()
        out.Resize(3)
        rc, err = eng.Execute(mi, it == 0, False, out)
        assert rc == RC_SUCCESS, err
        toks.append(out.output_token.copy())
        logits.append(eng.logits(3))
        step = ref.build_step(desc, [[int(t)] for t in out.output_token], pos, 3, page_tables=pages)
        pos = [p + 1 for p in pos]
    res.close()
    mine = torch.tensor(np.stack(toks).astype(np.int64), device="cuda")
    allt = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allt, mine)
    result = {"quant_method": {0: "none", 1: "online_i8i8", 2: "w4a16"}[quant], "tp": world, "steps": 4,
              "model": f"2 layers, hidden {desc.hidden_dim}, {desc.num_heads} heads / {desc.num_kv_heads} kv heads, vocab {desc.vocab_size}"}
    if rank == 0:
        orc = ref.LlamaOracle(desc, SynthWeights(desc, 0xB200), 512, tp=world)
        step = ref.build_step(desc, prompts, [0, 0, 0], 0, page_tables=pages)
        pos = [len(p) for p in prompts]
        ok_tok, worst, med = True, 0.0, 0.0
        for it in range(4):
            exp = orc.forward(step)
            etok, _ = sampler_ref.sample_topk_topp(exp, None, None, None, desc.vocab_size, 1, 0.0)
            for r in range(world):
                ok_tok &= allt[r][it].cpu().numpy().tolist() == etok.tolist()
            rel = np.abs(logits[it] - exp).max(axis=1) / np.abs(exp).max(axis=1)
            worst, med = max(worst, float(rel.max())), max(med, float(np.median(rel)))
            step = ref.build_step(desc, [[int(t)] for t in etok], pos, 3, page_tables=pages)
            pos = [p + 1 for p in pos]
        bound = 5e-2 if quant == 1 else 2e-3
        result.update({"tokens_match_oracle_all_ranks": bool(ok_tok), "logits_rel_err_max": worst, "logits_rel_err_median_max": med,
                       "passed": bool(ok_tok and worst <= bound and med <= 1e-3),
                       "oracle": "oracle.llama_ref.LlamaOracle(tp=N), PARITY UNPINNED (builder-written; SURVEY F1)"})
    flag = torch.tensor([1 if (rank != 0 or result.get("passed")) else 0], dtype=torch.int32, device="cuda")
    dist.broadcast(flag, src=0)
    result["passed"] = bool(int(flag.item()))
    return result


def tp_leg(args, torch, dist, world, rank, local_rank, hbm_peak):
    """tensor parallelism over all `world` GPUs of the box: the reference's --tensor-parallel-size (its only sharding)."""
    from ppl_llm_serving_b200 import nccl
    from ppl_llm_serving_b200.engine import ModelConfig, LLAMA2_7B, LLAMA2_13B, LLAMA2_70B
    comm = nccl.create_comm(world, rank)

    def barrier():
        dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    out = {"tensor_parallel_size": world,
           "collectives": "after o_proj and after down_proj of every layer (the two residual join points): one fused kernel over "
                          "NVLink peer memory = all-reduce of the fp16 partials + residual add + RMSNorm + int8 quant (ncclAllReduce + "
                          "separate kernels with B2LLM_TP_JOIN=nccl); + one ncclAllGather(fp32) of the vocab-parallel logits per step",
           "parity_gate": [], "runs": []}
    quants = [1, 2]   # W8A8 and W4A16 (the latter is timed at N = 8 only, but its kernel is gated at every N)
    for q in quants:
        try:
            out["parity_gate"].append(tp_parity_gate(torch, dist, world, rank, local_rank, comm, q))
        except Exception as ex:  # a failed gate is reported, the timing below still runs (and is then unvalidated)
            out["parity_gate"].append({"quant_method": q, "passed": False, "error": repr(ex)[:300]})
    shapes = [("LLaMA-2-7B W8A8", LLAMA2_7B, 1, BATCH, args.kv_len or 512,
               f"BASELINE configs[1] sharded TP={world}: the N=1 step (B=1024, kv_len 512) strong-scaled")]
    if world == 8:
        shapes.append(("LLaMA-2-7B W8A8", LLAMA2_7B, 1, BATCH, 2048, "BASELINE configs[1] at its LITERAL shape (1024 x 2048 fits at TP=8)"))
        shapes.append(("LLaMA-2-70B GQA W4A16", LLAMA2_70B, 2, 256, 8192, "BASELINE configs[3], literal shape"))
    if world == 4:
        shapes.append(("LLaMA-2-13B W8A8", LLAMA2_13B, 1, 512, 4096,
                       "BASELINE configs[2]: B=512 at the largest kv_len <= 4096 the reference's KV budget fits (literal needs 268 GB/GPU)"))
    if args.tp_filter:
        shapes = [sh for sh in shapes if args.tp_filter in sh[0]]
        out["shape_filter"] = args.tp_filter
    for name, dims, quant, batch, kv_target, note in shapes:
        cfg = ModelConfig(**dims, page_size=PAGE, max_position=max(4096, kv_target + 16), quant_method=quant)
        if args.layers != 32:
            cfg.num_layers = args.layers
        entry = {"model": name, "tp": world, "batch": batch, "note": note}
        run = None
        try:
            run = DecodeRun(torch, cfg, batch, kv_target, local_rank, tp=world, rank=rank, comm=comm, dist=dist,
                            kv_budget_tokens=args.kv_budget_tokens or None)
            if not run.ok:
                entry["error"] = run.err
            else:
                for _ in range(args.warmup):
                    run.device_step()
                js = (C.c_double * 8)()
                run.lib.b2llm_engine_tp_join_stats(run.res.engine, js)   # reset: count the timed loop only
                ms = run.time_device(args.steps, barrier)
                run.lib.b2llm_engine_tp_join_stats(run.res.engine, js)
                cls_ms, cls_n = run.profile_classes(min(args.steps, 5))
                ms_e2e = run.time_e2e(min(args.steps, 10), barrier) / min(args.steps, 10) * args.steps
                t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms, ms_e2e = float(t[0]) / args.steps, float(t[1]) / args.steps
                step_bytes, attn_bytes = run.bytes_per_gpu()
                other = run.profiled_ms_per_step - sum(cls_ms)
                entry.update({
                    "kv_len": run.kv_len, "kv_budget_tokens": run.max_tokens, "layers": cfg.num_layers,
                    "ms_per_step": ms, "tokens_per_s": batch / (ms * 1e-3), "e2e_tokens_per_s": batch / (ms_e2e * 1e-3),
                    "device_ms_by_class_per_step": {"attention": cls_ms[0], "layer_gemms": cls_ms[1], "lm_head": cls_ms[2],
                                                    "collectives": cls_ms[3], "other (norm/quant/rope/sampler + gaps)": other,
                                                    "profiled_pass_ms_per_step": run.profiled_ms_per_step,
                                                    "note": "separate pass with CUDA events around every kernel class"},
                    "allreduce_ms_per_step": cls_ms[3], "collective_calls_per_step": cls_n[3],
                    "collective_share_of_step": cls_ms[3] / ms,
                    "per_gpu_roofline": {"algorithmic_bytes_per_step_per_gpu": step_bytes,
                                         "hbm_bound_ms": step_bytes / (hbm_peak * 1e9) * 1e3,
                                         "frac_of_hbm_roofline": step_bytes / (hbm_peak * 1e9) * 1e3 / ms},
                    "attention_GBps_per_gpu": attn_bytes / (cls_ms[0] / max(1.0, cls_n[0]) * 1e-3) / 1e9 if cls_ms[0] > 0 else None,
                })
                if js[0] > 0:  # the fused join's own phase clocks (rank 0), microseconds per call in the timed loop
                    entry["fused_join_us_per_call"] = {"calls_per_step": js[0] / args.steps,
                                                       "waiting_for_peers_partials (rank skew)": js[1] / js[0] * 1e-3,
                                                       "reduce_norm_quant_deliver": js[2] / js[0] * 1e-3,
                                                       "waiting_for_peers_rows": js[3] / js[0] * 1e-3,
                                                       "first_row_of_cta0": {"peer_loads": js[5] / js[0] * 1e-3,
                                                                             "reduce_quant_issue_stores": js[6] / js[0] * 1e-3,
                                                                             "system_fence": js[7] / js[0] * 1e-3}}
                    entry["exchange"] = "fused all-reduce + residual + RMSNorm + quant kernel over NVLink peer memory (csrc/tp_join.cu)"
                else:
                    entry["exchange"] = "ncclAllReduce + separate residual / RMSNorm / quant kernels (B2LLM_TP_JOIN=nccl)"
        except Exception as ex:
            entry["error"] = repr(ex)[:400]
        finally:
            if run is not None:
                run.close()
        out["runs"].append(entry)
    ok_runs = [r for r in out["runs"] if "ms_per_step" in r]
    if ok_runs:
        worst = max(ok_runs, key=lambda r: r["collective_share_of_step"])
        out["limiting_collective"] = (f"residual join x {2 * worst['layers']} per step (+ logits all-gather): {worst['allreduce_ms_per_step']:.2f} ms = "
                                      f"{100 * worst['collective_share_of_step']:.0f} % of the {worst['model']} TP={world} step")
    nccl.destroy_comm(comm)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--kv-len", type=int, default=0, help="override the fitted uniform KV length")
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-tp", action="store_true", help="N > 1: skip the tensor-parallel leg")
    ap.add_argument("--no-alt", action="store_true", help="N = 1: skip the config-2b alternative shape")
    ap.add_argument("--tp-filter", default="", help="N > 1: time only the tensor-parallel shapes whose model name contains this "
                                                    "(e.g. 70B); the replica leg and the parity gates still run")
    ap.add_argument("--kv-budget-tokens", type=int, default=0,
                    help="allocate exactly this many KV tokens instead of the reference's 0.94 x free-memory budget "
                         "(profiling runs: ncu saves / restores all device memory on every replay pass)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        reference_arm(args, args.kv_len or 512)
        return

    # stdout carries exactly ONE JSON line: libraries that chat on fd 1 (NCCL prints its version there) go to stderr
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import b200_import
    b200_import.load()
    from ppl_llm_serving_b200.engine import ModelConfig, LLAMA2_7B

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE {world}"
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # no device_id: the NCCL communicator is created lazily at the first collective (the barrier before the timed
        # region), i.e. AFTER the KV budget has been taken from the free memory -- with eager creation NCCL's buffers eat
        # the 0.7 GB of slack that decides between kv_len 512 and 496, and the N > 1 runs would time a different workload
        dist.init_process_group("nccl")

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    cfg = ModelConfig(**LLAMA2_7B, page_size=PAGE, max_position=4096)
    cfg.num_layers = args.layers
    run = DecodeRun(torch, cfg, BATCH, args.kv_len or 2048, local_rank, kv_budget_tokens=args.kv_budget_tokens or None)
    if not run.ok:
        raise SystemExit(f"engine init failed: {run.err}")
    kv_len, max_tokens = run.kv_len, run.max_tokens

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        run.device_step()
    torch.cuda.synchronize()
    sampler.t_begin = time.perf_counter()
    ms_dev = run.time_device(args.steps, barrier)
    sampler.stop_flag = True
    launches_per_step = run.launches_per_step()
    cls_ms, cls_n = run.profile_classes(min(args.steps, 5))
    profiled_ms = run.profiled_ms_per_step
    ms_e2e = run.time_e2e(args.steps, barrier)
    mi = run.mi
    h2d = int(mi.token_inputs.nbytes + mi.seq_starts.nbytes + mi.kv_starts.nbytes + mi.start_pos.nbytes + BATCH * 4)
    d2h = BATCH * 8

    if world > 1:
        t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e = float(t[0]), float(t[1])
    value = world * BATCH * args.steps / (ms_dev * 1e-3)
    e2e = world * BATCH * args.steps / (ms_e2e * 1e-3)
    hbm_peak, peak_src = peaks()
    step_bytes, attn_bytes = run.bytes_per_gpu()

    # N = 1: SURVEY 8(d) config 2b on the same engine and KV memory -- the literal seq 2048 at the batch that fits
    alt = None
    if world == 1 and not args.no_alt and args.layers == 32 and not args.kv_len:
        alt = alt_shape_2b(torch, run, hbm_peak, barrier, max(3, min(args.steps, 6)))
    run.close()
    del run

    tp = None
    if world > 1 and not args.no_tp:
        tp = tp_leg(args, torch, dist, world, rank, local_rank, hbm_peak)

    if rank == 0:
        attn_ms = cls_ms[0] / max(1.0, cls_n[0])
        achieved = attn_bytes / (attn_ms * 1e-3) / 1e9 if attn_ms > 0 else 0.0
        ms_step = ms_dev / args.steps
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8", "data": "synthetic",
            "config": {
                "arithmetic": "W8A8: int8 x int8 -> int32 tensor-core GEMMs, int8 group-8 KV cache, fp16 activations, fp32 reductions",
                "workload": f"LLaMA-2-7B W8A8 TP=1, running batch {BATCH}, uniform kv_len {kv_len} "
                            f"(largest that fits: KV budget {max_tokens} tokens at max_tokens_scale {MAX_TOKENS_SCALE}; "
                            f"literal seq 2048 needs 687 GB), int8 group-8 paged KV page_size {PAGE} layout 3, greedy",
                "layers": cfg.num_layers, "replicas": world,
                "l2": "inputs larger than L2 (KV read per step %.1f GB >> 126 MB)" % (cfg.num_layers * attn_bytes / 1e9),
                "step_roofline": {"algorithmic_bytes_per_step": step_bytes,
                                  "hbm_bound_ms": step_bytes / (hbm_peak * 1e9) * 1e3,
                                  "frac_of_hbm_roofline": step_bytes / (hbm_peak * 1e9) * 1e3 / ms_step},
                "device_ms_by_class_per_step": {"attention": cls_ms[0], "layer_gemms": cls_ms[1], "lm_head": cls_ms[2],
                                                "other (norm/quant/rope/sampler + gaps)": profiled_ms - sum(cls_ms[:3]),
                                                "profiled_pass_ms_per_step": profiled_ms,
                                                "note": "separate profiled pass (CUDA events per kernel class), not the timed loop"},
                "parity": "unpinned: the oracle is builder-written (the reference holds no kernels and no golden vectors, SURVEY F1/F6)",
            },
            "roofline": {"kernel": "attn_decode_kernel", "bound": "hbm", "achieved": achieved, "peak": hbm_peak,
                         "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": None,  # not measured in this run; see traffic_reference_capture
                         "traffic_reference_capture": "profiles/r2_ncu_step_run25.txt: 5389960000 B dram read+write per launch at this shape "
                                                      "(ncu --set full, LOADER 3, round 2 run 25) vs 5368709120 algorithmic = 1.004x",
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": attn_bytes, "avg_launch_ms": attn_ms,
                         "launches_timed": int(cls_n[0] * min(args.steps, 5))},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps, "host_inputs": "pinned"},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": sampler.summary(),
        }
        if alt is not None:
            line["config"]["alt_shapes"] = [alt]
        if tp is not None:
            line["config"]["tp"] = tp
        if not args.no_cpu and world == 1:  # the CPU leg runs on rank 0 at N = 1 only
            r = cpu_arm(kv_len, 2, 1)
            line["cpu_baseline"] = {"value": r["tokens_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def alt_shape_2b(torch, run, hbm_peak, barrier, steps):
    """kv_len 2048 (BASELINE's literal seq) at the running batch that fits the same KV budget"""
    from ppl_llm_serving_b200.engine import ModelInput
    kv_len = 2048
    batch = min(BATCH, run.max_tokens // kv_len)
    if batch < 8:
        return None
    pages_per = kv_len // PAGE
    rng = np.random.default_rng(1003)
    perm = rng.permutation(batch * pages_per)
    mi = ModelInput()
    mi.token_inputs = rng.integers(0, run.cfg.vocab_size, batch).astype(np.int64)
    mi.seq_starts = np.arange(batch + 1, dtype=np.int64)
    mi.start_pos = np.full(batch, kv_len - 1, dtype=np.int64)
    mi.kv_starts = np.arange(batch + 1, dtype=np.int64) * kv_len
    mi.page_list = (perm.reshape(batch, pages_per) * PAGE).astype(np.int64).reshape(-1)
    mi.max_pages, mi.decoding_batches, mi.max_seq_len, mi.max_kv_len = pages_per, batch, 1, kv_len
    old = (run.mi, run.batch, run.kv_len)
    run.mi, run.batch, run.kv_len = mi, batch, kv_len
    try:
        assert run.engine.SetInput(mi, True) == 0, run.lib.b2llm_last_error()
        for _ in range(3):
            run.device_step()
        ms = run.time_device(steps, barrier) / steps
        cls_ms, cls_n = run.profile_classes(min(steps, 3))
        step_bytes, attn_bytes = run.bytes_per_gpu()
        return {"workload": f"config 2b: LLaMA-2-7B W8A8 TP=1, running batch {batch} (largest that fits), uniform kv_len {kv_len}",
                "ms_per_step": ms, "tokens_per_s": batch / (ms * 1e-3), "steps": steps,
                "frac_of_hbm_roofline": step_bytes / (hbm_peak * 1e9) * 1e3 / ms,
                "attention_frac_of_hbm_peak": attn_bytes / (cls_ms[0] / max(1.0, cls_n[0]) * 1e-3) / 1e9 / hbm_peak}
    finally:
        run.mi, run.batch, run.kv_len = old


if __name__ == "__main__":
    main()
