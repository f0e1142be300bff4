#!/usr/bin/env python
"""bench.py -- decode tokens/sec of the B200-native ppl.llm.serving hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the CPU arm (oracle port; the reference
                                                           # has neither kernels nor a CPU backend)

Workload (config.workload): BASELINE.json configs[1] "LLaMA-2-7B W8A8, running-batch 1024, seq 2048,
TP=1" at the largest uniform KV length that fits ONE B200 with the reference's own KV budget formula
(resource_manager.cc:329-342, --max-tokens-scale 0.94): the literal 1024 x 2048 needs 687 GB of int8
KV.  A step = one decode forward of 1024 running sequences (each attending to kv_len
cached tokens through a shuffled page table) + greedy sampling.  N > 1 runs N independent replicas
(requests are independent; the reference's only sharding is TP) -> "scaling": "weak".

One JSON line on stdout (rank 0).  `value` = device-timed steps with inputs resident in HBM;
`e2e` = the same steps through LLMEngine.Execute with host ModelInput vectors (H2D of the step
inputs + D2H of tokens/logprobs inside the timed region).
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
L2_BYTES = 126 * 2 ** 20


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


# ----------------------------------------------------------------------------------------------
def cpu_step_time(batch, kv_len, layers, seed=0):
    """seconds for the oracle (numpy, BLAS threads = all cores) to run `layers` transformer layers +
    head of a LLaMA-2-7B W8A8 decode step for `batch` sequences of `kv_len` cached tokens.
    The ONLY place bench.py executes oracle/ (cpu_baseline leg and --impl reference)."""
    from oracle import llama_ref as ref
    from oracle.weights import ModelDesc, SynthWeights
    desc = ModelDesc(4096, 11008, layers, 32, 32, 32000, cache_layout=3, cache_mode=1, page_size=PAGE,
                     quant_method=1, max_position=max(4096, kv_len + 1))
    w = SynthWeights(desc, 0xB200)
    rng = np.random.default_rng(seed)
    # weight VALUES do not affect timing: fill the oracle's weight cache directly instead of running the
    # (slow, single-threaded) reproducible hash generator over 0.7 G elements
    h, inter, D = desc.hidden_dim, desc.intermediate_dim, desc.head_dim
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


def run_cpu(batch, kv_len, layers, steps, warmup):
    orc, step = cpu_step_time(batch, kv_len, layers)
    from oracle import sampler_ref
    t_layers, t_total = [], []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        logits = orc.forward(step)
        sampler_ref.sample_topk_topp(logits, None, None, None, 32000, 1, 0.0)
        t1 = time.perf_counter()
        if i >= warmup:
            t_total.append(t1 - t0)
    return float(np.median(t_total))


def cpu_tokens_per_s(batch, kv_len, steps=2, warmup=1):
    """measure 1 layer and 2 layers -> per-layer time and head time -> 32-layer step time."""
    t1 = run_cpu(batch, kv_len, 1, steps, warmup)
    t2 = run_cpu(batch, kv_len, 2, steps, warmup)
    per_layer = max(t2 - t1, 1e-9)
    head = max(t1 - per_layer, 0.0)
    step_time = 32 * per_layer + head
    return batch / step_time, step_time, per_layer, head


def reference_arm(args, kv_len):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    batch = 32
    tps, step_time, per_layer, head = cpu_tokens_per_s(batch, kv_len, steps=max(1, args.steps), warmup=max(1, min(args.warmup, 1)))
    line = {
        "impl": "reference", "metric": METRIC, "value": tps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_time * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int8", "data": "synthetic",
        "config": {"workload": f"LLaMA-2-7B W8A8 TP=1, running batch {BATCH}, uniform kv_len {kv_len} (the length the b200 arm fits on "
                               f"one 180 GB B200 at max_tokens_scale {MAX_TOKENS_SCALE}), int8 group-8 paged KV page_size {PAGE} layout 3, "
                               f"greedy; CPU arm measured on a batch-{batch} sample of this step (see cpu_baseline.sample)"},
        "cpu_baseline": {"value": tps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"batch {batch} of {BATCH}, 1- and 2-layer runs timed, per-layer {per_layer * 1e3:.1f} ms x 32 + head "
                                   f"{head * 1e3:.1f} ms; numpy/BLAS on all {cores} host threads; the reference has no CPU "
                                   f"backend and no kernels in-tree (SURVEY.md F1/F3), so the builder-written oracle is timed"},
        "e2e": {"value": tps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--kv-len", type=int, default=0, help="override the fitted uniform KV length")
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
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
    from ppl_llm_serving_b200 import capi
    from ppl_llm_serving_b200.engine import (CudaResourceManager, LLMEngine, ModelConfig, ModelInput, ModelOutput,
                                             LLAMA2_7B, RC_SUCCESS, INT64_MAX, _ptr)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE {world}"
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        # no device_id: the NCCL communicator is created lazily at the first collective (the barrier before the timed
        # region), i.e. AFTER the KV budget has been taken from the free memory -- with eager creation NCCL's buffers eat
        # the 0.7 GB of slack that decides between kv_len 512 and 496, and the N > 1 runs would time a different workload
        dist.init_process_group("nccl")

    cfg = ModelConfig(**LLAMA2_7B, page_size=PAGE, max_position=4096)
    cfg.num_layers = args.layers
    res = CudaResourceManager()
    rc = res.Init(cfg, MAX_TOKENS_SCALE, max_running_batch=BATCH, max_tokens_per_step=BATCH, enable_penalty=False,
                  kv_cache_max_tokens=args.kv_budget_tokens or None, seed=0xB200, device=local_rank)
    if rc != RC_SUCCESS:
        raise SystemExit(f"engine init failed: {capi.load_library().b2llm_last_error().decode()}")
    lib = res.lib
    max_tokens = res.kv_cache_max_tokens
    kv_len = args.kv_len or min(2048, (max_tokens // BATCH) // PAGE * PAGE)
    pages_per = kv_len // PAGE
    assert pages_per * PAGE * BATCH <= max_tokens
    # synthetic cache contents (values do not affect timing; scales finite)
    res.kv_cache_mem.random_(-127, 128)
    res.kv_scale_mem.fill_(0.01)
    rng = np.random.default_rng(1002)
    perm = rng.permutation(BATCH * pages_per)
    page_list = (perm.reshape(BATCH, pages_per) * PAGE).astype(np.int64)

    engine = LLMEngine(res, False, 1, 0.0)
    mi = ModelInput()
    mi.token_inputs = rng.integers(0, cfg.vocab_size, BATCH).astype(np.int64)
    mi.seq_starts = np.arange(BATCH + 1, dtype=np.int64)
    mi.start_pos = np.full(BATCH, kv_len - 1, dtype=np.int64)
    mi.kv_starts = np.arange(BATCH + 1, dtype=np.int64) * kv_len
    mi.page_list = page_list.reshape(-1)
    mi.max_pages, mi.decoding_batches, mi.max_seq_len, mi.max_kv_len = pages_per, BATCH, 1, kv_len
    mi.temperatures = np.ones(BATCH, np.float32)
    mi.top_p_list = np.zeros(BATCH, np.float32)
    mi.top_k_list = [1] * BATCH
    out = ModelOutput()
    out.Resize(BATCH)
    # the step's host inputs live in pinned memory (what the e2e leg copies from every step); pageable as a fallback
    pinned_keep, pinned = [], True
    for name in ("token_inputs", "seq_starts", "start_pos", "kv_starts", "page_list", "temperatures", "top_p_list"):
        try:
            t = torch.from_numpy(np.ascontiguousarray(getattr(mi, name))).pin_memory()
            pinned_keep.append(t)
            setattr(mi, name, t.numpy())
        except Exception:
            pinned = False

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    stream = res.stream
    sptr = C.c_void_p(stream.cuda_stream)
    dev_tok = torch.empty(BATCH, dtype=torch.int32, device="cuda")
    dev_lp = torch.empty(BATCH, dtype=torch.float32, device="cuda")

    def device_step():
        """inputs already staged in HBM: forward + greedy sampler, nothing crosses PCIe"""
        rc = engine.RunModel(False)
        assert rc == RC_SUCCESS, lib.b2llm_last_error()
        rc = lib.b2llm_sample_topk_topp(sptr, C.c_void_p(engine.logits_ptr), None, None, None, BATCH, cfg.vocab_size,
                                        engine.logits_stride, 1, 0.0, 0.0, None, _ptr(dev_tok), _ptr(dev_lp))
        assert rc == RC_SUCCESS

    # stage inputs once for the device-resident measurement
    assert engine.SetInput(mi, True) == RC_SUCCESS
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        device_step()
    lib.b2llm_engine_profile(res.engine, 1)
    barrier()
    sampler.t_begin = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        device_step()
    ev1.record(stream)
    barrier()
    sampler.stop_flag = True
    ms_dev = ev0.elapsed_time(ev1)
    launches_per_step = lib.b2llm_engine_last_launch_count(res.engine) + 1
    ms_cls = (C.c_double * 3)()
    n_cls = (C.c_int64 * 3)()
    lib.b2llm_engine_profile_read(res.engine, ms_cls, n_cls, 3)
    lib.b2llm_engine_profile(res.engine, 0)

    # e2e: the public call a user makes, host vectors in, host tokens out
    for _ in range(2):
        rc, err = engine.Execute(mi, False, False, out)
        assert rc == RC_SUCCESS, err
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rc, err = engine.Execute(mi, False, False, out)
        assert rc == RC_SUCCESS, err
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3
    h2d = int(mi.token_inputs.nbytes + mi.seq_starts.nbytes + mi.kv_starts.nbytes + mi.start_pos.nbytes + BATCH * 4)
    d2h = BATCH * 8

    if world > 1:
        t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e = float(t[0]), float(t[1])
    value = world * BATCH * args.steps / (ms_dev * 1e-3)
    e2e = world * BATCH * args.steps / (ms_e2e * 1e-3)

    if rank == 0:
        hbm_peak, peak_src = peaks()
        kv_b_tok_layer = 2 * cfg.num_kv_heads * cfg.head_dim * (1 + 2 / cfg.cache_quant_group)  # int8 + fp16 scale / 8
        attn_bytes = BATCH * kv_len * kv_b_tok_layer          # algorithmic bytes of one attention launch
        attn_ms = ms_cls[0] / max(1, n_cls[0])
        achieved = attn_bytes / (attn_ms * 1e-3) / 1e9 if attn_ms > 0 else 0.0
        w_bytes = cfg.num_layers * (3 * cfg.hidden_dim * cfg.hidden_dim + cfg.hidden_dim * cfg.hidden_dim
                                    + 3 * cfg.hidden_dim * cfg.intermediate_dim) + cfg.vocab_size * cfg.hidden_dim * 2
        step_bytes = w_bytes + cfg.num_layers * attn_bytes
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
                "device_ms_by_class_per_step": {"attention": ms_cls[0] / args.steps, "layer_gemms": ms_cls[1] / args.steps,
                                                "lm_head": ms_cls[2] / args.steps},
            },
            "roofline": {"kernel": "attn_decode_kernel", "bound": "hbm", "achieved": achieved, "peak": hbm_peak,
                         "unit": "GB/s", "frac": achieved / hbm_peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full capture of this
                         # exact shape (profiles/r1_ncu_attn.txt); null for any other shape
                         "traffic": 5585427456 if (kv_len == 512 and cfg.num_layers == 32) else None,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": attn_bytes, "avg_launch_ms": attn_ms, "launches_timed": int(n_cls[0])},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps, "host_inputs": "pinned" if pinned else "pageable"},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": sampler.summary(),
        }
        if not args.no_cpu and world == 1:  # the CPU leg runs on rank 0 at N = 1 only
            cores = os.cpu_count()
            cb = 32
            tps, st, per_layer, head = cpu_tokens_per_s(cb, kv_len)
            line["cpu_baseline"] = {
                "value": tps, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"oracle (numpy/BLAS, all {cores} host threads) on batch {cb} of {BATCH} at kv_len {kv_len}: 1- and 2-layer "
                          f"runs timed, per-layer {per_layer * 1e3:.1f} ms x 32 + head {head * 1e3:.1f} ms = {st * 1e3:.0f} ms/step; "
                          f"builder-written port (the reference has no CPU backend, SURVEY.md F3)"}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    res.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
