#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/.

The reference holds no golden vectors for this path and its arithmetic lives in
un-vendored ppl.nn @ master, so these fixtures are produced by the oracle (oracle/) and serve to
(a) freeze the oracle -- tests/test_oracle_cpu.py recomputes them on every CPU run, so a silent
change of a numeric convention shows up as a diff -- and (b) give the GPU tests a fixed
input/output pair that does not depend on the oracle being importable.

Where the reference's own sources compile here (host-side integer logic: HashCombine,
PrefixCacheManager; see oracle/ref_build.sh) the fixture host_kat.json additionally stores the
output of the compiled reference test (oracle/_ref/test_prefix_cache_mgr) so that the GPU box,
which has no /root/reference, still checks against reference-produced numbers.

Run:  python scripts/make_golden.py
"""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import host_ref, llama_ref as ref, sampler_ref  # noqa: E402
from oracle.weights import ModelDesc, SynthWeights, synth_tensor  # noqa: E402

OUT = ROOT / "tests" / "golden"


def small_desc(quant_method=1, layout=3, mode=1, kvh=4, kvbit=8):
    return ModelDesc(256, 512, 2, 4, kvh, 512, cache_layout=layout, cache_mode=mode, page_size=16,
                     quant_method=quant_method, max_position=128, cache_quant_bit=kvbit, cache_quant_group=8 if kvbit else 1)


def gen_step_fixture(name, desc):
    """prefill of a ragged batch then two decode steps; stores inputs, logits and greedy tokens."""
    rng = np.random.default_rng(7)
    w = SynthWeights(desc, 0xB200)
    orc = ref.LlamaOracle(desc, w, 256)
    prompts = [list(map(int, rng.integers(0, desc.vocab_size, n))) for n in (5, 17, 9)]
    pages = [[32, 0], [96, 16], [64, 48]]
    idx = [0, 40, 100]
    kw = dict(page_tables=pages) if desc.cache_mode == 1 else dict(cache_indices=idx)
    step = ref.build_step(desc, prompts, [0, 0, 0], 0, **kw)
    pos = [len(p) for p in prompts]
    rec = {}
    for it in range(3):
        logits = orc.forward(step)
        tok, lp = sampler_ref.sample_topk_topp(logits, None, None, None, desc.vocab_size, 1, 0.0)
        rec[f"s{it}_token_inputs"] = step.token_inputs
        rec[f"s{it}_seq_starts"] = step.seq_starts
        rec[f"s{it}_kv_starts"] = step.kv_starts
        rec[f"s{it}_start_pos"] = step.start_pos
        rec[f"s{it}_decoding_batches"] = np.int64(step.decoding_batches)
        rec[f"s{it}_logits"] = logits.astype(np.float32)
        rec[f"s{it}_tokens"] = tok
        rec[f"s{it}_logprobs"] = lp
        step = ref.build_step(desc, [[int(t)] for t in tok], pos, 3, **kw)
        pos = [p + 1 for p in pos]
    if desc.cache_mode == 1:
        rec["page_list"] = ref.build_step(desc, prompts, [0, 0, 0], 0, **kw).page_list
        rec["max_pages"] = np.int64(2)
    else:
        rec["cache_indices"] = np.asarray(idx, np.int64)
    c, s = orc.cache.export()
    rec["kv_cache_final"], rec["kv_scale_final"] = c, s
    rec["desc"] = np.asarray([desc.hidden_dim, desc.intermediate_dim, desc.num_layers, desc.num_heads, desc.num_kv_heads,
                              desc.vocab_size, desc.cache_layout, desc.cache_mode, desc.page_size, desc.quant_method,
                              desc.max_position, desc.cache_quant_bit, desc.cache_quant_group], np.int64)
    np.savez_compressed(OUT / f"{name}.npz", **rec)


def gen_ops_fixture():
    rng = np.random.default_rng(11)
    rec = {}
    rec["synth_t5"] = synth_tensor(0xB200, 5, (3, 64), 0.02)
    x = (rng.standard_normal((6, 256)) * 1.5).astype(np.float16)
    g = (1 + 0.02 * rng.standard_normal(256)).astype(np.float16)
    y = ref.rmsnorm_f32(x, g, 1e-5)
    q, s = ref.quant_rows(y)
    rec.update(rms_x=x, rms_g=g, rms_q=q, rms_s=s)
    a8 = rng.integers(-127, 128, (5, 96), dtype=np.int8)
    w8 = rng.integers(-127, 128, (24, 96), dtype=np.int8)
    rec.update(gemm_a=a8, gemm_w=w8, gemm_acc=ref.gemm_i8_acc_numpy(a8, w8))
    kx = (rng.standard_normal((4, 2, 128))).astype(np.float16)
    kq, ks = ref.kv_quant(kx, 8)
    rec.update(kv_x=kx, kv_q=kq, kv_s=ks)
    cos, sin = ref.rope_table(32, 128, 10000.0)
    rec.update(rope_out=ref.apply_rope(kx, np.array([0, 3, 17, 31]), cos, sin), rope_cos_31=cos[31], rope_sin_31=sin[31])
    logits = rng.standard_normal((4, 300)).astype(np.float32) * 3
    rnd = np.array([0.1, 0.5, 0.9, 0.3], np.float32)
    tok, lp = sampler_ref.sample_topk_topp(logits, np.array([0.7, 1.0, 1.3, 0.5], np.float32),
                                           np.array([0.9, 0.5, 1.0, 0.0], np.float32), rnd, 300, 8, 0.0)
    rec.update(samp_logits=logits, samp_rand=rnd, samp_tok=tok, samp_lp=lp)
    np.savez_compressed(OUT / "ops.npz", **rec)


def gen_host_kat():
    kat = {
        "hash_combine_0_12345": host_ref.hash_combine(0, [1, 2, 3, 4, 5]),
        "hash_combine_chain": host_ref.hash_combine(host_ref.hash_combine(0, list(range(16))), list(range(16, 32))),
        "hash_combine_negative": host_ref.hash_combine(12345, [-1, -2, 2147483647, -2147483648]),
        "page_count_examples": [[p, g, ps, host_ref.page_count(p, g, ps)] for p, g, ps in
                                [(1, 1, 16), (16, 1, 16), (16, 2, 16), (1000, 1024, 128), (17, 8, 16)]],
        "kv_budget_7b_178e9": list(host_ref.kv_cache_max_tokens(0.94, 178_000_000_000, 32, 32, 1, 4096, 32, 8, 8)),
        "kv_budget_70b_tp8_170e9": list(host_ref.kv_cache_max_tokens(0.94, 170_000_000_000, 80, 8, 8, 8192, 64, 8, 8)),
        "prefix_cache_sequence": None,
        "reference_test_output": None,
    }
    m = host_ref.PrefixCacheModel()
    for h, p in zip([0, 1, 2, 3], [11, 12, 13, 14]):
        m.insert(h, p)
    for h, p in zip([5, 6, 7, 8], [15, 16, 17, 18]):
        m.insert(h, p)
    m.dec_ref([0, 1, 2, 3])
    s1 = m.size()
    m.dec_ref([5, 6, 7, 8])
    e1 = m.evict(4)
    s2 = m.size()
    e2 = m.evict(4)
    kat["prefix_cache_sequence"] = {"size_after_inserts": s1, "evict1": e1, "size_after_evict1": s2, "evict2": e2,
                                    "size_end": m.size()}
    ref_bin = ROOT / "oracle" / "_ref" / "test_prefix_cache_mgr"
    if ref_bin.exists():
        out = subprocess.run([str(ref_bin)], capture_output=True, text=True, timeout=30).stdout
        kat["reference_test_output"] = out.strip().splitlines()
    (OUT / "host_kat.json").write_text(json.dumps(kat, indent=1) + "\n")


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    gen_ops_fixture()
    gen_step_fixture("step_w8a8_paged_l3", small_desc(1, 3, 1))
    gen_step_fixture("step_fp16_contig_l1_gqa", small_desc(0, 1, 0, kvh=2))
    gen_step_fixture("step_w8a8_fp16kv_paged_l2_gqa", small_desc(1, 2, 1, kvh=2, kvbit=0))   # cache_quant_bit 0: fp16 cache
    gen_host_kat()
    for f in sorted(OUT.iterdir()):
        print(f.name, f.stat().st_size)


if __name__ == "__main__":
    main()
