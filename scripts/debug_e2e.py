"""debug / calibration: GPU-vs-oracle logits error statistics and the oracle's own 1-ulp sensitivity."""
import sys
import numpy as np
sys.path.insert(0, ".")
import b200_import; b200_import.load()
from ppl_llm_serving_b200.engine import CudaResourceManager, LLMEngine, ModelInput, ModelOutput
from oracle import llama_ref as ref
from oracle.weights import ModelDesc, SynthWeights

def errs(a, b):
    mx = np.abs(a - b).max(axis=1) / np.abs(b).max(axis=1)
    l2 = np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)
    return mx, l2

def run(desc, prompt_len, steps, seed=5, kv=1024):
    w = SynthWeights(desc, seed)
    res = CudaResourceManager(); assert res.Init(desc, 0.9, 16, 256, kv_cache_max_tokens=kv, seed=seed) == 0
    eng = LLMEngine(res, False, 1, 0.0)
    orc = ref.LlamaOracle(desc, w, kv)
    orc2 = ref.LlamaOracle(desc, w, kv)
    rng = np.random.default_rng(seed)
    B = len(prompt_len)
    prompts = [list(map(int, rng.integers(0, desc.vocab_size, n))) for n in prompt_len]
    pages = [[64 * b + 16 * i for i in range(4)] for b in range(B)]
    step = ref.build_step(desc, prompts, [0] * B, 0, page_tables=pages)
    pos = list(prompt_len)
    for it in range(steps):
        mi = ModelInput(token_inputs=step.token_inputs.tolist(), seq_starts=step.seq_starts.tolist(), kv_starts=step.kv_starts.tolist(),
                        start_pos=step.start_pos.tolist(), page_list=step.page_list.tolist(), max_pages=step.max_pages,
                        decoding_batches=step.decoding_batches, max_seq_len=step.max_seq_len, max_kv_len=step.max_kv_len,
                        temperatures=[1.0] * B, top_p_list=[0.0] * B, top_k_list=[1] * B)
        out = ModelOutput(); out.Resize(B)
        rc, err = eng.Execute(mi, it == 0, False, out); assert rc == 0, err
        exp = orc.forward(step)
        nud = orc2.forward(step, ulp_nudge=True)
        got = eng.logits(B)
        mx, l2 = errs(got, exp)
        nmx, nl2 = errs(nud, exp)
        etok = exp.argmax(axis=1)
        print(f"step {it}: gpu max-norm {mx.max():.2e} (med {np.median(mx):.2e}) l2 {l2.max():.2e} (med {np.median(l2):.2e}) | "
              f"oracle 1-ulp nudge: max-norm {nmx.max():.2e} l2 {nl2.max():.2e} | tok match {int((out.output_token == etok).sum())}/{B}", flush=True)
        step = ref.build_step(desc, [[int(t)] for t in etok], pos, B, page_tables=pages)
        pos = [p + 1 for p in pos]
    res.close()

for qm in (0, 1):
    print("small dims 3 layers quant", qm)
    run(ModelDesc(512, 1024, 3, 4, 4, 1024, cache_layout=3, cache_mode=1, page_size=16, quant_method=qm, max_position=256), [5, 17, 1, 33, 8, 9, 20, 3], 4)
    print("7B dims 2 layers quant", qm)
    run(ModelDesc(4096, 11008, 2, 32, 32, 32000, cache_layout=3, cache_mode=1, page_size=16, quant_method=qm, max_position=256), [16, 3, 9, 30], 3, seed=0xB205)
