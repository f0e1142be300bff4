import sys, ctypes as C
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import b200_import; b200_import.load()
from oracle.weights import ModelDesc
from ppl_llm_serving_b200 import capi
from ppl_llm_serving_b200.engine import CudaResourceManager, LLMEngine, ModelInput, ModelOutput, RC_SUCCESS

def run(B, KV, layers=1, hidden=4096, inter=11008, heads=32):
    PAGE = 16
    desc = ModelDesc(hidden, inter, layers, heads, heads, 32000, cache_layout=3, cache_mode=1, page_size=PAGE, quant_method=1, max_position=1024)
    T = B * KV
    res = CudaResourceManager()
    assert res.Init(desc, 0.9, B, B, kv_cache_max_tokens=T, seed=0xB200) == RC_SUCCESS
    engine = LLMEngine(res, False, 1, 0.0)
    g = torch.Generator(device="cuda").manual_seed(1)
    H, D = heads, 128
    logical = torch.randint(-127, 128, (layers * 2 * H, T, D), dtype=torch.int8, device="cuda", generator=g)
    logical_s = (torch.rand((layers * 2 * H, T, D // 8), device="cuda", generator=g) * 0.02 + 0.001).to(torch.float16)
    tokens = np.random.default_rng(2).integers(0, desc.vocab_size, B).astype(np.int64)
    outs = []
    for seed in (10, 10, 11):
        perm = np.random.default_rng(seed).permutation(T // PAGE)
        page_list = (perm.reshape(B, KV // PAGE) * PAGE).astype(np.int64)
        slot = torch.from_numpy((page_list[:, :, None] + np.arange(PAGE)[None, None, :]).reshape(-1)).cuda()
        cache = res.kv_cache_mem.view(layers * 2 * H, T, D)
        scale = res.kv_scale_mem.view(layers * 2 * H, T, D // 8)
        cache.index_copy_(1, slot, logical)
        scale.index_copy_(1, slot, logical_s)
        torch.cuda.synchronize()
        mi = ModelInput(token_inputs=tokens, seq_starts=np.arange(B + 1, dtype=np.int64), kv_starts=np.arange(B + 1, dtype=np.int64) * KV,
                        start_pos=np.full(B, KV - 1, dtype=np.int64), page_list=page_list.reshape(-1), max_pages=KV // PAGE,
                        decoding_batches=B, max_seq_len=1, max_kv_len=KV, temperatures=[1.0] * B, top_p_list=[0.0] * B, top_k_list=[1] * B)
        out = ModelOutput(); out.Resize(B)
        rc, err = engine.Execute(mi, True, False, out)
        assert rc == RC_SUCCESS, err
        attn = engine.debug_read(2, (B, H * D), np.float16).astype(np.float32)
        qkv = engine.debug_read(1, (B, 3 * H * D), np.float16).astype(np.float32)
        outs.append((engine.logits(B), attn, qkv))
    for name, i, j in (("same placement twice", 0, 1), ("different placement", 0, 2)):
        dl = np.abs(outs[i][0] - outs[j][0]).max(); da = np.abs(outs[i][1] - outs[j][1]).max(); dq = np.abs(outs[i][2] - outs[j][2]).max()
        rows = np.nonzero(np.abs(outs[i][1] - outs[j][1]).max(axis=1) > 0)[0]
        print(f"B={B} KV={KV} L={layers} h={hidden}: {name}: max|dlogits| {dl:.3e} max|dattn| {da:.3e} max|dqkv| {dq:.3e} bad rows {len(rows)} {rows[:8]}", flush=True)
    res.close()

run(8, 32, hidden=512, inter=1024, heads=4)
run(64, 128, hidden=512, inter=1024, heads=4)
run(64, 512)
run(1024, 64)
run(1024, 512)
