import sys, ctypes as C
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import b200_import; b200_import.load()
from oracle.weights import ModelDesc
from ppl_llm_serving_b200 import capi
from ppl_llm_serving_b200.engine import CudaResourceManager, LLMEngine, ModelInput, ModelOutput, RC_SUCCESS

B, KV, PAGE = 1024, 512, 16
desc = ModelDesc(4096, 11008, 1, 32, 32, 32000, cache_layout=3, cache_mode=1, page_size=PAGE, quant_method=1, max_position=1024)
T = B * KV
res = CudaResourceManager()
assert res.Init(desc, 0.9, B, B, kv_cache_max_tokens=T, seed=0xB200) == RC_SUCCESS
engine = LLMEngine(res, False, 1, 0.0)
g = torch.Generator(device="cuda").manual_seed(1)
H, D = 32, 128
logical = torch.randint(-127, 128, (2 * H, T, D), dtype=torch.int8, device="cuda", generator=g)
logical_s = (torch.rand((2 * H, T, D // 8), device="cuda", generator=g) * 0.02 + 0.001).to(torch.float16)
tokens = np.random.default_rng(2).integers(0, desc.vocab_size, B).astype(np.int64)
dup = "--dup" in sys.argv
if dup:
    tokens[1] = tokens[0]
    logical[:, KV:2 * KV] = logical[:, 0:KV]
    logical_s[:, KV:2 * KV] = logical_s[:, 0:KV]
chk0 = (logical.to(torch.int32).sum().item(), logical_s.float().sum().item())
outs = {}
for seed in (10, 11):
    perm = np.random.default_rng(seed).permutation(T // PAGE)
    page_list = (perm.reshape(B, KV // PAGE) * PAGE).astype(np.int64)
    slot = torch.from_numpy((page_list[:, :, None] + np.arange(PAGE)[None, None, :]).reshape(-1)).cuda()
    cache = res.kv_cache_mem.view(2 * H, T, D)
    scale = res.kv_scale_mem.view(2 * H, T, D // 8)
    cache.index_copy_(1, slot, logical)
    scale.index_copy_(1, slot, logical_s)
    torch.cuda.synchronize()
    mi = ModelInput(token_inputs=tokens, seq_starts=np.arange(B + 1, dtype=np.int64), kv_starts=np.arange(B + 1, dtype=np.int64) * KV,
                    start_pos=np.full(B, KV - 1, dtype=np.int64), page_list=page_list.reshape(-1), max_pages=KV // PAGE,
                    decoding_batches=B, max_seq_len=1, max_kv_len=KV, temperatures=[1.0] * B, top_p_list=[0.0] * B, top_k_list=[1] * B)
    for mode in (1, 0, 2, 1):
        capi.check(res.lib.b2llm_engine_configure(res.engine, 3, mode), "cfg")
        out = ModelOutput(); out.Resize(B)
        rc, err = engine.Execute(mi, True, False, out)
        assert rc == RC_SUCCESS, err
        outs[(seed, mode, len([k for k in outs if k[0] == seed and k[1] == mode]))] = (engine.logits(B), engine.debug_read(2, (B, H * D), np.float16).astype(np.float32),
                                                                                      engine.debug_read(1, (B, 3 * H * D), np.float16).astype(np.float32))
    print("logical checksum unchanged:", chk0 == (logical.to(torch.int32).sum().item(), logical_s.float().sum().item()))
keys = list(outs)
base = outs[keys[0]]
for k in keys[1:]:
    o = outs[k]
    print(k, "vs", keys[0], "dlogits %.3e dattn %.3e dqkv %.3e" % (np.abs(o[0] - base[0]).max(), np.abs(o[1] - base[1]).max(), np.abs(o[2] - base[2]).max()),
          "rows differing in attn:", int((np.abs(o[1] - base[1]).max(axis=1) > 0).sum()), flush=True)
if dup:
    print("rows 0/1 equal:", np.array_equal(base[0][0], base[0][1]))
