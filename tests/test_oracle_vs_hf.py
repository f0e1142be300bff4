"""Independent cross-check of the oracle's MODEL DEFINITION (CPU): the fp16-weight mode of the oracle against
Hugging Face transformers' LlamaForCausalLM in fp32 carrying the same synthetic weights.

The reference does not contain the model math (Runtime::Run() is external, llm_engine.cc:113-116) and has no
golden logits, so "LLaMA-2" is anchored on the public definition: pre-norm RMSNorm, rotate-half RoPE
(theta 10000), SwiGLU, GQA.  transformers is library code in this image; it is NOT the reference, only a
second implementation of the same public definition.  Differences are the oracle's fp16 activation
roundings and its int8 KV cache (decode step), so the bar is loose (2 % of the logit range) but catches
any structural mistake (head ordering, RoPE pairing, norm placement, gate/up swap) which move logits by
O(100 %).
"""
import numpy as np
import pytest

from oracle import llama_ref as ref
from oracle.weights import ModelDesc, SynthWeights

torch = pytest.importorskip("torch")
transformers = pytest.importorskip("transformers")


def _hf_model(desc, w):
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(vocab_size=desc.vocab_size, hidden_size=desc.hidden_dim, intermediate_size=desc.intermediate_dim,
                      num_hidden_layers=desc.num_layers, num_attention_heads=desc.num_heads,
                      num_key_value_heads=desc.num_kv_heads, rms_norm_eps=desc.norm_eps, rope_theta=desc.rope_theta,
                      max_position_embeddings=desc.max_position, tie_word_embeddings=False, attention_bias=False,
                      mlp_bias=False, hidden_act="silu")
    cfg._attn_implementation = "eager"
    m = LlamaForCausalLM(cfg).to(torch.float32).eval()
    t = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))
    D, Hq, Hkv = desc.head_dim, desc.num_heads, desc.num_kv_heads
    sd = {"model.embed_tokens.weight": t(w.embedding()), "model.norm.weight": t(w.final_norm()),
          "lm_head.weight": t(w.lm_head())}
    for l in range(desc.num_layers):
        lw = w.layer(l)
        p = f"model.layers.{l}."
        sd[p + "input_layernorm.weight"] = t(lw["attn_norm"])
        sd[p + "post_attention_layernorm.weight"] = t(lw["ffn_norm"])
        sd[p + "self_attn.q_proj.weight"] = t(lw["wqkv"][: Hq * D])
        sd[p + "self_attn.k_proj.weight"] = t(lw["wqkv"][Hq * D: (Hq + Hkv) * D])
        sd[p + "self_attn.v_proj.weight"] = t(lw["wqkv"][(Hq + Hkv) * D:])
        sd[p + "self_attn.o_proj.weight"] = t(lw["wo"])
        sd[p + "mlp.gate_proj.weight"] = t(lw["wgate"])
        sd[p + "mlp.up_proj.weight"] = t(lw["wup"])
        sd[p + "mlp.down_proj.weight"] = t(lw["wdown"])
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("rotary" in k or "inv_freq" in k for k in missing), (missing, unexpected)
    return m


@pytest.mark.parametrize("kvh", [4, 2])
def test_oracle_fp16_mode_matches_hf_llama(kvh):
    desc = ModelDesc(256, 512, 2, 4, kvh, 320, cache_layout=2, cache_mode=0, quant_method=0, max_position=64)
    w = SynthWeights(desc, 0xB200)
    hf = _hf_model(desc, w)
    rng = np.random.default_rng(3)
    prompts = [list(map(int, rng.integers(0, desc.vocab_size, n))) for n in (11, 4)]
    orc = ref.LlamaOracle(desc, w, 64)
    logits = orc.forward(ref.build_step(desc, prompts, [0, 0], 0, cache_indices=[0, 32]))
    with torch.no_grad():
        want = np.stack([hf(torch.tensor([p])).logits[0, -1].numpy() for p in prompts])
    scale = np.abs(want).max()
    assert np.abs(logits - want).max() <= 2e-2 * scale, np.abs(logits - want).max() / scale
    assert logits.argmax(axis=1).tolist() == want.argmax(axis=1).tolist()
    # one decode step (reads the int8 group-8 cache for the prefix): still the same function
    nxt = [int(t) for t in want.argmax(axis=1)]
    l2 = orc.forward(ref.build_step(desc, [[t] for t in nxt], [11, 4], 2, cache_indices=[0, 32]))
    with torch.no_grad():
        want2 = np.stack([hf(torch.tensor([p + [t]])).logits[0, -1].numpy() for p, t in zip(prompts, nxt)])
    assert np.abs(l2 - want2).max() <= 3e-2 * np.abs(want2).max()


def test_oracle_w8a8_mode_close_to_hf_llama():
    """W8A8 adds quantisation noise of ~1/127 per operand; logits stay within 10 % of range of the fp32 model
    and the int8 model's arg-max stays among the fp32 model's top 5."""
    desc = ModelDesc(256, 512, 2, 4, 4, 320, cache_layout=3, cache_mode=1, page_size=16, quant_method=1, max_position=64)
    w = SynthWeights(desc, 0xB200)
    hf = _hf_model(desc, w)
    rng = np.random.default_rng(4)
    p = list(map(int, rng.integers(0, desc.vocab_size, 13)))
    orc = ref.LlamaOracle(desc, w, 64)
    logits = orc.forward(ref.build_step(desc, [p], [0], 0, page_tables=[[16]]))
    with torch.no_grad():
        want = hf(torch.tensor([p])).logits[0, -1].numpy()
    assert np.abs(logits[0] - want).max() <= 0.1 * np.abs(want).max()
    assert int(logits[0].argmax()) in np.argsort(-want)[:5].tolist()
