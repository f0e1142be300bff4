"""Writer of the model directory the reference's tools point at (``--model-dir``, ``--model-param-path``):

    <dir>/params.json                       keys of src/common/config.cc:41-145
    <dir>/model_slice_<rank>/model.onnx     resource_manager.cc:280-286 -- here a b2llm model-slice descriptor
    <dir>/weights.fp16                      optional fp16 blob (full tensors; every rank slices its part)

The reference addresses the model only by that path and by tensor index (llm_engine.h:124-138); what is inside
``model.onnx`` is private to the runtime behind ``ppl::nn::onnx::RuntimeBuilder``, which reads two kinds of file here:
a real ppl.pmx ONNX export (host/src/{onnx_model,pmx_llama}.cc; written for tests by ``pmx_onnx_writer.py``) and this
text descriptor for synthetic (seeded) weights or an fp16 blob -- what bench.py, the steady-state probe and most tests
use, because it needs no multi-GB files.  See INTEGRATION.md section 4.
"""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

BLOB_ORDER = ("embedding, final_norm, lm_head, then per layer: attn_norm, wqkv [(nq+2nkv)*D, h] (q heads, k heads, "
              "v heads), wo [h, nq*D], ffn_norm, wgate [I, h], wup [I, h], wdown [h, I]; all fp16, row-major, unsharded")


def write_model_dir(path, cfg, tensor_parallel_size: int = 1, seed: int | None = 0xB200, weights=None) -> Path:
    """``cfg``: object with ModelConfig attributes (engine.ModelConfig or oracle.weights.ModelDesc).
    ``weights``: optional object with embedding()/final_norm()/lm_head()/layer(l) (e.g. oracle SynthWeights) to be
    written as an fp16 blob; otherwise the runtime generates the seeded synthetic weights on the device."""
    path = Path(path)
    path.mkdir(parents=True, exist_ok=True)
    params = {
        "num_heads": cfg.num_heads, "num_kv_heads": cfg.num_kv_heads, "num_layers": cfg.num_layers,
        "hidden_dim": cfg.hidden_dim, "intermediate_dim": cfg.intermediate_dim, "vocab_size": cfg.vocab_size,
        "cache_quant_bit": cfg.cache_quant_bit, "cache_quant_group": cfg.cache_quant_group,
        "cache_layout": cfg.cache_layout, "cache_mode": cfg.cache_mode, "page_size": cfg.page_size,
        "dynamic_batching": True, "auto_causal": True,
    }
    (path / "params.json").write_text(json.dumps(params, indent=1) + "\n")
    source = f"synthetic:{int(seed)}"
    if weights is not None:
        with open(path / "weights.fp16", "wb") as f:
            def put(a):
                f.write(np.ascontiguousarray(a, dtype=np.float16).tobytes())
            put(weights.embedding()); put(weights.final_norm()); put(weights.lm_head())
            for l in range(cfg.num_layers):
                w = weights.layer(l)
                for k in ("attn_norm", "wqkv", "wo", "ffn_norm", "wgate", "wup", "wdown"):
                    put(w[k])
        source = "file:../weights.fp16"
    for r in range(tensor_parallel_size):
        d = path / f"model_slice_{r}"
        d.mkdir(exist_ok=True)
        lines = ["b2llm-model-slice 1", f"# {BLOB_ORDER}"]
        for k in ("hidden_dim", "intermediate_dim", "num_layers", "num_heads", "num_kv_heads", "vocab_size",
                  "cache_quant_bit", "cache_quant_group", "cache_layout", "cache_mode", "page_size", "max_position"):
            lines.append(f"{k} = {int(getattr(cfg, k))}")
        lines.append(f"norm_eps = {float(cfg.norm_eps)!r}")
        lines.append(f"rope_theta = {float(cfg.rope_theta)!r}")
        lines.append(f"tensor_parallel_size = {tensor_parallel_size}")
        lines.append(f"rank = {r}")
        lines.append(f"weights = {source}")
        (d / "model.onnx").write_text("\n".join(lines) + "\n")
    return path
