"""Writer of a model directory in the layout of a ppl.pmx LLaMA export (docs/llama_guide.md:12-36 of the reference):

    <dir>/params.json                       keys of src/common/config.cc:41-145
    <dir>/model_slice_<rank>/model.onnx     ONNX ModelProto (schema: src/onnx/onnx.proto): pmx-domain nodes carrying
                                            the graph constants as attributes + this rank's weight shards as
                                            initializers (inline raw_data, or ONNX external data files)

It exists so that the C++ model-slice loader (host/src/{onnx_model,pmx_llama}.cc, behind
``ppl::nn::onnx::RuntimeBuilder::LoadModel``) can be exercised without a real export: no ppl.pmx checkout and no
checkpoint exist offline.  The bytes are produced by the official ``google.protobuf`` runtime from a descriptor
built here with the field numbers of the reference's ``onnx.proto`` -- an encoder independent of the hand-written
wire reader it tests.  Tensor names, shard shapes and node attributes follow ppl.pmx's LLaMA model
(``tok_embeddings``, ``layers.<i>.attention.wqkv`` ..., ``RMSNorm.eps``, ``RotaryPositionEmbedding.theta``,
``MultiHeadCacheAttention.{num_heads, head_dim, cache_layout ...}``); see host/src/pmx_llama.cc for the list.
"""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

# TensorProto.DataType / AttributeProto.AttributeType (onnx.proto:480-507, 118-134)
DT_FLOAT, DT_FLOAT16, DT_BFLOAT16 = 1, 10, 16
AT_FLOAT, AT_INT, AT_STRING = 1, 2, 3

_MESSAGES: dict = {}


def onnx_messages(syntax: str = "proto3") -> dict:
    """message classes of the ONNX subset; ``syntax`` "proto3" packs repeated scalars (dims, ints), "proto2" writes one
    tag per element -- real files contain both encodings and a reader has to take either"""
    if syntax in _MESSAGES:
        return _MESSAGES[syntax]
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

    F = descriptor_pb2.FieldDescriptorProto
    pkg = "b2onnx_" + syntax
    fdp = descriptor_pb2.FileDescriptorProto(name=f"{pkg}.proto", package=pkg, syntax=syntax)

    def msg(name, *fields):
        m = fdp.message_type.add(name=name)
        for fname, number, ftype, label, type_name in fields:
            f = m.field.add(name=fname, number=number, type=ftype, label=label)
            if type_name:
                f.type_name = f".{pkg}.{type_name}"

    OPT, REP = F.LABEL_OPTIONAL, F.LABEL_REPEATED
    msg("StringStringEntryProto", ("key", 1, F.TYPE_STRING, OPT, None), ("value", 2, F.TYPE_STRING, OPT, None))
    msg("OperatorSetIdProto", ("domain", 1, F.TYPE_STRING, OPT, None), ("version", 2, F.TYPE_INT64, OPT, None))
    msg("TensorProto", ("dims", 1, F.TYPE_INT64, REP, None), ("data_type", 2, F.TYPE_INT32, OPT, None),
        ("float_data", 4, F.TYPE_FLOAT, REP, None), ("int32_data", 5, F.TYPE_INT32, REP, None),
        ("name", 8, F.TYPE_STRING, OPT, None), ("raw_data", 9, F.TYPE_BYTES, OPT, None),
        ("doc_string", 12, F.TYPE_STRING, OPT, None),
        ("external_data", 13, F.TYPE_MESSAGE, REP, "StringStringEntryProto"), ("data_location", 14, F.TYPE_INT32, OPT, None))
    msg("AttributeProto", ("name", 1, F.TYPE_STRING, OPT, None), ("f", 2, F.TYPE_FLOAT, OPT, None),
        ("i", 3, F.TYPE_INT64, OPT, None), ("s", 4, F.TYPE_BYTES, OPT, None), ("floats", 7, F.TYPE_FLOAT, REP, None),
        ("ints", 8, F.TYPE_INT64, REP, None), ("doc_string", 13, F.TYPE_STRING, OPT, None), ("type", 20, F.TYPE_INT32, OPT, None))
    msg("NodeProto", ("input", 1, F.TYPE_STRING, REP, None), ("output", 2, F.TYPE_STRING, REP, None),
        ("name", 3, F.TYPE_STRING, OPT, None), ("op_type", 4, F.TYPE_STRING, OPT, None),
        ("attribute", 5, F.TYPE_MESSAGE, REP, "AttributeProto"), ("doc_string", 6, F.TYPE_STRING, OPT, None),
        ("domain", 7, F.TYPE_STRING, OPT, None))
    msg("ValueInfoProto", ("name", 1, F.TYPE_STRING, OPT, None), ("doc_string", 3, F.TYPE_STRING, OPT, None))
    msg("GraphProto", ("node", 1, F.TYPE_MESSAGE, REP, "NodeProto"), ("name", 2, F.TYPE_STRING, OPT, None),
        ("initializer", 5, F.TYPE_MESSAGE, REP, "TensorProto"), ("doc_string", 10, F.TYPE_STRING, OPT, None),
        ("input", 11, F.TYPE_MESSAGE, REP, "ValueInfoProto"), ("output", 12, F.TYPE_MESSAGE, REP, "ValueInfoProto"))
    msg("ModelProto", ("ir_version", 1, F.TYPE_INT64, OPT, None), ("producer_name", 2, F.TYPE_STRING, OPT, None),
        ("producer_version", 3, F.TYPE_STRING, OPT, None), ("domain", 4, F.TYPE_STRING, OPT, None),
        ("model_version", 5, F.TYPE_INT64, OPT, None), ("doc_string", 6, F.TYPE_STRING, OPT, None),
        ("graph", 7, F.TYPE_MESSAGE, OPT, "GraphProto"), ("opset_import", 8, F.TYPE_MESSAGE, REP, "OperatorSetIdProto"))
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fdp)
    names = [m.name for m in fdp.message_type]
    classes = {n: message_factory.GetMessageClass(pool.FindMessageTypeByName(f"{pkg}.{n}")) for n in names}
    _MESSAGES[syntax] = classes
    return classes


def _to_bf16_bits(a: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)  # round to nearest even


def shard_weights(cfg, weights, rank: int, tp: int, fused_qkv: bool = True, emb_split: str = "hidden",
                  head_split: str = "vocab") -> dict:
    """rank ``rank``'s tensors of a ``tp``-way export, by ppl.pmx parameter name -> np.float16 array"""
    h, D = cfg.hidden_dim, cfg.hidden_dim // cfg.num_heads
    NQ, NKV, I, V = cfg.num_heads, cfg.num_kv_heads, cfg.intermediate_dim, cfg.vocab_size
    nq, nkv, il = NQ // tp, NKV // tp, I // tp
    out = {}
    emb = np.asarray(weights.embedding(), dtype=np.float16)
    head = np.asarray(weights.lm_head(), dtype=np.float16)
    if tp > 1 and emb_split == "hidden":
        emb = emb[:, rank * (h // tp):(rank + 1) * (h // tp)]
    elif tp > 1 and emb_split == "vocab":
        emb = emb[rank * (V // tp):(rank + 1) * (V // tp)]
    if tp > 1 and head_split == "vocab":
        head = head[rank * (V // tp):(rank + 1) * (V // tp)]
    out["tok_embeddings.weight"] = emb
    for l in range(cfg.num_layers):
        w = weights.layer(l)
        p = f"layers.{l}."
        wqkv = np.asarray(w["wqkv"], dtype=np.float16)
        q = wqkv[rank * nq * D:(rank + 1) * nq * D]
        k = wqkv[NQ * D + rank * nkv * D: NQ * D + (rank + 1) * nkv * D]
        v = wqkv[(NQ + NKV) * D + rank * nkv * D:(NQ + NKV) * D + (rank + 1) * nkv * D]
        out[p + "attention_norm.weight"] = np.asarray(w["attn_norm"], dtype=np.float16)
        if fused_qkv:
            out[p + "attention.wqkv.weight"] = np.concatenate([q, k, v], axis=0)
        else:
            out[p + "attention.wq.weight"], out[p + "attention.wk.weight"], out[p + "attention.wv.weight"] = q, k, v
        out[p + "attention.wo.weight"] = np.asarray(w["wo"], dtype=np.float16)[:, rank * nq * D:(rank + 1) * nq * D]
        out[p + "ffn_norm.weight"] = np.asarray(w["ffn_norm"], dtype=np.float16)
        out[p + "feed_forward.w1.weight"] = np.asarray(w["wgate"], dtype=np.float16)[rank * il:(rank + 1) * il]
        out[p + "feed_forward.w3.weight"] = np.asarray(w["wup"], dtype=np.float16)[rank * il:(rank + 1) * il]
        out[p + "feed_forward.w2.weight"] = np.asarray(w["wdown"], dtype=np.float16)[:, rank * il:(rank + 1) * il]
    out["norm.weight"] = np.asarray(weights.final_norm(), dtype=np.float16)
    out["output.weight"] = head
    return {k: np.ascontiguousarray(v) for k, v in out.items()}


def write_pmx_export(path, cfg, weights, tensor_parallel_size: int = 1, fused_qkv: bool = True,
                     external_data: bool = False, dtype: str = "fp16", syntax: str = "proto3",
                     emb_split: str = "hidden", head_split: str = "vocab", typed_data: bool = False,
                     write_params: bool = True, attrs: bool = True, domain: str = "pmx") -> Path:
    """``weights``: object with embedding()/final_norm()/lm_head()/layer(l) (e.g. the oracle's SynthWeights).
    ``dtype``: "fp16" | "fp32" | "bf16" payload type; ``external_data``: one file per initializer beside model.onnx, as
    torch.onnx.export does for models over 2 GB; ``typed_data``: payload in int32_data / float_data instead of raw_data;
    ``attrs`` False drops the pmx attributes (the loader then falls back to params.json)."""
    P = onnx_messages(syntax)
    path = Path(path)
    path.mkdir(parents=True, exist_ok=True)
    tp = tensor_parallel_size
    h, D = cfg.hidden_dim, cfg.hidden_dim // cfg.num_heads
    if write_params:
        params = {
            "num_heads": cfg.num_heads, "num_kv_heads": cfg.num_kv_heads, "num_layers": cfg.num_layers,
            "hidden_dim": cfg.hidden_dim, "intermediate_dim": cfg.intermediate_dim, "vocab_size": cfg.vocab_size,
            "cache_quant_bit": cfg.cache_quant_bit, "cache_quant_group": cfg.cache_quant_group,
            "cache_layout": cfg.cache_layout, "cache_mode": cfg.cache_mode, "page_size": cfg.page_size,
            "dynamic_batching": True, "auto_causal": True,
        }
        (path / "params.json").write_text(json.dumps(params, indent=1) + "\n")

    def attr(name, value):
        a = P["AttributeProto"](name=name)
        if isinstance(value, float):
            a.f, a.type = value, AT_FLOAT
        elif isinstance(value, str):
            a.s, a.type = value.encode(), AT_STRING
        else:
            a.i, a.type = int(value), AT_INT
        return a

    for r in range(tp):
        sd = path / f"model_slice_{r}"
        sd.mkdir(exist_ok=True)
        tensors = shard_weights(cfg, weights, r, tp, fused_qkv, emb_split, head_split)
        g = P["GraphProto"](name="torch_jit")
        for nm in ("token_ids", "attn_mask", "seqstarts", "kvstarts", "cachestarts", "decoding_batches", "start_pos",
                   "max_seqlen", "max_kvlen", "kv_cache", "kv_scale"):  # llm_engine.h:124-138
            g.input.add(name=nm)
        g.output.add(name="logits")

        def node(op, inputs, outputs, dom=domain, **kw):
            n = g.node.add(op_type=op, domain=dom, name=f"{op}_{len(g.node)}")
            n.input.extend(inputs)
            n.output.extend(outputs)
            if attrs:
                n.attribute.extend(attr(k, v) for k, v in kw.items())
            return n

        cache_kw = dict(num_layer=cfg.num_layers, quant_bit=cfg.cache_quant_bit, quant_group=cfg.cache_quant_group,
                        cache_mode=cfg.cache_mode, cache_layout=cfg.cache_layout, page_size=cfg.page_size)
        dyn = domain + ".dynamic_batching"
        node("ParallelEmbedding", ["token_ids", "tok_embeddings.weight"], ["x0"], num_embeddings=cfg.vocab_size,
             embedding_dims=h, padding_idx=-1, max_norm=0.0, norm_type=2.0)
        x = "x0"
        for l in range(cfg.num_layers):
            p = f"layers.{l}."
            node("RMSNorm", [x, p + "attention_norm.weight"], [f"n{l}a"], axis=-1, eps=float(cfg.norm_eps), skip_term=0)
            if fused_qkv:
                node("ColumnParallelLinear", [f"n{l}a", p + "attention.wqkv.weight"], [f"qkv{l}"], in_features=h,
                     out_features=(cfg.num_heads + 2 * cfg.num_kv_heads) * D, bias_term=0, gather_output=0)
                node("Reshape", [f"qkv{l}", "shape_heads"], [f"qkv{l}r"], dom="")
                node("Split", [f"qkv{l}r"], [f"q{l}", f"k{l}", f"v{l}"], dom="", axis=1)
            else:
                for nm, t, heads in (("q", "wq", cfg.num_heads), ("k", "wk", cfg.num_kv_heads), ("v", "wv", cfg.num_kv_heads)):
                    node("ColumnParallelLinear", [f"n{l}a", p + f"attention.{t}.weight"], [f"{nm}{l}"], in_features=h,
                         out_features=heads * D, bias_term=0, gather_output=0)
            node("RotaryPositionEmbedding", [f"q{l}", f"k{l}", "seqstarts", "start_pos", "max_seqlen"], [f"rq{l}", f"rk{l}"],
                 dom=dyn, bypass_key=0, rotary_dim=0, theta=float(cfg.rope_theta), max_position_embeddings=int(cfg.max_position),
                 scaling_type="", scaling_factor=1.0)
            node("MultiHeadCacheAttention",
                 [f"rq{l}", f"rk{l}", f"v{l}", "seqstarts", "kvstarts", "cachestarts", "start_pos", "decoding_batches", "max_seqlen",
                  "max_kvlen", "kv_cache", "kv_scale"], [f"a{l}"], dom=dyn, num_heads=cfg.num_heads // tp, head_dim=D,
                 is_causal=1, is_alibi=0, num_kv_heads=cfg.num_kv_heads // tp, layer_idx=l, **cache_kw)
            node("RowParallelLinear", [f"a{l}", p + "attention.wo.weight"], [f"o{l}"], in_features=h, out_features=h, bias_term=0,
                 input_is_parallel=1)
            node("Add", [x, f"o{l}"], [f"x{l}m"], dom="")
            node("RMSNorm", [f"x{l}m", p + "ffn_norm.weight"], [f"n{l}f"], axis=-1, eps=float(cfg.norm_eps), skip_term=0)
            node("ColumnParallelLinear", [f"n{l}f", p + "feed_forward.w1.weight"], [f"g{l}"], in_features=h,
                 out_features=cfg.intermediate_dim, bias_term=0, gather_output=0)
            node("ColumnParallelLinear", [f"n{l}f", p + "feed_forward.w3.weight"], [f"u{l}"], in_features=h,
                 out_features=cfg.intermediate_dim, bias_term=0, gather_output=0)
            node("SiLU", [f"g{l}", f"u{l}"], [f"s{l}"], gated=1)
            node("RowParallelLinear", [f"s{l}", p + "feed_forward.w2.weight"], [f"d{l}"], in_features=cfg.intermediate_dim,
                 out_features=h, bias_term=0, input_is_parallel=1)
            node("Add", [f"x{l}m", f"d{l}"], [f"x{l + 1}"], dom="")
            x = f"x{l + 1}"
        node("RMSNorm", [x, "norm.weight"], ["xn"], axis=-1, eps=float(cfg.norm_eps), skip_term=0)
        node("ColumnParallelLinear", ["xn", "output.weight"], ["logits"], in_features=h, out_features=cfg.vocab_size,
             bias_term=0, gather_output=1)

        for name, a in tensors.items():
            t = g.initializer.add(name=name)
            t.dims.extend(a.shape)
            if dtype == "fp32":
                t.data_type, payload = DT_FLOAT, a.astype(np.float32)
            elif dtype == "bf16":
                t.data_type, payload = DT_BFLOAT16, _to_bf16_bits(a.astype(np.float32))
            else:
                t.data_type, payload = DT_FLOAT16, a
            if external_data:
                (sd / name).write_bytes(payload.tobytes())
                t.data_location = 1
                t.external_data.add(key="location", value=name)
                t.external_data.add(key="offset", value="0")
                t.external_data.add(key="length", value=str(payload.nbytes))
            elif typed_data and dtype == "fp32":
                t.float_data.extend(payload.reshape(-1).tolist())
            elif typed_data:
                t.int32_data.extend(payload.reshape(-1).view(np.uint16).astype(np.int32).tolist())
            else:
                t.raw_data = payload.tobytes()
        m = P["ModelProto"](ir_version=8, producer_name="pytorch", producer_version="2.0.0", graph=g)
        m.opset_import.add(domain="", version=11)
        m.opset_import.add(domain=domain, version=1)
        m.opset_import.add(domain=dyn, version=1)
        (sd / "model.onnx").write_bytes(m.SerializeToString())
    return path
