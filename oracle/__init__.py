"""CPU oracle for the batched LLaMA decode hot path of OpenPPL/ppl.llm.serving.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker (or as
the CPU arm being timed), never as the thing measured or shipped.

PARITY UNPINNED.  The arithmetic of this path does not live in
``/root/reference``: it lives in OpenPPL/ppl.nn @ master (floating, fetched by
``cmake/deps.cmake:92-106``) and its sub-dependency ppl.llm.kernel.cuda, neither
of which is vendored, and the reference holds no golden vectors for the path
(``test/test_prefix_cache_mgr.cc:25-66`` asserts nothing).  This oracle therefore
restates (a) the tensor / IO contract that *is* in the reference
(``src/engine/llm_engine.h:40-73``, ``src/engine/llm_engine.cc:29-169``,
``src/generator/llm_generator.cc:263-298``) and (b) the published LLaMA-2 /
ppl.pmx operator definitions (RMSNorm, rotary embedding, online int8 per-token /
per-channel quantisation, int8 group-8 KV cache, SwiGLU, top-k/top-p sampling).
Host-side integer logic that *is* pinnable from reference source
(``HashCombine``, page counting, kv_starts construction, the KV budget formula,
PrefixCacheManager's refcount/LRU sequence) is pinned as known-answer tests in
``tests/test_host_kat.py``.
"""
