"""numpy restatement of the sampler / penalty operators behind ``CudaPostProcessor``.

Test infrastructure (see ``oracle/__init__.py``); PARITY UNPINNED: the kernels
``pmx::sample_topk_topp`` / ``pmx::apply_penalty`` are EXTERNAL (ppl.llm.kernel.cuda); the
reference holds only their call sites (``src/backends/cuda/post_processor.cc:190-193``,
``:271-274``).  Argument meaning follows those call sites:

  sample_topk_topp(logits[B, stride] fp32, temperatures|None, top_p|None, rand[B], batch, vocab,
                   stride, top_k, default_top_p, default_rand, out int32[B], logprobs f32[B])
    * ``temperatures`` / ``top_p`` are passed only on steps whose request list changed, otherwise
      null -> temperature 1 and ``default_top_p`` (post_processor.cc:154-177, a reference quirk).
    * ``top_k`` is one value for the whole batch, taken from request 0 (llm_engine.cc:219).
  Semantics fixed here:
    l' = l / T;  candidates = the top_k largest l' (ties: lower index first), descending;
    p_i = softmax over the candidates;  keep the shortest prefix whose cumulative p reaches top_p
    (always >= 1 candidate; top_p <= 0 keeps exactly one -> greedy);  draw r = rand * (kept mass),
    pick the first candidate with cumulative p > r (last kept one if none);
    logprob = log_softmax(l')[token] over the FULL vocabulary.
  top_k == 1 is greedy arg-max (lowest index on ties).

  apply_penalty(logits, temperatures, repetition, presence|None, frequency|None, batch_slots,
                token_inputs, seqstarts, start_pos, batch, vocab, count_map uint16[slots, vocab])
    * a sequence whose start_pos is 0 (first step) first clears its slot's count row;
    * this step's input tokens are counted into the row (saturating at 65535);
    * for every vocab entry with count > 0: l = l / rep if l > 0 else l * rep;
      l -= presence; l -= frequency * count;   then every entry: l /= temperature.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def sample_topk_topp(logits, temperatures, top_p, rand, vocab, top_k, default_top_p, default_rand=0.0):
    logits = np.asarray(logits, dtype=F32)
    B = logits.shape[0]
    out = np.zeros(B, dtype=np.int32)
    logprobs = np.zeros(B, dtype=F32)
    for b in range(B):
        T = F32(1.0) if temperatures is None else F32(temperatures[b])
        tp = F32(default_top_p) if top_p is None else F32(top_p[b])
        r = F32(default_rand) if rand is None else F32(rand[b])
        l = (logits[b, :vocab] / T).astype(F32)
        m = l.max()
        lse = F32(np.log(np.exp((l - m).astype(np.float64)).sum())) + m
        k = max(1, min(int(top_k), vocab))
        # stable: descending value, ascending index on ties
        order = np.lexsort((np.arange(vocab), -l.astype(np.float64)))[:k]
        cand = l[order]
        e = np.exp((cand - cand[0]).astype(F32)).astype(F32)
        p = (e / np.cumsum(e, dtype=F32)[-1]).astype(F32)  # sequential fp32 sum, as the kernel
        cum = np.cumsum(p, dtype=F32)
        keep = k
        if tp <= 0:
            keep = 1
        else:
            reach = np.nonzero(cum >= tp)[0]
            keep = int(reach[0]) + 1 if len(reach) else k
        mass = cum[keep - 1]
        thr = F32(r * mass)
        sel = np.nonzero(cum[:keep] > thr)[0]
        i = int(sel[0]) if len(sel) else keep - 1
        out[b] = order[i]
        logprobs[b] = l[order[i]] - lse
    return out, logprobs


def apply_penalty(logits, temperatures, repetition, presence, frequency, batch_slots, token_inputs,
                  seqstarts, start_pos, vocab, count_map, next_pos=None):
    """in-place on ``logits`` [B, >=vocab] and ``count_map`` uint16 [slots, vocab].

    A slot's counts are cleared on the first step of a request: ``start_pos == 0``, or -- with ``next_pos`` (a dict
    slot -> expected next position, kept by the caller across steps) -- whenever the step does not continue the slot's
    previous one.  The second rule is what catches a request admitted on a prefix-cache hit, which enters at
    ``start_pos = cache_hit_count`` (llm_generator.cc:229-242) into a slot another request used before."""
    B = len(start_pos)
    for b in range(B):
        slot = int(batch_slots[b])
        reset = int(start_pos[b]) == 0
        if next_pos is not None:
            reset = reset or next_pos.get(slot, -1) != int(start_pos[b])
            next_pos[slot] = int(start_pos[b]) + int(seqstarts[b + 1]) - int(seqstarts[b])
        if reset:
            count_map[slot, :] = 0
        for t in token_inputs[int(seqstarts[b]): int(seqstarts[b + 1])]:
            if count_map[slot, int(t)] < 65535:
                count_map[slot, int(t)] += 1
        c = count_map[slot, :vocab].astype(F32)
        l = logits[b, :vocab].astype(F32)
        seen = c > 0
        rep = F32(repetition[b])
        l = np.where(seen, np.where(l > 0, l / rep, l * rep), l).astype(F32)
        if presence is not None:
            l = np.where(seen, l - F32(presence[b]), l).astype(F32)
        if frequency is not None:
            l = (l - F32(frequency[b]) * c).astype(F32)
        logits[b, :vocab] = (l / F32(temperatures[b])).astype(F32)
    return logits
