#!/bin/bash
# run 35 (TEMPORARY instrumentation, not committed): epilogue cycles of the transposed W4 kernel, direct vs split-K partial path
export FUSED_ONLY=1
for sk in 1 2; do
 for sh in o qkv gate_up; do
  echo "## SPLITK=$sk $sh"; SHAPES=$sh B2LLM_W4_SPLITK=$sk timeout 120 python scripts/gemm_w4_bench.py 256 2>&1 | tail -3
 done
done
