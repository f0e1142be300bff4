#!/bin/bash
# NOTE: the W4TPROF lines this run printed came from clock64 counters added to gemm_w4t_kernel for this run only (never
# committed); the findings are in profiles/r2_gemm_w4_transposed.txt section 6.  Without them the script just times the shapes.
# run 35 (TEMPORARY instrumentation, not committed): epilogue cycles of the transposed W4 kernel, direct vs split-K partial path
export FUSED_ONLY=1
for sk in 1 2; do
 for sh in o qkv gate_up; do
  echo "## SPLITK=$sk $sh"; SHAPES=$sh B2LLM_W4_SPLITK=$sk timeout 120 python scripts/gemm_w4_bench.py 256 2>&1 | tail -3
 done
done
