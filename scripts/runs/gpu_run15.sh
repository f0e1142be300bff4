#!/bin/bash
# round 1, run 15 (2 x B200): the tensor-parallel path of the ONNX model-slice loader -- two pmx-style slices, embedding /
# lm head re-assembled from both, b2llm_engine_load_weight_shard per rank -- through the reference's own TP bring-up.
set -x
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_host_cpp.py -x -q -m gpu -k "pmx_onnx_export_tensor_parallel_2 or generator_tensor_parallel_2" > gpurun_out/pytest15_tp2.log 2>&1; echo "rc=$?" | tee -a gpurun_out/pytest15_tp2.log; tail -8 gpurun_out/pytest15_tp2.log | cut -c1-300
