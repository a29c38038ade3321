#!/bin/bash
# Runs each kernel family's GPU parity tests in its own process (a trapping kernel cannot poison the others).
mkdir -p gpurun_out/k1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/k1/smi.txt 2>&1
: > gpurun_out/k1/summary.txt
for k in gemm_kmajor gemm_epilogues gemm_dgrad gemm_wgrad gemm_batched gemm_seq gemm_strided gemm_rejects layernorm rope attention latent cnn token_helpers spectral; do
  timeout 240 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k $k --tb=short > gpurun_out/k1/$k.log 2>&1
  echo "$k rc=$? $(tail -1 gpurun_out/k1/$k.log)" >> gpurun_out/k1/summary.txt
done
cat gpurun_out/k1/summary.txt
