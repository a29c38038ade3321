#!/bin/bash
mkdir -p gpurun_out/r8
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k attention --tb=short > gpurun_out/r8/k_attention.log 2>&1
echo "kernels:attention rc=$? $(tail -1 gpurun_out/r8/k_attention.log)"
grep -E "^FAILED|^E  " gpurun_out/r8/k_attention.log | head -30
KB_TAG=r8/kernel_bench timeout 600 python tools/kernel_bench.py attn > gpurun_out/r8/kernel_bench.txt 2>&1
echo "kernel_bench rc=$?"; cat gpurun_out/r8/kernel_bench.txt | tail -12
