#!/bin/bash
mkdir -p gpurun_out/r39
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "rope or cnn or gemm" --tb=short > gpurun_out/r39/k.log 2>&1
echo "kernels rc=$? $(tail -1 gpurun_out/r39/k.log)"; grep -E "^E  |FAILED" gpurun_out/r39/k.log | head -10
KB_TAG=r39/kernel_bench timeout 600 python tools/kernel_bench.py misc cnn > gpurun_out/r39/kernel_bench.txt 2>&1
echo "kernel_bench rc=$?"; grep -E "rope|cnn" gpurun_out/r39/kernel_bench.txt
