#!/bin/bash
mkdir -p gpurun_out/r34
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short > gpurun_out/r34/k.log 2>&1
echo "kernels rc=$? $(tail -1 gpurun_out/r34/k.log)"; grep -E "^E  |FAILED" gpurun_out/r34/k.log | head -10
KB_TAG=r34/kernel_bench timeout 600 python tools/kernel_bench.py ln misc > gpurun_out/r34/kernel_bench.txt 2>&1
echo "kernel_bench rc=$?"; cat gpurun_out/r34/kernel_bench.txt
