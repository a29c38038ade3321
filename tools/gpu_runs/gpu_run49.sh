#!/bin/bash
mkdir -p gpurun_out/r49
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "rope" --tb=short > gpurun_out/r49/k.log 2>&1
echo "kernels:rope rc=$? $(tail -1 gpurun_out/r49/k.log)"; grep -E "^E  |FAILED" gpurun_out/r49/k.log | head
KB_TAG=r49/kernel_bench timeout 600 python tools/kernel_bench.py misc > gpurun_out/r49/kernel_bench.txt 2>&1
grep rope gpurun_out/r49/kernel_bench.txt
