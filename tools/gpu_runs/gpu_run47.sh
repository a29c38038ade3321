#!/bin/bash
mkdir -p gpurun_out/r47
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k cnn --tb=short > gpurun_out/r47/k_cnn.log 2>&1
echo "kernels:cnn rc=$? $(tail -1 gpurun_out/r47/k_cnn.log)"
KB_TAG=r47/kernel_bench timeout 600 python tools/kernel_bench.py cnn > gpurun_out/r47/kernel_bench.txt 2>&1
cat gpurun_out/r47/kernel_bench.txt
