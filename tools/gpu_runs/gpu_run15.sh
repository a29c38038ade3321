#!/bin/bash
mkdir -p gpurun_out/r15
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "spectral or attention" --tb=short > gpurun_out/r15/kernels.log 2>&1
echo "kernels rc=$? $(tail -1 gpurun_out/r15/kernels.log)"
KB_TAG=r15/kernel_bench timeout 600 python tools/kernel_bench.py attn > gpurun_out/r15/kernel_bench.txt 2>&1
echo "kernel_bench rc=$?"; cat gpurun_out/r15/kernel_bench.txt | tail -8
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r15/bench.json 2> gpurun_out/r15/bench.err
echo "bench rc=$?"; head -c 220 gpurun_out/r15/bench.json; cp gpurun_out/bench_kernel_breakdown.json gpurun_out/r15/breakdown.json
