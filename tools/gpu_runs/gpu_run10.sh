#!/bin/bash
mkdir -p gpurun_out/r10
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k attention --tb=short > gpurun_out/r10/k_attention.log 2>&1
echo "kernels:attention rc=$? $(tail -1 gpurun_out/r10/k_attention.log)"
KB_TAG=r10/kernel_bench timeout 600 python tools/kernel_bench.py attn > gpurun_out/r10/kernel_bench.txt 2>&1
echo "kernel_bench rc=$?"; cat gpurun_out/r10/kernel_bench.txt | tail -9
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r10/bench.json 2> gpurun_out/r10/bench.err
echo "bench rc=$?"; cat gpurun_out/r10/bench.json | head -c 300; cp gpurun_out/bench_kernel_breakdown.json gpurun_out/r10/breakdown.json
