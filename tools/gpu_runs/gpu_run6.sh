#!/bin/bash
mkdir -p gpurun_out/r6
for k in gemm attention; do
  timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k $k --tb=short > gpurun_out/r6/k_$k.log 2>&1
  echo "kernels:$k rc=$? $(tail -1 gpurun_out/r6/k_$k.log)"
done
KB_TAG=r6/kernel_bench timeout 900 python tools/kernel_bench.py gemm attn > gpurun_out/r6/kernel_bench.txt 2>&1
echo "kernel_bench rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r6/bench.json 2> gpurun_out/r6/bench.err
echo "bench rc=$?"; cat gpurun_out/r6/bench.json | head -c 400; cp gpurun_out/bench_kernel_breakdown.json gpurun_out/r6/breakdown.json
