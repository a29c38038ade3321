#!/bin/bash
mkdir -p gpurun_out/r4
for k in gemm cnn; do
  timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k $k --tb=short > gpurun_out/r4/k_$k.log 2>&1
  echo "kernels:$k rc=$? $(tail -1 gpurun_out/r4/k_$k.log)"
done
KB_TAG=r4/kernel_bench timeout 900 python tools/kernel_bench.py gemm cnn > gpurun_out/r4/kernel_bench.txt 2>&1
echo "kernel_bench rc=$?"
timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short -s > gpurun_out/r4/model.log 2>&1
echo "model rc=$? $(tail -1 gpurun_out/r4/model.log)"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r4/bench.json 2> gpurun_out/r4/bench.err
echo "bench rc=$?"; cat gpurun_out/r4/bench.json | head -c 1800; cp gpurun_out/bench_kernel_breakdown.json gpurun_out/r4/breakdown.json
