#!/bin/bash
mkdir -p gpurun_out/r14
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "spectral or rope or token or layernorm" --tb=short > gpurun_out/r14/kernels.log 2>&1
echo "kernels rc=$? $(tail -1 gpurun_out/r14/kernels.log)"
timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short > gpurun_out/r14/model.log 2>&1
echo "model rc=$? $(tail -1 gpurun_out/r14/model.log)"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r14/bench.json 2> gpurun_out/r14/bench.err
echo "bench rc=$?"; head -c 220 gpurun_out/r14/bench.json; cp gpurun_out/bench_kernel_breakdown.json gpurun_out/r14/breakdown.json
