#!/bin/bash
mkdir -p gpurun_out/r42
timeout 900 python -m pytest tests -q -m gpu --tb=short > gpurun_out/r42/k.log 2>&1
echo "gpu tests rc=$? $(tail -1 gpurun_out/r42/k.log)"; grep -E "^E  |FAILED" gpurun_out/r42/k.log | head -10
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r42/bench.json 2> gpurun_out/r42/bench.err
echo "bench rc=$?"; head -c 220 gpurun_out/r42/bench.json; tail -2 gpurun_out/r42/bench.err
cp gpurun_out/bench_kernel_breakdown.json gpurun_out/bench_gemm_shapes.json gpurun_out/r42/
