#!/bin/bash
mkdir -p gpurun_out/r75
timeout 1200 python -m pytest tests -q -m gpu --tb=short > gpurun_out/r75/pytest_gpu.log 2>&1
echo "pytest -m gpu rc=$? $(tail -1 gpurun_out/r75/pytest_gpu.log)"
grep -E "^(FAILED|ERROR|E  )" gpurun_out/r75/pytest_gpu.log | head -20
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r75/smoke.log 2>&1
echo "smoke rc=$? $(tail -2 gpurun_out/r75/smoke.log)"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r75/bench.json 2> gpurun_out/r75/bench.err
echo "bench rc=$?"; cut -c1-330 gpurun_out/r75/bench.json; tail -3 gpurun_out/r75/bench.err
cp gpurun_out/bench_kernel_breakdown.json gpurun_out/bench_gemm_shapes.json gpurun_out/r75/ 2>/dev/null
