#!/bin/bash
# Round-end style validation: full GPU test suite, smoke, bench (graph), ncu launch list of one step.
mkdir -p gpurun_out/final
timeout 1200 python -m pytest tests -q -m gpu --tb=short > gpurun_out/final/pytest_gpu.log 2>&1
echo "pytest -m gpu rc=$? $(tail -1 gpurun_out/final/pytest_gpu.log)"
timeout 300 python __graft_entry__.py --smoke > gpurun_out/final/smoke.log 2>&1
echo "smoke rc=$? $(tail -1 gpurun_out/final/smoke.log)"
timeout 900 python bench.py > gpurun_out/final/bench.json 2> gpurun_out/final/bench.err
echo "bench rc=$?"; cat gpurun_out/final/bench.json; cp gpurun_out/bench_kernel_breakdown.json gpurun_out/bench_gemm_shapes.json gpurun_out/final/
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final/bench_reference.json 2> gpurun_out/final/bench_reference.err
echo "reference rc=$?"; cat gpurun_out/final/bench_reference.json
KB_TAG=final/kernel_bench timeout 900 python tools/kernel_bench.py > gpurun_out/final/kernel_bench.txt 2>&1
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-profile --no-cpu-baseline"
$CMD > gpurun_out/final/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 11000 -c 2200 --csv --log-file gpurun_out/final/launches.csv $CMD > gpurun_out/final/ncu_launches.log 2>&1
echo "ncu launch list rc=$?"; du -sh gpurun_out
