#!/bin/bash
# call 3: parity suite, smoke, kernel micro-benchmarks, ncu launch list + full capture of the GEMM kernel
mkdir -p gpurun_out/r3
timeout 900 python -m pytest tests -q -m gpu --tb=short -s -x > gpurun_out/r3/pytest_gpu.log 2>&1
echo "pytest -m gpu rc=$? $(tail -1 gpurun_out/r3/pytest_gpu.log)"
grep -E "^\[small" gpurun_out/r3/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r3/smoke.log 2>&1
echo "smoke rc=$? $(tail -1 gpurun_out/r3/smoke.log)"
KB_TAG=r3/kernel_bench timeout 900 python tools/kernel_bench.py > gpurun_out/r3/kernel_bench.txt 2>&1
echo "kernel_bench rc=$?"
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-profile --no-cpu-baseline"
$CMD > gpurun_out/r3/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 1700 --csv --log-file gpurun_out/r3/launches.csv $CMD > gpurun_out/r3/ncu_launches.log 2>&1
echo "ncu launch list rc=$?"
$CMD > gpurun_out/r3/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 200 -c 3 -o gpurun_out/r3/prof_gemm $CMD > gpurun_out/r3/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out/r3
