#!/bin/bash
mkdir -p gpurun_out/r71
timeout 600 python -m pytest tests/test_trainer_gpu.py -q --tb=short > gpurun_out/r71/pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/r71/pytest.log)"
timeout 600 python bench.py --no-profile --no-cpu-baseline --task reg > gpurun_out/r71/bench_reg.json 2> gpurun_out/r71/bench_reg.err
echo "bench reg rc=$?"; cut -c1-300 gpurun_out/r71/bench_reg.json; tail -3 gpurun_out/r71/bench_reg.err
