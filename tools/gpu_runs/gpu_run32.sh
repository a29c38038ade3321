#!/bin/bash
mkdir -p gpurun_out/r32
timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short -x > gpurun_out/r32/model.log 2>&1
echo "model tests rc=$? $(tail -1 gpurun_out/r32/model.log)"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r32/bench.json 2> gpurun_out/r32/bench.err
echo "bench rc=$?"; head -c 250 gpurun_out/r32/bench.json; tail -2 gpurun_out/r32/bench.err
cp gpurun_out/bench_kernel_breakdown.json gpurun_out/r32/
