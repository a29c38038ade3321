#!/bin/bash
mkdir -p gpurun_out/r48
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "sn or spectral" --tb=short > gpurun_out/r48/k_sn.log 2>&1
echo "kernels:sn rc=$? $(tail -1 gpurun_out/r48/k_sn.log)"
timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short -x > gpurun_out/r48/model.log 2>&1
echo "model tests rc=$? $(tail -1 gpurun_out/r48/model.log)"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r48/bench.json 2> gpurun_out/r48/bench.err
echo "bench rc=$?"; head -c 220 gpurun_out/r48/bench.json; tail -2 gpurun_out/r48/bench.err
cp gpurun_out/bench_kernel_breakdown.json gpurun_out/r48/
