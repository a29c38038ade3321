#!/bin/bash
mkdir -p gpurun_out/r7
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short > gpurun_out/r7/kernels.log 2>&1
echo "kernels rc=$? $(tail -1 gpurun_out/r7/kernels.log)"
KB_TAG=r7/kernel_bench timeout 900 python tools/kernel_bench.py attn misc > gpurun_out/r7/kernel_bench.txt 2>&1
echo "kernel_bench rc=$?"; grep -E "attn|rope_fwd|colsum" gpurun_out/r7/kernel_bench.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r7/bench.json 2> gpurun_out/r7/bench.err
echo "bench rc=$?"; cat gpurun_out/r7/bench.json | head -c 300; cp gpurun_out/bench_kernel_breakdown.json gpurun_out/r7/breakdown.json
