#!/bin/bash
mkdir -p gpurun_out/r9
for a in "2 64 12 4 bwd" "2 32 12 20 bwd" "2 176 12 44 bwd" "2 80 12 20 bwd"; do echo "== $a"; timeout 60 python tools/attn_debug.py $a 2>&1 | tail -2; done
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k attention --tb=short > gpurun_out/r9/k_attention.log 2>&1
echo "kernels:attention rc=$? $(tail -1 gpurun_out/r9/k_attention.log)"
KB_TAG=r9/kernel_bench timeout 600 python tools/kernel_bench.py attn > gpurun_out/r9/kernel_bench.txt 2>&1
echo "kernel_bench rc=$?"; cat gpurun_out/r9/kernel_bench.txt | tail -9
timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short -s > gpurun_out/r9/model.log 2>&1
echo "model rc=$? $(tail -1 gpurun_out/r9/model.log)"; grep -E "^\[small" gpurun_out/r9/model.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r9/bench.json 2> gpurun_out/r9/bench.err
echo "bench rc=$?"; cat gpurun_out/r9/bench.json | head -c 300; cp gpurun_out/bench_kernel_breakdown.json gpurun_out/r9/breakdown.json
