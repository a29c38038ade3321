#!/bin/bash
mkdir -p gpurun_out/r12
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r12/bench.json 2> gpurun_out/r12/bench.err
echo "bench rc=$?"; head -c 200 gpurun_out/r12/bench.json; echo; cp gpurun_out/bench_gemm_shapes.json gpurun_out/bench_kernel_breakdown.json gpurun_out/r12/
CMD="python tools/kernel_bench.py cnn"
KB_BATCH=64 $CMD > gpurun_out/r12/plain_cnn.log 2>&1 &&
KB_BATCH=64 ncu --set full --clock-control none --import-source on -k regex:cnn_ -s 8 -c 2 -o gpurun_out/r12/prof_cnn $CMD > gpurun_out/r12/ncu_cnn.log 2>&1
echo "ncu cnn rc=$?"
