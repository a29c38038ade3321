#!/bin/bash
mkdir -p gpurun_out/r16
timeout 900 python tools/res_check.py > gpurun_out/r16/res_check.log 2>&1; echo "res_check rc=$?"; tail -4 gpurun_out/r16/res_check.log
timeout 900 python bench.py --task reg --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r16/bench_reg.json 2> gpurun_out/r16/bench_reg.err
echo "bench reg rc=$?"; head -c 250 gpurun_out/r16/bench_reg.json; tail -3 gpurun_out/r16/bench_reg.err
