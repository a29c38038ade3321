#!/bin/bash
# ncu --set full of the row-staged RoPE kernels (first two launches = S224 fwd warm-up calls of kernel_bench misc)
mkdir -p gpurun_out/r62
CMD="python tools/kernel_bench.py misc"
KB_TAG=r62/kb timeout 300 $CMD > gpurun_out/r62/plain.log 2>&1 || exit 1
KB_TAG=r62/kb timeout 600 ncu --set full --clock-control none --import-source on -k regex:rope_rows -c 2 -o gpurun_out/r62/prof_rope_fwd $CMD > gpurun_out/r62/ncu_fwd.log 2>&1
echo "ncu fwd rc=$?"
KB_TAG=r62/kb timeout 600 ncu --set full --clock-control none --import-source on -k regex:rope_rows_bwd -c 1 -o gpurun_out/r62/prof_rope_bwd $CMD > gpurun_out/r62/ncu_bwd.log 2>&1
echo "ncu bwd rc=$?"
ls -la gpurun_out/r62
