#!/bin/bash
# same-box sweep of the spectral-norm work-item size
mkdir -p gpurun_out/r77
for w in 16384 8192 4096 16384 8192; do
CALM_SN_ITEM_WEIGHTS=$w timeout 300 python bench.py --no-profile --no-cpu-baseline --steps 10 > gpurun_out/r77/bench_$w.json 2> gpurun_out/r77/bench_$w.err
echo "$w rc=$? $(python -c "import json;b=json.load(open('gpurun_out/r77/bench_$w.json'));print(b['ms_per_step'])")"
done
