#!/bin/bash
mkdir -p gpurun_out/r79
for w in 4096 2048 1024 4096; do
CALM_SN_ITEM_WEIGHTS=$w timeout 120 python bench.py --no-profile --no-cpu-baseline --steps 10 > gpurun_out/r79/bench_$w.json 2> gpurun_out/r79/bench_$w.err
echo "$w rc=$? $(python -c "import json;b=json.load(open('gpurun_out/r79/bench_$w.json'));print(b['ms_per_step'])")"
done
