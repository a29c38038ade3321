#!/bin/bash
# same-box A/B: bf16 side outputs of LN backward / token transpose / CNN backward on vs off (consumer-side casts)
mkdir -p gpurun_out/r73
for rep in 1 2; do
CALM_BF16_SIDE=0 timeout 600 python bench.py --no-profile --no-cpu-baseline > gpurun_out/r73/bench_off_$rep.json 2> gpurun_out/r73/bench_off_$rep.err
echo "off rc=$? $(cut -c1-160 gpurun_out/r73/bench_off_$rep.json)"
CALM_BF16_SIDE=1 timeout 600 python bench.py --no-profile --no-cpu-baseline > gpurun_out/r73/bench_on_$rep.json 2> gpurun_out/r73/bench_on_$rep.err
echo "on  rc=$? $(cut -c1-160 gpurun_out/r73/bench_on_$rep.json)"
done
