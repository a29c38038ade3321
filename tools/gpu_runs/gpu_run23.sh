#!/bin/bash
mkdir -p gpurun_out/r23
CMD="python tools/kernel_bench.py gemm"
for idx in 10 25; do
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s $idx -c 1 -o /tmp/prof_gemm_$idx $CMD > gpurun_out/r23/ncu_$idx.log 2>&1
echo "ncu $idx rc=$?"
python tools/ncu_source_top.py /tmp/prof_gemm_$idx.ncu-rep 60 > gpurun_out/r23/src_top_$idx.txt 2>&1
ncu -i /tmp/prof_gemm_$idx.ncu-rep --page source --csv > gpurun_out/r23/src_$idx.csv 2>/dev/null
done
ls -la gpurun_out/r23
