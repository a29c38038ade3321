#!/bin/bash
mkdir -p gpurun_out/r22
CMD="python tools/kernel_bench.py gemm"
$CMD > gpurun_out/r22/plain.log 2>&1 || { echo "plain failed"; exit 0; }
for idx in 10 25 40; do
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s $idx -c 1 -o /tmp/prof_gemm_$idx $CMD > gpurun_out/r22/ncu_$idx.log 2>&1
echo "ncu $idx rc=$?"
python tools/ncu_extract.py /tmp/prof_gemm_$idx.ncu-rep pct_of_peak_sustained_elapsed stalled lts__t_ l1tex__m_ smsp__pcsamp > gpurun_out/r22/gemm_$idx.txt 2>&1
ncu -i /tmp/prof_gemm_$idx.ncu-rep --page details > gpurun_out/r22/gemm_${idx}_details.txt 2>&1
done
ls -la /tmp/*.ncu-rep
