#!/bin/bash
mkdir -p gpurun_out/r37
CMD="python tools/kernel_bench.py gemm"
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 40 -c 1 -o /tmp/prof_gemm_gelu $CMD > gpurun_out/r37/ncu.log 2>&1
echo "ncu rc=$?"
python tools/ncu_extract.py /tmp/prof_gemm_gelu.ncu-rep > gpurun_out/r37/gemm_gelu_metrics.txt 2>&1
python tools/ncu_source_lines.py /tmp/prof_gemm_gelu.ncu-rep 50 > gpurun_out/r37/gemm_gelu_lines.txt 2>&1
head -26 gpurun_out/r37/gemm_gelu_metrics.txt
