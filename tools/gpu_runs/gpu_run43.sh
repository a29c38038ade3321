#!/bin/bash
mkdir -p gpurun_out/r43
CMD="python tools/kernel_bench.py gemm"
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 605 -c 1 -o /tmp/prof_gemm_seq $CMD > gpurun_out/r43/ncu.log 2>&1
echo "ncu rc=$?"
python tools/ncu_extract.py /tmp/prof_gemm_seq.ncu-rep > gpurun_out/r43/gemm_seq_metrics.txt 2>&1
python tools/ncu_source_lines.py /tmp/prof_gemm_seq.ncu-rep 45 > gpurun_out/r43/gemm_seq_lines.txt 2>&1
head -24 gpurun_out/r43/gemm_seq_metrics.txt
