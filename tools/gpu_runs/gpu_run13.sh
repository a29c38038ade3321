#!/bin/bash
mkdir -p gpurun_out/r13
CMD="python tools/kernel_bench.py gemm"
$CMD > gpurun_out/r13/plain_gemm.log 2>&1 &&
ncu --set full --clock-control none -k regex:gemm_tcgen05 -s 10 -c 36 -o /tmp/prof_gemm $CMD > gpurun_out/r13/ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
python tools/ncu_extract.py /tmp/prof_gemm.ncu-rep > gpurun_out/r13/ncu_gemm_metrics.txt 2>&1
CMD2="python tools/kernel_bench.py attn"
$CMD2 > gpurun_out/r13/plain_attn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_.*tc_kernel -s 6 -c 2 -o gpurun_out/r13/prof_attn $CMD2 > gpurun_out/r13/ncu_attn.log 2>&1
echo "ncu attn rc=$?"
python tools/ncu_extract.py gpurun_out/r13/prof_attn.ncu-rep > gpurun_out/r13/ncu_attn_metrics.txt 2>&1
ls -la gpurun_out/r13; du -sh gpurun_out
