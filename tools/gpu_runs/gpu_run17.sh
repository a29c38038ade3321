#!/bin/bash
mkdir -p gpurun_out/r17
CMD="python tools/kernel_bench.py cnn"
KB_BATCH=64 $CMD > gpurun_out/r17/plain_cnn.log 2>&1 &&
KB_BATCH=64 ncu --set full --clock-control none --import-source on -k regex:cnn_bwd -s 4 -c 1 -o gpurun_out/r17/prof_cnn_bwd $CMD > gpurun_out/r17/ncu_cnn.log 2>&1
echo "ncu cnn rc=$?"
python tools/ncu_extract.py gpurun_out/r17/prof_cnn_bwd.ncu-rep > gpurun_out/r17/ncu_cnn_bwd_metrics.txt 2>&1
cat gpurun_out/r17/ncu_cnn_bwd_metrics.txt
