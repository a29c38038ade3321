#!/bin/bash
mkdir -p gpurun_out/r46
KB_BATCH=64 ncu --set full --clock-control none --import-source on -k regex:cnn_bwd -s 4 -c 1 -o /tmp/prof_cnn_bwd python tools/kernel_bench.py cnn > gpurun_out/r46/ncu_cnn.log 2>&1
echo "ncu cnn rc=$?"
python tools/ncu_source_lines.py /tmp/prof_cnn_bwd.ncu-rep 90 > gpurun_out/r46/cnn_bwd_lines.txt 2>&1
head -5 gpurun_out/r46/cnn_bwd_lines.txt
