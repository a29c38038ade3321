#!/bin/bash
mkdir -p gpurun_out/r52
KB_BATCH=64 ncu --set full --clock-control none --import-source on -k regex:cnn_fwd -s 4 -c 1 -o /tmp/prof_cnn_fwd python tools/kernel_bench.py cnn > gpurun_out/r52/ncu.log 2>&1
echo "ncu rc=$?"
python tools/ncu_extract.py /tmp/prof_cnn_fwd.ncu-rep > gpurun_out/r52/cnn_fwd_metrics.txt 2>&1
python tools/ncu_source_lines.py /tmp/prof_cnn_fwd.ncu-rep 45 > gpurun_out/r52/cnn_fwd_lines.txt 2>&1
grep -E "time_duration|issue_active|pipe_|inst_executed.sum|registers" gpurun_out/r52/cnn_fwd_metrics.txt
