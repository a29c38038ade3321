#!/bin/bash
mkdir -p gpurun_out/r56
KB_BATCH=256 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc -s 4 -c 1 -o /tmp/prof_attn_bwd python tools/kernel_bench.py attn > gpurun_out/r56/ncu_attn.log 2>&1
echo "ncu rc=$?"
python tools/ncu_extract.py /tmp/prof_attn_bwd.ncu-rep > gpurun_out/r56/attn_bwd_metrics.txt 2>&1
python tools/ncu_source_lines.py /tmp/prof_attn_bwd.ncu-rep 70 > gpurun_out/r56/attn_bwd_lines.txt 2>&1
head -28 gpurun_out/r56/attn_bwd_metrics.txt
