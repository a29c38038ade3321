#!/bin/bash
mkdir -p gpurun_out/r24
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k cnn --tb=short > gpurun_out/r24/k_cnn.log 2>&1
echo "kernels:cnn rc=$? $(tail -1 gpurun_out/r24/k_cnn.log)"; grep -E "^E  |FAILED" gpurun_out/r24/k_cnn.log | head -10
KB_TAG=r24/kernel_bench timeout 600 python tools/kernel_bench.py cnn > gpurun_out/r24/kernel_bench.txt 2>&1
echo "kernel_bench rc=$?"; cat gpurun_out/r24/kernel_bench.txt
KB_BATCH=64 ncu --set full --clock-control none --import-source on -k regex:cnn_bwd -s 4 -c 1 -o /tmp/prof_cnn_bwd python tools/kernel_bench.py cnn > gpurun_out/r24/ncu_cnn.log 2>&1
echo "ncu cnn rc=$?"
python tools/ncu_extract.py /tmp/prof_cnn_bwd.ncu-rep > gpurun_out/r24/ncu_cnn_bwd_metrics.txt 2>&1
python tools/ncu_source_lines.py /tmp/prof_cnn_bwd.ncu-rep 80 > gpurun_out/r24/cnn_bwd_lines.txt 2>&1
head -30 gpurun_out/r24/ncu_cnn_bwd_metrics.txt
