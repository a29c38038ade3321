#!/bin/bash
mkdir -p gpurun_out/r26
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "cnn or gemm" --tb=short > gpurun_out/r26/k.log 2>&1
echo "kernels rc=$? $(tail -1 gpurun_out/r26/k.log)"; grep -E "^E  |FAILED" gpurun_out/r26/k.log | head -10
for occ in 3 2 3 2; do
CALM_CNN_BWD_OCC=$occ KB_TAG=r26/kb_cnn_$occ timeout 600 python tools/kernel_bench.py cnn > gpurun_out/r26/kb_cnn_$occ.txt 2>&1
echo "occ $occ"; grep bwd gpurun_out/r26/kb_cnn_$occ.txt
done
grep fwd gpurun_out/r26/kb_cnn_3.txt
KB_TAG=r26/kernel_bench timeout 600 python tools/kernel_bench.py gemm > gpurun_out/r26/kernel_bench.txt 2>&1
grep -E "GELU|S224" gpurun_out/r26/kernel_bench.txt | cut -c1-110
