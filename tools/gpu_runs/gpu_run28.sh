#!/bin/bash
mkdir -p gpurun_out/r28
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention or attn" --tb=short > gpurun_out/r28/k_attn.log 2>&1
echo "kernels:attn rc=$? $(tail -1 gpurun_out/r28/k_attn.log)"; grep -E "^E  |FAILED" gpurun_out/r28/k_attn.log | head -10
KB_TAG=r28/kernel_bench timeout 600 python tools/kernel_bench.py attn > gpurun_out/r28/kernel_bench.txt 2>&1
echo "kernel_bench rc=$?"; cat gpurun_out/r28/kernel_bench.txt
python tools/attn_trace.py 224 > gpurun_out/r28/trace224.txt 2>&1; head -60 gpurun_out/r28/trace224.txt | tail -32
