#!/bin/bash
mkdir -p gpurun_out/r33
KB_AB=0,64,32 KB_TAG=r33/kb_ab timeout 800 python tools/kernel_bench.py gemm > gpurun_out/r33/kb_ab.txt 2>&1
echo "rc=$?"; cat gpurun_out/r33/kb_ab.txt
