#!/bin/bash
mkdir -p gpurun_out/r55
KB_AB=0,8 KB_TAG=r55/kb_ab timeout 800 python tools/kernel_bench.py gemm > gpurun_out/r55/kb_ab.txt 2>&1
grep "^gemm" gpurun_out/r55/kb_ab.txt | cut -c10-140
