#!/bin/bash
mkdir -p gpurun_out/r44
KB_AB=0,0x400 KB_TAG=r44/kb_ab timeout 800 python tools/kernel_bench.py gemm > gpurun_out/r44/kb_ab.txt 2>&1
grep "^gemm" gpurun_out/r44/kb_ab.txt | cut -c10-140
