#!/bin/bash
mkdir -p gpurun_out/r20
for f in 0 0x100 0x200 0x300; do
KB_FLAGS=$f KB_TAG=r20/kb_$f timeout 600 python tools/kernel_bench.py gemm > gpurun_out/r20/kb_$f.txt 2>&1
echo "flags $f rc=$?"; grep -E "S224" gpurun_out/r20/kb_$f.txt | cut -c1-110
done
