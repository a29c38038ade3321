#!/bin/bash
mkdir -p gpurun_out/r31
for f in 0 0x1000 0x2000 0x4000 0x7000; do
KB_FLAGS=$f KB_TAG=r31/kb_$f timeout 300 python tools/kernel_bench.py attn > gpurun_out/r31/kb_$f.txt 2>&1
echo "flags $f"; grep "bwd" gpurun_out/r31/kb_$f.txt | cut -c1-80
done
