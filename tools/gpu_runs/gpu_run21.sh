#!/bin/bash
mkdir -p gpurun_out/r21
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "epilogue_forms" --tb=short -x > gpurun_out/r21/k_epi.log 2>&1
rc=$?; echo "epi rc=$rc $(tail -1 gpurun_out/r21/k_epi.log)"; grep -E "^E  |FAILED" gpurun_out/r21/k_epi.log | head -10
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k gemm --tb=short > gpurun_out/r21/k_gemm.log 2>&1
echo "kernels:gemm rc=$? $(tail -1 gpurun_out/r21/k_gemm.log)"; grep -E "^E  |FAILED" gpurun_out/r21/k_gemm.log | head -20
KB_TAG=r21/kernel_bench timeout 600 python tools/kernel_bench.py gemm > gpurun_out/r21/kernel_bench.txt 2>&1
echo "kernel_bench rc=$?"; grep -E "S224|S176 qkv|S128 qkv|S80 qkv" gpurun_out/r21/kernel_bench.txt | cut -c1-110
if [ $rc -ne 0 ]; then exit 0; fi
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r21/bench.json 2> gpurun_out/r21/bench.err
echo "bench rc=$?"; head -c 220 gpurun_out/r21/bench.json; tail -2 gpurun_out/r21/bench.err
