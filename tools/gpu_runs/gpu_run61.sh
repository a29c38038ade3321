#!/bin/bash
# row-staged RoPE kernels: parity + per-kernel bench; trainer graph test
mkdir -p gpurun_out/r61
timeout 600 python -m pytest tests/test_trainer_gpu.py tests/test_kernels_gpu.py -q --tb=short -k "rope or trainer or cross_entropy or huber or state_dict" > gpurun_out/r61/pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/r61/pytest.log)"
grep -E "^(FAILED|ERROR|E  )" gpurun_out/r61/pytest.log | head -30
KB_TAG=r61/kb_misc timeout 600 python tools/kernel_bench.py misc > gpurun_out/r61/kb_misc.txt 2>&1
grep "rope" gpurun_out/r61/kb_misc.txt
