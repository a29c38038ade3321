#!/bin/bash
mkdir -p gpurun_out/r68
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_kernels_gpu.py -q -s --tb=short -k "highres or colsum" > gpurun_out/r68/pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/r68/pytest.log)"
grep -E "^(FAILED|ERROR|E  )|out rel" gpurun_out/r68/pytest.log | head -20
