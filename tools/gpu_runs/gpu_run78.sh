#!/bin/bash
mkdir -p gpurun_out/r78
timeout 600 python -m pytest tests -q -m gpu --tb=short -x > gpurun_out/r78/pytest_gpu.log 2>&1
echo "pytest -m gpu rc=$? $(tail -1 gpurun_out/r78/pytest_gpu.log)"
grep -E "^(FAILED|ERROR|E  )" gpurun_out/r78/pytest_gpu.log | head -10
timeout 300 python bench.py --no-profile --no-cpu-baseline > gpurun_out/r78/bench.json 2> gpurun_out/r78/bench.err
echo "bench rc=$?"; cut -c1-200 gpurun_out/r78/bench.json
