#!/bin/bash
mkdir -p gpurun_out/r65
timeout 1200 python -m pytest tests -q -m gpu --tb=short > gpurun_out/r65/pytest_gpu.log 2>&1
echo "pytest -m gpu rc=$? $(tail -1 gpurun_out/r65/pytest_gpu.log)"
grep -E "^(FAILED|ERROR|E  )" gpurun_out/r65/pytest_gpu.log | head -20
timeout 600 python bench.py --no-profile --no-cpu-baseline > gpurun_out/r65/bench.json 2> gpurun_out/r65/bench.err
echo "bench rc=$?"; cut -c1-330 gpurun_out/r65/bench.json; tail -3 gpurun_out/r65/bench.err
