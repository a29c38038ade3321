#!/bin/bash
mkdir -p gpurun_out/r67
timeout 900 python -m pytest tests/test_trainer_gpu.py -q --tb=short > gpurun_out/r67/pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/r67/pytest.log)"
grep -E "^(FAILED|ERROR|E  )" gpurun_out/r67/pytest.log | head -20
