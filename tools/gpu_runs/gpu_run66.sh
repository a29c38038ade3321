#!/bin/bash
mkdir -p gpurun_out/r66
timeout 600 python -m pytest tests/test_model_gpu.py -q -s -k "matches_oracle" --tb=line 2>&1 | grep -E "output rel err|passed|failed" > gpurun_out/r66/new.txt
CALM_DEBUG_FLAGS=64 timeout 600 python -m pytest tests/test_model_gpu.py -q -s -k "matches_oracle" --tb=line 2>&1 | grep -E "output rel err|passed|failed" > gpurun_out/r66/legacy.txt
echo new; cat gpurun_out/r66/new.txt; echo legacy; cat gpurun_out/r66/legacy.txt
