#!/bin/bash
mkdir -p gpurun_out/r69
timeout 600 python -m pytest tests/test_model_gpu.py -q -s --tb=line -k "highres and 384" 2>&1 | grep -E "distance|out rel|passed|failed" > gpurun_out/r69/new.txt
CALM_DEBUG_FLAGS=64 timeout 600 python -m pytest tests/test_model_gpu.py -q -s --tb=line -k "highres and 384" 2>&1 | grep -E "distance|out rel|passed|failed" > gpurun_out/r69/legacy_rope.txt
echo new; cat gpurun_out/r69/new.txt; echo legacy rope; cat gpurun_out/r69/legacy_rope.txt
