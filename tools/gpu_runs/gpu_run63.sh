#!/bin/bash
mkdir -p gpurun_out/r63
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_trainer_gpu.py -q --tb=short -k "rope or graph" > gpurun_out/r63/pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/r63/pytest.log)"
grep -E "^(FAILED|ERROR|E  )" gpurun_out/r63/pytest.log | head -20
timeout 600 python bench.py --no-profile --no-cpu-baseline > gpurun_out/r63/bench.json 2> gpurun_out/r63/bench.err
echo "bench rc=$?"; cut -c1-330 gpurun_out/r63/bench.json; tail -3 gpurun_out/r63/bench.err
CMD="python tools/kernel_bench.py misc"
KB_TAG=r63/kb timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rope_ --csv --log-file gpurun_out/r63/rope_launches.csv $CMD > gpurun_out/r63/ncu.log 2>&1
echo "ncu rc=$?"
