#!/bin/bash
mkdir -p gpurun_out/r74
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q --tb=short -k "latent or matches_oracle or mutation or rng" > gpurun_out/r74/pytest.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/r74/pytest.log)"
grep -E "^(FAILED|ERROR|E  )" gpurun_out/r74/pytest.log | head -20
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-profile --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:latent_ -c 30 --csv --log-file gpurun_out/r74/latent_launches.csv $CMD > gpurun_out/r74/ncu.log 2>&1
echo "ncu rc=$?"
