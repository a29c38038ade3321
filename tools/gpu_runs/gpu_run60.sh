#!/bin/bash
# trainer-glue parity + A/B of the step with torch glue vs calm_trainer glue
mkdir -p gpurun_out/r60
timeout 600 python -m pytest tests/test_trainer_gpu.py -q --tb=short > gpurun_out/r60/pytest_trainer.log 2>&1
echo "pytest trainer rc=$? $(tail -1 gpurun_out/r60/pytest_trainer.log)"
grep -E "^(FAILED|ERROR|E  )" gpurun_out/r60/pytest_trainer.log | head -30
timeout 600 python bench.py --no-profile --no-cpu-baseline --torch-glue > gpurun_out/r60/bench_torch_glue.json 2> gpurun_out/r60/bench_torch_glue.err
echo "bench torch-glue rc=$?"; cut -c1-400 gpurun_out/r60/bench_torch_glue.json
timeout 600 python bench.py --no-profile --no-cpu-baseline > gpurun_out/r60/bench_calm_glue.json 2> gpurun_out/r60/bench_calm_glue.err
echo "bench calm-glue rc=$?"; cut -c1-400 gpurun_out/r60/bench_calm_glue.json; tail -5 gpurun_out/r60/bench_calm_glue.err
timeout 600 python bench.py --no-profile --no-cpu-baseline --task reg > gpurun_out/r60/bench_calm_glue_reg.json 2> gpurun_out/r60/bench_calm_glue_reg.err
echo "bench reg rc=$?"; cut -c1-300 gpurun_out/r60/bench_calm_glue_reg.json; tail -5 gpurun_out/r60/bench_calm_glue_reg.err
