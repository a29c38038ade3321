#!/bin/bash
# Round-end style validation on a B200 box (one gpurun call): full GPU test suite, smoke, bench (all configs), the reference arm,
# per-kernel micro-benchmarks, and the ncu launch list of one step (durations + dram bytes + tensor-pipe activity).
# Usage: bash tools/gpu_final.sh [tag]   -> gpurun_out/<tag>/
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu --tb=short -s > $OUT/pytest_gpu.log 2>&1
echo "pytest -m gpu rc=$? $(tail -1 $OUT/pytest_gpu.log)"
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1
echo "smoke rc=$? $(tail -1 $OUT/smoke.log)"
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?"; head -c 700 $OUT/bench.json; echo; cp gpurun_out/bench_kernel_breakdown.json gpurun_out/bench_gemm_shapes.json $OUT/
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
echo "reference rc=$?"; head -c 400 $OUT/bench_reference.json; echo
if [ "$2" != "short" ]; then
  timeout 600 python bench.py --task reg --no-cpu-baseline > $OUT/bench_reg.json 2> $OUT/bench_reg.err; echo "reg rc=$?"; head -c 300 $OUT/bench_reg.json; echo
  timeout 900 python bench.py --res 384 --no-cpu-baseline > $OUT/bench_384.json 2> $OUT/bench_384.err; echo "384 rc=$?"; head -c 300 $OUT/bench_384.json; echo
  timeout 900 python bench.py --res 512 --no-cpu-baseline > $OUT/bench_512.json 2> $OUT/bench_512.err; echo "512 rc=$?"; head -c 300 $OUT/bench_512.json; echo
  KB_TAG=$TAG/kernel_bench timeout 900 python tools/kernel_bench.py > $OUT/kernel_bench.txt 2>&1
fi
CMD="python bench.py --steps 1 --warmup 3 --no-graph --no-profile --no-cpu-baseline --no-reference-gpu"
$CMD > $OUT/plain.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -s 10500 -c 3200 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
echo "ncu launch list rc=$?"
python tools/ncu_step_summary.py $OUT/launches.csv $OUT/ncu_summary.json > $OUT/ncu_launch_shares_step.txt 2>&1; head -30 $OUT/ncu_launch_shares_step.txt
du -sh gpurun_out
