#!/bin/bash
mkdir -p gpurun_out/r11
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python bench.py --steps 5 --warmup 3 --no-graph --no-cpu-baseline --no-profile > gpurun_out/r11/bench_eager_1gpu.json 2> gpurun_out/r11/bench_eager_1gpu.err
echo "eager 1gpu rc=$?"; head -c 260 gpurun_out/r11/bench_eager_1gpu.json; echo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r11/bench_2gpu.json 2> gpurun_out/r11/bench_2gpu.err
echo "2gpu rc=$?"; head -c 400 gpurun_out/r11/bench_2gpu.json; echo; tail -5 gpurun_out/r11/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r11/ref_2gpu.json 2> gpurun_out/r11/ref_2gpu.err
echo "ref 2gpu rc=$?"; head -c 300 gpurun_out/r11/ref_2gpu.json
