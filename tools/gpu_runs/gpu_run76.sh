#!/bin/bash
# 2-GPU data-parallel bench with the device-side trainer glue (graph capture incl. NCCL buckets)
mkdir -p gpurun_out/r76
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-profile > gpurun_out/r76/bench_2gpu.json 2> gpurun_out/r76/bench_2gpu.err
echo "bench 2gpu rc=$?"; cut -c1-600 gpurun_out/r76/bench_2gpu.json; tail -5 gpurun_out/r76/bench_2gpu.err
