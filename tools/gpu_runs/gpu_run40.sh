#!/bin/bash
mkdir -p gpurun_out/r40
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r40/bench2.json 2> gpurun_out/r40/bench2.err
echo "rc=$?"; tail -c 1500 gpurun_out/r40/bench2.json; echo; tail -3 gpurun_out/r40/bench2.err
