#!/bin/bash
mkdir -p gpurun_out/r80
timeout 50 python bench.py --no-profile --no-cpu-baseline --steps 5 > gpurun_out/r80/bench.json 2> gpurun_out/r80/bench.err
echo "bench rc=$?"; cut -c1-240 gpurun_out/r80/bench.json; grep -o '"cuda_graph": [a-z]*\|graph_note[^,]*' gpurun_out/r80/bench.json; tail -2 gpurun_out/r80/bench.err
