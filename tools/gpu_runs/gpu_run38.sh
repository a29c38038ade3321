#!/bin/bash
mkdir -p gpurun_out/r38
CMD="python tools/kernel_bench.py misc"
ncu --set full --clock-control none --import-source on -k regex:rope_ -s 6 -c 14 -o /tmp/prof_rope $CMD > gpurun_out/r38/ncu.log 2>&1
echo "ncu rc=$?"
python tools/ncu_extract.py /tmp/prof_rope.ncu-rep lts__t_sector_hit l1tex__t_sector_hit sm__warps_active launch__occupancy achieved_occupancy dram__throughput > gpurun_out/r38/rope_metrics.txt 2>&1
grep -E "^===|time_duration|dram__bytes|warps_active|issue_active|registers|block_size|grid_size|long_scoreboard|lts__t_sector_hit_rate.pct|l1tex__t_sector_hit_rate.pct|dram__throughput.avg.pct" gpurun_out/r38/rope_metrics.txt | head -80
