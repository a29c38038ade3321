#!/bin/bash
# 2-GPU session: NCCL gradient test, then bench.py at N=2 for several caps of the all-reduce CTAs (CALM_DDP_CTAS; 0 = NCCL default)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_ddp_nccl_gpu.py -q -m gpu 2>&1 | tail -4
port=29540
for ctas in ${CTAS_LIST:-0 2 4 8}; do
  port=$((port + 1))
  CALM_DDP_CTAS=$ctas timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus 2 --no-profile > gpurun_out/scale2_ctas$ctas.log 2>&1
  python - "$ctas" <<'PY'
import json, sys
ctas = sys.argv[1]
try:
    line = [l for l in open("gpurun_out/scale2_ctas%s.log" % ctas) if l.startswith("{")][-1]
    d = json.loads(line)
    print("ctas", ctas, "ms/step %.3f" % d["ms_per_step"], "img/s %.0f" % d["value"], "comm", d.get("comm"))
except Exception as e:
    print("ctas", ctas, "FAILED", e)
    print(open("gpurun_out/scale2_ctas%s.log" % ctas).read()[-1500:])
PY
done
