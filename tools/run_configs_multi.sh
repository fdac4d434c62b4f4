#!/bin/bash
# usage: tools/run_configs_multi.sh N [steps] [warmup]   -> gpurun_out/bench_<config>_n<N>.json for the four configurations
N=${1:-8}; K=${2:-3}; W=${3:-3}
mkdir -p gpurun_out
for c in power_scan stiff twothick finegrid; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $N --steps $K --warmup $W --config $c 2> gpurun_out/bench_${c}_n${N}.err | grep '^{' > gpurun_out/bench_${c}_n${N}.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${c}_n${N}.json"))
    print("$c N=$N value %.1f ms/step %.2f kernel_ms %.2f e2e %.1f strong %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d.get("strong", {}).get("value")))
except Exception as e:
    print("$c N=$N FAILED", e)
PY
done
