#!/bin/bash
# 8-GPU box: high-priority stitch stream, gather joined at the end / before the rendering.
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
run() {
  local NAME=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 297$((RANDOM % 90 + 10)) \
      bench.py --gpus 8 "$@" > $OUT/bench_${TAG}_${NAME}_g8.json 2> $OUT/bench_${TAG}_${NAME}_g8.err
  echo "$NAME rc=$? $(python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_${TAG}_${NAME}_g8.json"))
    print("value %.4g  ms/step %.4g  stages %s  e2e %s %s" % (d["value"], d["ms_per_step"],
          {k: round(v, 3) for k, v in d["stages"].items() if k.endswith("_ms")}, (d.get("e2e") or {}).get("ms_per_step"), d.get("selfcheck", {}).get("nvswitch_multicast")))
except Exception as e:
    print("no line:", e)
PY
)"
}
PVK_GATHER_JOIN=end run metric_prio_joinend --steps 5 --warmup 3 --no-e2e
PVK_GATHER_JOIN=pack run metric_prio_joinpack --steps 5 --warmup 3 --no-e2e
PVK_GATHER_JOIN=end PVK_PEER_GATHER=0 run metric_prio_joinend_nccl --steps 5 --warmup 3 --no-e2e
