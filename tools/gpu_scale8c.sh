#!/bin/bash
# 8-GPU box: metric bench at N = 8: gather variants x where the main stream joins the gather.
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
PVK_PEER_GATHER=0 run metric_nccl_joinpack --steps 5 --warmup 3 --no-e2e
PVK_PEER_GATHER=1 PVK_PEER_MULTICAST=0 run metric_peer_joinpack --steps 5 --warmup 3 --no-e2e
PVK_PEER_GATHER=1 PVK_PEER_MULTICAST=1 run metric_mcast_joinpack --steps 5 --warmup 3 --no-e2e
PVK_PEER_GATHER=0 PVK_GATHER_JOIN=end run metric_nccl_joinend --steps 5 --warmup 3 --no-e2e
PVK_PEER_GATHER=0 run cfg4_nccl_joinpack --workload cfg4 --steps 3 --warmup 2
PVK_PEER_GATHER=1 PVK_PEER_MULTICAST=1 run cfg4_mcast_joinpack --workload cfg4 --steps 3 --warmup 2
