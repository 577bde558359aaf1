#!/bin/bash
# Run on an 8-GPU box (gpurun --gpus 8): NCCL / peer-memory tests, the metric bench at N = 8 with the
# peer-memory gather and with NCCL, and the strong-scaling workloads (cfg4: 8-hour signal, cfg3: 4096 clips)
# at N = 1, 2, 4, 8.   Usage: tools/gpu_scale8.sh <tag> [cfg4 hours]
TAG=${1:-r2}
HOURS=${2:-8}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo_$TAG.txt 2>&1
timeout 600 python -W ignore -m pytest tests/test_gpu_multi.py -x -q -m gpu -s 2>&1 | grep -v "^NCCL" | tail -6 | tee $OUT/pytest_multi_$TAG.log
run() {  # run <n> <name> <args...>
  local N=$1 NAME=$2; shift 2
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --gpus 1 "$@" > $OUT/bench_${TAG}_${NAME}_g1.json 2> $OUT/bench_${TAG}_${NAME}_g1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 296$N$((RANDOM % 10)) \
      bench.py --gpus $N "$@" > $OUT/bench_${TAG}_${NAME}_g$N.json 2> $OUT/bench_${TAG}_${NAME}_g$N.err
  fi
  echo "$NAME N=$N rc=$? $(python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_${TAG}_${NAME}_g$N.json"))
    print("value %.4g %s  ms/step %.4g  stages %s  e2e %s" % (d["value"], d["unit"], d["ms_per_step"],
          {k: round(v, 3) for k, v in d["stages"].items() if k.endswith("_ms")}, (d.get("e2e") or {}).get("ms_per_step")))
except Exception as e:
    print("no line:", e)
PY
)"
  grep -v "^\[" $OUT/bench_${TAG}_${NAME}_g$N.err | grep -i "error\|fail\|Traceback" | head -3
}
PVK_PEER_GATHER=1 run 8 metric_peer1 --steps 5 --warmup 3
PVK_PEER_GATHER=0 run 8 metric_peer0 --steps 5 --warmup 3 --no-e2e
for N in 8 4 2 1; do
  run $N cfg4 --workload cfg4 --cfg4-hours $HOURS --steps 3 --warmup 2
  run $N cfg3 --workload cfg3 --steps 3 --warmup 2
done
for N in 4 2; do
  run $N metric --steps 5 --warmup 3 --no-e2e
done
