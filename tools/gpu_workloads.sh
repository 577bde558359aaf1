#!/bin/bash
# Run on the GPU box: the auxiliary bench workloads (BASELINE configs[2] clip batch, configs[3] 8-hour signal).
# Usage: tools/gpu_workloads.sh <tag> <ngpus> [cfg4 hours] [cfg3 clips] [steps]
TAG=${1:-r2}
N=${2:-1}
HOURS=${3:-8}
CLIPS=${4:-4096}
STEPS=${5:-3}
OUT=gpurun_out
mkdir -p $OUT
for WL in cfg3 cfg4; do
if [ "$N" = "1" ]; then
timeout 900 python bench.py --workload $WL --cfg4-hours $HOURS --cfg3-clips $CLIPS --steps $STEPS --warmup 2 \
    > $OUT/bench_${TAG}_${WL}_g1.json 2> $OUT/bench_${TAG}_${WL}_g1.err
else
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$N \
    bench.py --gpus $N --workload $WL --cfg4-hours $HOURS --cfg3-clips $CLIPS --steps $STEPS --warmup 2 \
    > $OUT/bench_${TAG}_${WL}_g$N.json 2> $OUT/bench_${TAG}_${WL}_g$N.err
fi
echo "$WL rc=$?"; tail -c 1800 $OUT/bench_${TAG}_${WL}_g$N.json; grep -v "^\[" $OUT/bench_${TAG}_${WL}_g$N.err | tail -8
done
