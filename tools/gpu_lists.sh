#!/bin/bash
# Run on the GPU box: ncu launch lists (gpu__time_duration per kernel) of the auxiliary workloads.
# Usage: tools/gpu_lists.sh <tag> [cfg4 hours] [cfg3 clips]
TAG=${1:-r2}
HOURS=${2:-0.5}
CLIPS=${3:-512}
OUT=gpurun_out
mkdir -p $OUT
for WL in cfg4 cfg3; do
ncu --kernel-name-base demangled -k "regex:pvk::" --metrics gpu__time_duration.sum --clock-control none -c 80 \
    --csv --log-file $OUT/launches_${TAG}_$WL.csv python bench.py --workload $WL --cfg4-hours $HOURS --cfg3-clips $CLIPS \
    --steps 1 --warmup 1 > $OUT/ncu_bench_${TAG}_$WL.log 2>&1
python tools/ncu_launches.py $OUT/launches_${TAG}_$WL.csv | tee $OUT/launches_${TAG}_${WL}_summary.txt
done
