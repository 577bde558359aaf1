#!/bin/bash
# Run on the GPU box (via gpurun): GPU tests, bench, ncu launch list and full captures.
# Usage: tools/gpu_round.sh <tag> [steps]
TAG=${1:-r1}
STEPS=${2:-5}
OUT=gpurun_out
mkdir -p $OUT
python -W ignore -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee $OUT/pytest_gpu_$TAG.log
python bench.py --steps $STEPS --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -c 3000 $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
if [ "$3" != "noncu" ]; then
ncu --kernel-name-base demangled -k "regex:pvk::" --metrics gpu__time_duration.sum --clock-control none -c 60 \
    --csv --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_bench_$TAG.log 2>&1
tail -3 $OUT/ncu_bench_$TAG.log
if [ "$3" = "list" ]; then exit 0; fi
ncu --kernel-name-base demangled -k "regex:pvk::analyze_kernel" --set full --clock-control none --import-source on -s 2 -c 1 \
    -f -o $OUT/prof_analyze_$TAG python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
ncu --kernel-name-base demangled -k "regex:pvk::resynth_(tile_)?kernel" --set full --clock-control none --import-source on -s 2 -c 1 \
    -f -o $OUT/prof_resynth_$TAG python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
if [ "$3" = "more" ]; then
ncu --kernel-name-base demangled -k "regex:pvk::track_link" --set full --clock-control none --import-source on -s 2 -c 1 \
    -f -o $OUT/prof_link_$TAG python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
ncu --kernel-name-base demangled -k "regex:pvk::resynth_prepare" --set full --clock-control none --import-source on -s 2 -c 1 \
    -f -o $OUT/prof_prepare_$TAG python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
fi
ls -la $OUT
fi
