#!/bin/bash
# analysis-kernel iteration on one GPU: A/B of build variants, the parity tests, one ncu --set full capture
# Usage: tools/gpu_an.sh <tag> [noprof]
TAG=$1
OUT=gpurun_out
mkdir -p $OUT
PVK_CASES=${PVK_CASES:-metric_10min,cfg2_10min,cfg5_60s} python tools/tune_analyze.py pypevoc_b200/libpvk_r1.so pypevoc_b200/libpvk.so $(ls pypevoc_b200/libpvk_?_*.so 2>/dev/null) 2>&1 | tee $OUT/tune_$TAG.txt
python -W ignore -m pytest tests/test_gpu_parity.py tests/test_gpu_params.py tests/test_gpu_extras.py -x -q -m gpu 2>&1 | tail -4 | tee $OUT/pytest_an_$TAG.log
if [ "$2" != "noprof" ]; then
ncu --kernel-name-base demangled -k "regex:pvk::analyze_kernel" --set full --clock-control none --import-source on -s 2 -c 1 \
    -f -o $OUT/prof_analyze_$TAG python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
fi
