#!/bin/bash
# Turn the artefacts of tools/gpu_round.sh <tag> (in gpurun_out/) into the tracked summaries under profiles/.
# Usage: tools/profiles_from_round.sh <tag> [analysis region spec ...]   (region spec: see tools/ncu_regions.py)
T=$1; shift
cp gpurun_out/bench_$T.json profiles/${T}_bench.json
cp gpurun_out/launches_$T.csv profiles/${T}_launches.csv
cp gpurun_out/pytest_gpu_$T.log profiles/${T}_pytest_gpu.log
python tools/ncu_launches.py gpurun_out/launches_$T.csv > profiles/${T}_launches_summary.txt
for k in analyze resynth link prepare; do
  [ -f gpurun_out/prof_${k}_$T.ncu-rep ] || continue
  python tools/ncu_raw.py gpurun_out/prof_${k}_$T.ncu-rep > profiles/${T}_${k}_ncu_raw.txt
  ncu -i gpurun_out/prof_${k}_$T.ncu-rep --page source --csv --print-source cuda,sass > /tmp/${k}_src_$T.csv 2>/dev/null
  python tools/ncu_lines.py /tmp/${k}_src_$T.csv 40 > profiles/${T}_${k}_hot_lines.txt
done
if [ $# -gt 0 ]; then python tools/ncu_regions.py /tmp/analyze_src_$T.csv "$@" > profiles/${T}_analyze_regions.txt; fi
