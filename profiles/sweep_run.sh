#!/bin/bash
# Runs profiles/merge_profile.py against every scratch/libdbsp_*.so built by sweep_build.py.
#   bash profiles/sweep_run.sh [rows_per_input=50000000] [value_lanes=1]
rows=${1:-50000000}; nv=${2:-1}
mkdir -p gpurun_out
for v in scratch/libdbsp_*.so; do
  n=$(basename "$v" .so)
  DBSP_B200_LIB=$PWD/$v timeout 150 python profiles/merge_profile.py "$rows" "$nv" > "gpurun_out/sweep_${n}_v${nv}.log" 2>&1
  echo "$n nv=$nv rc=$? $(grep -o "'frac': [0-9.]*" "gpurun_out/sweep_${n}_v${nv}.log" | head -1) $(grep -o "rows -> [0-9]*" "gpurun_out/sweep_${n}_v${nv}.log" | head -1)"
done
