#!/bin/bash
# Run on the GPU box (gpurun): plain run first, then the ncu launch list and one --set full capture of k_extend.
# Usage: bash profiles/run_ncu.sh <tag> [bench args...]
set -u
TAG=$1; shift
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline $*"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_extend -s 4 -c 2 -f -o gpurun_out/${TAG}_extend $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shade_lambert -s 4 -c 1 -f -o gpurun_out/${TAG}_lambert $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
ls -la gpurun_out/ | grep ${TAG}
ncu --set full --clock-control none --import-source on -k regex:k_shade_terminal -s 4 -c 1 -f -o gpurun_out/${TAG}_terminal $CMD > gpurun_out/${TAG}_ncu4.log 2>&1
ls -la gpurun_out/ | grep ${TAG}
