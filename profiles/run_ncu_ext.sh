#!/bin/bash
# GPU box: one --set full capture of the extend kernel (+ launch list) for a workload at reduced spp.
# Usage: bash profiles/run_ncu_ext.sh <tag> <workload> <spp>
set -u
TAG=$1; WL=$2; SPP=$3
CMD="python tools/quick.py ${WL}:${SPP}"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_extend -s 12 -c 1 -f -o gpurun_out/${TAG}_extend $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
ls -la gpurun_out/ | grep ${TAG}
