#!/bin/bash
# A/B on the GPU box: leaf-test parking threshold (RTB_OPT = T << 8) x workloads
W=${1:-"C1:200 C3:200 C4:128"}
for T in 0 6 10 14 18 24; do
  RTB_OPT=$((T << 8)) python tools/quick.py $W 2>&1 | sed "s/^/park=$T /"
done
