#!/bin/bash
# A/B on the GPU box: baseline library (ray_tracer_archive_b200/librtb200_base.so, built from the previous commit) vs the
# current one, same box, interleaved.  Usage: bash tools/ab.sh [workloads...]
W=${@:-"C1:100 C2:200 C3:100 C4:16"}
B=${BASE:-ray_tracer_archive_b200/librtb200_base.so}
for rep in 1 2; do
  [ -f $B ] && RTB200_LIB=$B python tools/quick.py $W 2>&1 | sed 's/^/base /'
  python tools/quick.py $W 2>&1 | sed 's/^/new  /'
done
