#!/bin/bash
# occupancy / lane matrix on the GPU box: bash tools/matrix.sh "<workloads>" <lib or ''>
W=${1:-"C1:100 C3:100"}
LIB=$2
for so in 1 2 3; do for ln in 1 2 3; do
  if [ -n "$LIB" ]; then RTB200_LIB=$LIB RTB_SHADE_OCC=$so RTB_LANES=$ln python tools/quick.py $W 2>&1 | sed "s|^|$(basename $LIB) |"
  else RTB_SHADE_OCC=$so RTB_LANES=$ln python tools/quick.py $W 2>&1; fi
done; done
