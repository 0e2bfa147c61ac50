#!/bin/bash
A=$PWD/ray_tracer_archive_b200/librtb200_A.so
for rep in 1 2; do
  RTB200_LIB=$A python tools/quick.py C1:500 C3:500 C4:256 2>&1 | grep Mrays | cut -c1-120 | sed 's/^/A /'
  python tools/quick.py C1:500 C3:500 C4:256 2>&1 | grep Mrays | cut -c1-120 | sed 's/^/HEAD /'
done
