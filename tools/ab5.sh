#!/bin/bash
A=$PWD/ray_tracer_archive_b200/librtb200_A.so
RTB200_LIB=$A ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/r3m_A_c1_launches.csv python tools/quick.py C1:100 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/r3m_H_c1_launches.csv python tools/quick.py C1:100 > /dev/null 2>&1
