#!/bin/bash
run() { env "$@" python tools/quick.py $W 2>&1 | grep Mrays | cut -c1-165; }
W="C1:300 C3:200"
run RTB_MAX_LEAF=1
run RTB_MAX_LEAF=2
run RTB_MAX_LEAF=3
run RTB_OPEN_MIN_REL=0.0625
run RTB_OPEN_MIN_REL=0.25
W="C4:128"
run RTB_MAX_LEAF=1
run RTB_MAX_LEAF=3
