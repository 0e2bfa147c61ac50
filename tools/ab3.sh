#!/bin/bash
run() { env "$@" python tools/quick.py $W 2>&1 | grep Mrays | cut -c1-150; }
W="C1:200 C2:400 C3:200"
run RTB_OPT=0
run RTB_OPT=2
W="C4:128"
run RTB_OPT=3584
run RTB_OPT=3586
