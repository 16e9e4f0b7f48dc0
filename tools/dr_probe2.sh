#!/bin/bash
N=${1:-131072}
nvcc -O2 -o /tmp/kbench tools/kbench.cu -Ldcd_isaac_b200 -lmgplr -Xlinker -rpath=$PWD/dcd_isaac_b200 2>/dev/null
run() { echo "== $*"; env $1 /tmp/kbench $N 15 $2 8 0 1 $3 $4 2>&1 | grep "us/launch"; }
# amode=1: turn-only actions -> no goals; T=200 < 250: no time-limit resets either: the pure hot path of both variants
run MGPLR_X=0 200 1 0
run MGPLR_X=0 200 1 1
run MGPLR_RR_DYN=0 200 1 1
run MGPLR_RR_CTAS=4 200 1 1
run MGPLR_RR_SPEC=0 200 1 1
