#!/bin/bash
N=${1:-131072}
nvcc -O2 -o /tmp/kbench tools/kbench.cu -Ldcd_isaac_b200 -lmgplr -Xlinker -rpath=$PWD/dcd_isaac_b200 2>/dev/null
run() { echo "== $1 amode=$2"; env $1 /tmp/kbench $N 15 200 8 0 1 $2 1 2>&1 | grep "us/launch"; }
run MGPLR_RR_RGRID=0 1
run MGPLR_RR_PDL=0 1
run MGPLR_X=1 1
run MGPLR_RR_RGRID=0 0
run MGPLR_RR_PDL=0 0
run MGPLR_X=1 0
