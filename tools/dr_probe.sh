#!/bin/bash
# DR step-kernel probe on the GPU box: builds tools/kbench into /tmp and runs the A/B knobs.
#   gpurun -- 'bash tools/dr_probe.sh [N]'
N=${1:-131072}
nvcc -O2 -o /tmp/kbench tools/kbench.cu -Ldcd_isaac_b200 -lmgplr -Xlinker -rpath=$PWD/dcd_isaac_b200 2>/dev/null
run() { echo "== $*"; env "$@" /tmp/kbench $N 15 256 8 0 1 0 1 2>&1 | grep "us/launch"; }
run MGPLR_X=0
run MGPLR_RR_DYN=0
run MGPLR_RR_RGRID=1
run MGPLR_RR_RGRID=4
run MGPLR_RR_CTAS=4 MGPLR_RR_RGRID=1
run MGPLR_RR_CTAS=2 MGPLR_RR_RGRID=4
run MGPLR_RR_SPEC=0
echo "== prof"; MGPLR_RR_PROF=1 /tmp/kbench $N 15 256 2 0 1 0 1 2>&1 | grep -A3 "prof step [45]" | head -10
