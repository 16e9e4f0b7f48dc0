timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
MGPLR_RR_PROF=1 timeout 120 ./tools/kbench 131072 15 256 3 0 1 0 1 | grep -v reset_random
timeout 120 ./tools/kbench 131072 15 256 5 0 1 0 1 | grep -v reset_random
KB_TRACE=1 timeout 120 ./tools/kbench 131072 15 300 1 0 1 0 1 | grep trace | sort -k4 -n -r | head -4
timeout 120 ./tools/kbench 131072 25 256 5 0 0 0 1 | grep -v reset_random
