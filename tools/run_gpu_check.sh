timeout 600 python -m pytest tests/test_gpu_env_parity.py -m gpu -x -q 2>&1 | tail -3
MGPLR_RR_PROF=1 timeout 120 ./tools/kbench 131072 15 256 3 0 1 0 1 | tail -5
MGPLR_RR_SPEC=0 timeout 120 ./tools/kbench 131072 15 256 5 0 1 0 1 | grep -v reset_random
timeout 120 ./tools/kbench 131072 25 256 5 0 0 0 1
timeout 120 ./tools/kbench 131072 15 256 5 0 1 0 0 | grep -v reset_random
