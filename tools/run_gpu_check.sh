MGPLR_RR_PROF=1 timeout 120 ./tools/kbench 131072 15 256 3 0 1 0 1 | grep -v reset_random | tail -5
timeout 600 python -m pytest tests/test_gpu_env_parity.py -m gpu -x -q -k "random or spec or oracle_large or rollout" 2>&1 | tail -3
