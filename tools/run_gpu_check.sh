for t in 1 0; do
MGPLR_TMA=$t ./tools/kbench 131072 15 256 5 0 1 0 0 | grep -v reset_random
MGPLR_TMA=$t ./tools/kbench 524288 15 128 5 0 1 0 0 | grep -v reset_random
MGPLR_TMA=$t ./tools/kbench 131072 25 256 5 0 0 0 0 | grep -v reset_random
done
MGPLR_TMA=0 timeout 600 python -m pytest tests/test_gpu_env_parity.py -m gpu -x -q 2>&1 | tail -3
