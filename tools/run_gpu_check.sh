timeout 600 python -m pytest tests/test_gpu_plr_parity.py tests/test_gpu_plr_loop.py -m gpu -x -q 2>&1 | tail -8
