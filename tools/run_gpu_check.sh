python tools/bench_dropin.py --steps 1000 2>&1 | tail -20
