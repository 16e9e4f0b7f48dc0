#!/bin/bash
# What the round-end driver runs on a fresh B200 box, in one go:  gpurun --timeout 1500 -- 'bash tools/run_gpu_check.sh'
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/check_bench_reference.json 2> gpurun_out/check_bench.err
python bench.py --steps 5 > gpurun_out/check_bench_main.json 2>> gpurun_out/check_bench.err
python profiles/summarize_bench.py gpurun_out/check_bench_reference.json gpurun_out/check_bench_main.json
tail -3 gpurun_out/check_bench.err
timeout 120 python tools/bench_dropin.py > gpurun_out/check_dropin.json 2>> gpurun_out/check_bench.err; grep -c step_env gpurun_out/check_dropin.json
