#!/bin/bash
# Round-2 evidence run on one B200:  gpurun --timeout 2400 -- 'bash tools/r02_gpu_session.sh'
# Everything lands in gpurun_out/r02/ ; the files quoted in profiles/README.md are copied to profiles/r02_* afterwards.
O=gpurun_out/r02; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/pytest_gpu.txt
python __graft_entry__.py smoke 2>&1 | tail -1 | tee $O/smoke.txt
tools/bw_probe 4096 | tee $O/bw_probe.txt
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench.err
python bench.py --steps 5 > $O/bench_main.json 2>> $O/bench.err
python profiles/summarize_bench.py $O/bench_reference.json $O/bench_main.json | tee $O/bench_summary.txt
tail -3 $O/bench.err
# ncu passes only after the same command exited 0 without ncu
python bench.py --steps 2 --warmup 1 --T 32 --no-cpu --no-e2e --no-variants --no-python-ref > $O/bench_T32.json 2>> $O/bench.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench_T32.csv \
  python bench.py --steps 2 --warmup 1 --T 32 --no-cpu --no-e2e --no-variants --no-python-ref > $O/ncu_launches.log 2>&1
python bench.py --steps 1 --warmup 1 --T 8 --no-cpu --no-e2e --no-variants --no-python-ref --no-graph > $O/bench_T8.json 2>> $O/bench.err && \
ncu --set full --clock-control none --import-source on -k regex:k_step_env -s 10 -c 1 -o $O/step_env_full \
  python bench.py --steps 1 --warmup 1 --T 8 --no-cpu --no-e2e --no-variants --no-python-ref --no-graph > $O/ncu_full.log 2>&1
ncu -i $O/step_env_full.ncu-rep --page raw --csv > $O/step_env_full_raw.csv 2>/dev/null
timeout 300 python tools/bench_dropin.py > $O/dropin.json 2>> $O/bench.err
timeout 600 python tools/bench_configs.py > $O/configs.json 2>> $O/bench.err
tail -3 $O/bench.err
ls -la $O
