#!/bin/bash
# Steady-state DRAM traffic of k_step_env: six consecutive launches from the middle of a rollout, caches NOT flushed between
# them (ncu --cache-control none) and only single-pass metrics, so nothing is replayed: what the launch really moves when the
# level rows and hot records of the previous launch are still in L2.   gpurun -- 'bash tools/steady_traffic.sh [bench flags]'
O=gpurun_out/r02; mkdir -p $O
ARGS="--steps 1 --warmup 1 --T 64 --no-cpu --no-e2e --no-variants --no-python-ref --no-graph $*"
TAG=$(echo "$*" | tr -c 'a-zA-Z0-9\n' '_')
python bench.py $ARGS > $O/bench_T64$TAG.json 2> $O/t64.err && \
ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum \
  -k regex:k_step_env -s 100 -c 6 --csv --log-file $O/step_env_steady_traffic$TAG.csv python bench.py $ARGS > $O/ncu_steady.log 2>&1
tail -2 $O/ncu_steady.log; grep -c k_step_env $O/step_env_steady_traffic$TAG.csv
