#!/bin/bash
# Multi-GPU evidence (one box, N GPUs):  gpurun --gpus 8 --timeout 1500 -- 'bash tools/r02_multi_gpu.sh'
O=gpurun_out/r02; mkdir -p $O
NG=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3 | tee $O/pytest_multi_${NG}gpu.txt
for N in $NG $((NG / 2)) ; do
  [ $N -lt 2 ] && continue
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 \
    > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err
  python profiles/summarize_bench.py $O/bench_${N}gpu.json
done
python bench.py --steps 5 --no-cpu --no-python-ref --no-variants > $O/bench_1gpu_samebox.json 2>> $O/bench_${NG}gpu.err
python profiles/summarize_bench.py $O/bench_1gpu_samebox.json
