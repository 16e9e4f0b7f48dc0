#!/bin/bash
# A/B of NCCL channel counts for the overlapped record all-gather (N = all GPUs of the box)
NG=$(nvidia-smi -L | wc -l)
for ch in default 2 1 4; do
  if [ $ch = default ]; then unset NCCL_MAX_NCHANNELS; else export NCCL_MAX_NCHANNELS=$ch; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $NG --steps 8 --warmup 3 \
     --no-cpu --no-python-ref --no-variants --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('channels=$ch', '%.4e' % d['value'], 'ms/step %.3f' % d['ms_per_step'], 'us/launch %.2f' % d['roofline']['avg_launch_us'])"
done
unset NCCL_MAX_NCHANNELS
python bench.py --steps 8 --no-cpu --no-python-ref --no-variants --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('1 gpu', '%.4e' % d['value'], 'ms/step %.3f' % d['ms_per_step'], 'us/launch %.2f' % d['roofline']['avg_launch_us'])"
