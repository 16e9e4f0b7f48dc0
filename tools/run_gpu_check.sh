timeout 600 python -m pytest tests/test_gpu_env_parity.py -m gpu -x -q -k "host" 2>&1 | tail -3
for d in 0 1; do
MGPLR_HOST_DMA=$d python bench.py --steps 3 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('dma=$d envs 524288 e2e %.3e value %.3e'%(d['e2e']['value'], d['value']))"
MGPLR_HOST_DMA=$d python bench.py --steps 3 --no-cpu --envs 131072 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('dma=$d envs 131072 e2e %.3e value %.3e'%(d['e2e']['value'], d['value']))"
done
python tools/bench_levelops.py --envs 131072 2>/dev/null | tail -25
timeout 300 python -m pytest tests/test_gpu_plr_parity.py tests/test_gpu_storage.py tests/test_gpu_plr_loop.py -m gpu -x -q 2>&1 | tail -3
