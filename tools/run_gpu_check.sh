python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r1d_bench_8gpu.json 2> gpurun_out/r1d_bench_8gpu.err
tail -c 600 gpurun_out/r1d_bench_8gpu.json; tail -3 gpurun_out/r1d_bench_8gpu.err
