python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 3 --reset-random 1 --no-cpu --envs 131072 > gpurun_out/r1d_b_rr.json 2>gpurun_out/r1d.err; python profiles/summarize_bench.py gpurun_out/r1d_b_rr.json
python bench.py --steps 3 --reset-random 1 --no-cpu > gpurun_out/r1d_b_rr_512k.json 2>>gpurun_out/r1d.err; python profiles/summarize_bench.py gpurun_out/r1d_b_rr_512k.json
tail -2 gpurun_out/r1d.err
