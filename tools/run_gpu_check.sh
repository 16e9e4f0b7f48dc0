set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for p in 0 1; do
MGPLR_PDL=$p ./tools/kbench 131072 15 256 5 0 1 0 0 | grep -v reset_random
MGPLR_PDL=$p ./tools/kbench 1048576 15 64 5 0 1 0 0 | grep -v reset_random
MGPLR_PDL=$p ./tools/kbench 4096 15 256 5 0 1 0 0 | grep -v reset_random
MGPLR_PDL=$p ./tools/kbench 131072 15 256 5 0 1 0 1 | grep -v reset_random
done
python bench.py --steps 5 > gpurun_out/b_r1c.json 2> gpurun_out/b_r1c.err; tail -c 2500 gpurun_out/b_r1c.json
