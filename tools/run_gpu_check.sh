timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python tools/bench_levelops.py --envs 131072 2>/dev/null | grep -E "reset_agent|gae|episode scores" | cut -c1-200
