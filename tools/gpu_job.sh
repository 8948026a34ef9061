mkdir -p gpurun_out
python tools/ab_elem.py --orders 3,2,4 --libs head --reps 15
python bench.py --steps 30 --warmup 5 --no-pcg --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('k1 kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'ms_per_step', d['ms_per_step'])"
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
