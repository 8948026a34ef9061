for v in "" s370 s740 s1480; do
  echo "variant '$v'"
  if [ -n "$v" ]; then export HDG_B200_LIB=$PWD/hdiscontinuousgalerkin.jl_b200/variants/lib_$v.so; fi
  python bench.py --steps 50 --warmup 5 --no-pcg --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  ms_per_step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'])"
done
unset HDG_B200_LIB
python tools/ab_elem.py --orders 2 --reps 10 | grep -v "max |v1"
