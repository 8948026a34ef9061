#!/bin/bash
# One GPU-box session: parity tests, smoke, the headline bench line, ncu evidence of the kernel changed last.
#   gpurun --timeout 900 -- 'bash tools/gpu_round.sh'          (add "sweep" for the single-GPU C5 order sweep)
# Everything lands in gpurun_out/ (copied into profiles/ by hand afterwards).  Most important steps first.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1

echo "== pytest -m gpu"; date +%s
timeout 540 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/pytest_gpu.log

echo "== smoke"; date +%s
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
tail -3 gpurun_out/smoke.log

echo "== bench N=1 (default)"; date +%s
timeout 300 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
cut -c1-400 gpurun_out/bench_n1.json

echo "== ncu: element_quad_kernel<2> (--set full, one launch)"; date +%s
for k in 2; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:element_quad -s 3 -c 1 -f \
      -o gpurun_out/prof_elem_k${k}_quad python bench.py --order $k --steps 1 --warmup 3 --no-pcg --no-cpu \
      > gpurun_out/ncu_k${k}.log 2>&1; echo "ncu k=$k exit $?"
done

echo "== C3 per GPU: k=2, 4 M elements"; date +%s
timeout 240 python bench.py --order 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_k2_4M.json 2> gpurun_out/bench_k2_4M.err; echo "k2 4M exit $?"

if [ "$1" = "sweep" ]; then
echo "== C5: order sweep at ~8M trace dofs, one GPU"; date +%s
for cfg in "1 1155" "2 943" "3 816" "4 730"; do
  set -- $cfg
  timeout 240 python bench.py --order $1 --nx $2 --ny $2 --lx 1 --ly 1 --steps 20 --warmup 3 --no-cpu \
      > gpurun_out/bench_c5_k$1.json 2> gpurun_out/bench_c5_k$1.err; echo "c5 k=$1 exit $?"
done
echo "== ncu launch list of the default bench (shares of the step)"; date +%s
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_k1.csv \
    python bench.py --steps 3 --warmup 3 --maxit 64 --no-cpu > gpurun_out/launches_k1.log 2>&1; echo "launch list exit $?"
fi
date +%s
