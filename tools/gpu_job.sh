mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r2_gpu.txt 2>&1
echo "== pytest -m gpu"; date +%s
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_final.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/r2_pytest_gpu_final.log
echo "== bench N=1"; date +%s
timeout 600 python bench.py > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err; echo "bench exit $?"
echo "== ncu launch list of the default bench"; date +%s
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2_launches_k1.csv python bench.py --steps 3 --warmup 3 --maxit 64 --no-cpu --no-sweep > gpurun_out/r2_launches_k1.log 2>&1; echo "launch list exit $?"
echo "== ncu --set full: element_schur_kernel<1>"; date +%s
timeout 200 ncu --set full --clock-control none --import-source on -k regex:element_schur -s 3 -c 1 -f -o gpurun_out/r2_prof_elem_k1 python bench.py --steps 1 --warmup 3 --no-pcg --no-cpu --no-e2e --no-sweep > gpurun_out/ncu_k1.log 2>&1; echo "ncu k1 exit $?"
echo "== ncu --set full: one multigrid-PCG iteration (spmv, update_blk, vcycle, dir_mg)"; date +%s
timeout 300 ncu --set full --clock-control none -k regex:"pcg_spmv|pcg_update_blk|mg_vcycle|pcg_dir_mg" -s 8 -c 4 -f -o gpurun_out/r2_prof_mgiter_k1 python tools/mg_trace.py 1 1000 500 > gpurun_out/ncu_mgiter.log 2>&1; echo "ncu mg iter exit $?"
date +%s
