mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
date +%s
timeout 900 $TR bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_bench_n8_final.json 2> gpurun_out/r2_bench_n8_final.err; echo "bench n8 exit $?"
date +%s
cut -c1-250 gpurun_out/r2_bench_n8_final.json | tail -2
HDG_MG_TRACE=1 timeout 300 $TR tools/mg_trace.py 3 4000 2000 > gpurun_out/r2_mgtrace_c4_n8_final.txt 2>&1; echo "trace exit $?"; grep -v "^\[W\|^W1\|^\*\*\*" gpurun_out/r2_mgtrace_c4_n8_final.txt | tail -6
date +%s
