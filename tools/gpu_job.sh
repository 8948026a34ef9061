mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2_bench_n4_final.json 2> gpurun_out/r2_bench_n4_final.err; echo "bench n4 exit $?"
cut -c1-200 gpurun_out/r2_bench_n4_final.json | tail -1
