mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/multi_gpu_check_n2.log 2>&1; echo "multi_gpu_check exit $?"; tail -12 gpurun_out/multi_gpu_check_n2.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"; tail -c 1500 gpurun_out/bench_n2.json
