mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR tests/multi_gpu_check.py > gpurun_out/r2_mgcheck_n2b.log 2>&1; echo "multi_gpu_check exit $?"; grep -i "set_mesh\|parity\|FAIL" gpurun_out/r2_mgcheck_n2b.log | tail -12
timeout 400 $TR tools/partition_bench.py 2000 1000 --shuffle-blocks 4096 > gpurun_out/r2_partition_shuffled_n2.json 2> gpurun_out/r2_partition_shuffled_n2.err; echo "shuffled exit $?"; tail -1 gpurun_out/r2_partition_shuffled_n2.json
timeout 400 $TR tools/partition_bench.py 2000 1000 --shuffle-blocks 4096 --morton > gpurun_out/r2_partition_morton_n2.json 2> gpurun_out/r2_partition_morton_n2.err; echo "morton exit $?"; tail -1 gpurun_out/r2_partition_morton_n2.json
