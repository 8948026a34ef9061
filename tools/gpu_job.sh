mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
timeout 700 $TR tests/multi_gpu_check.py > gpurun_out/r2_mgcheck_n8_final.log 2>&1; echo "multi_gpu_check exit $?"; grep -i "GPUs:\|limit\|fallback\|parity\|FAIL" gpurun_out/r2_mgcheck_n8_final.log | tail -30
