mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR tests/multi_gpu_check.py > gpurun_out/r2_mgcheck_n2c.log 2>&1; echo "multi_gpu_check exit $?"; grep -i "limit\|fallback\|parity\|FAIL\|Error" gpurun_out/r2_mgcheck_n2c.log | tail -12
