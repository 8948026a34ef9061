import os, sys, ctypes as C
sys.path.insert(0, os.getcwd())
import torch, torch.distributed as dist
import hdg_b200 as hdg
rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ctx = hdg._Context(1, 2, 1.0, 1, lr)
ctx.comm_init(dist, device=torch.device("cuda", lr))
us = C.c_double()
for it in (1000, 10000):
    hdg.check(ctx.lib.hdg_comm_pingpong(ctx.h, it, C.byref(us)), ctx.h)
    if rank == 0: print("pingpong", it, "usec/exchange", us.value, flush=True)
dist.barrier(); ctx.close(); dist.destroy_process_group()
