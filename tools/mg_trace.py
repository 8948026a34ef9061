"""Stage timing of the multigrid V-cycle kernel (HDG_MG_TRACE=1): barrier entry / exit timestamps of the last V-cycle.
   HDG_MG_TRACE=1 python tools/mg_trace.py [order nx ny]            (one GPU)
   HDG_MG_TRACE=1 python -m torch.distributed.run --nproc-per-node N ... tools/mg_trace.py order nx ny    (ny = global rows)"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("HDG_MG_TRACE", "1")
import hdg_b200 as hdg  # noqa: E402

world, rank, lr = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
order, nx, ny = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (1, 1000, 500)))
qd = {1: 2, 2: 4, 3: 6, 4: 9}[order]
dist = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ctx = hdg._Context(order, qd, 1.0, 1, lr)
if dist is not None:
    ctx.comm_init(dist, device=torch.device("cuda", lr))
lib = ctx.lib
hdg.check(lib.hdg_set_rectangle_mesh(ctx.h, nx, ny, 0.0, 0.0, 2.0, 1.0), ctx.h)
hdg.check(lib.hdg_assemble(ctx.h), ctx.h)
hdg.check(lib.hdg_apply_dirichlet(ctx.h, None), ctx.h)
hdg.check(lib.hdg_set_preconditioner(ctx.h, 2), ctx.h)
info = hdg.api.SolveInfo()
for _ in range(2):
    hdg.check(lib.hdg_solve(ctx.h, 1e-12, 1000, C.byref(info)), ctx.h)
full_ms, its = info.solve_ms, info.iterations
ph = {}
for name in ("mg_setup", "solve_loop"):
    v = C.c_double()
    lib.hdg_last_phase_ms(ctx.h, name.encode(), C.byref(v))
    ph[name] = v.value
lib.hdg_solve(ctx.h, 1e-12, 4, C.byref(info))        # set-up + one graph chunk of 4 iterations (status 7: not converged)
short_ms = info.solve_ms
per_it = (full_ms - short_ms) / max(its - 4, 1)
us = np.zeros(64)
n = lib.hdg_mg_trace(ctx.h, hdg.api.f64p(us), 64)
t = us[:n]
if rank == 0:
    print(f"k={order} {nx}x{ny} on {world} GPU(s): {its} iterations, {full_ms:.3f} ms; 4 iterations + set-up {short_ms:.3f} ms -> "
          f"{per_it * 1e3:.1f} us / iteration, set-up ~{short_ms - 4 * per_it:.3f} ms")
    print(f"phases of the full solve: mg_setup {ph['mg_setup']:.3f} ms, iteration loop {ph['solve_loop']:.3f} ms "
          f"({ph['solve_loop'] / its * 1e3:.1f} us / iteration incl. graph capture and the host round trips), rest {full_ms - ph['mg_setup'] - ph['solve_loop']:.3f} ms")
    print("timestamps (us):", np.round(t, 1).tolist())
    # entries alternate: [start, b1_in, b1_out, b2_in, b2_out, ..., end]
    work = [t[1] - t[0]] + [t[i + 1] - t[i] for i in range(2, n - 1, 2)]
    bar = [t[i + 1] - t[i] for i in range(1, n - 1, 2)]
    print("work per stage (us):", np.round(work, 1).tolist(), "sum", round(float(np.sum(work)), 1))
    print("barrier wait of block 0 (us):", np.round(bar, 1).tolist(), "sum", round(float(np.sum(bar)), 1))
    print("V-cycle kernel total (us):", round(float(t[-1] - t[0]), 1))
ctx.close()
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
