"""Stage timing of the multigrid V-cycle kernel (HDG_MG_TRACE=1): barrier entry / exit timestamps of the last V-cycle.
   HDG_MG_TRACE=1 python tools/mg_trace.py [order nx ny]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("HDG_MG_TRACE", "1")
import hdg_b200 as hdg  # noqa: E402

order, nx, ny = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (1, 1000, 500)))
qd = {1: 2, 2: 4, 3: 6, 4: 9}[order]
ctx = hdg._Context(order, qd)
lib = ctx.lib
hdg.check(lib.hdg_set_rectangle_mesh(ctx.h, nx, ny, 0.0, 0.0, 2.0, 1.0), ctx.h)
hdg.check(lib.hdg_assemble(ctx.h), ctx.h)
hdg.check(lib.hdg_apply_dirichlet(ctx.h, None), ctx.h)
hdg.check(lib.hdg_set_preconditioner(ctx.h, 2), ctx.h)
info = hdg.api.SolveInfo()
for _ in range(2):
    hdg.check(lib.hdg_solve(ctx.h, 1e-12, 1000, C.byref(info)), ctx.h)
print(f"k={order} {nx}x{ny}: {info.iterations} iterations, {info.solve_ms:.3f} ms, {info.solve_ms / info.iterations * 1e3:.1f} us / iteration")
us = np.zeros(64)
n = lib.hdg_mg_trace(ctx.h, hdg.api.f64p(us), 64)
t = us[:n]
print("timestamps (us):", np.round(t, 1).tolist())
# entries alternate: [start, b1_in, b1_out, b2_in, b2_out, ..., end]
work = [t[1] - t[0]] + [t[i + 1] - t[i] for i in range(2, n - 1, 2)]
bar = [t[i + 1] - t[i] for i in range(1, n - 1, 2)]
print("work per stage (us):", np.round(work, 1).tolist(), "sum", round(float(np.sum(work)), 1))
print("barrier wait of block 0 (us):", np.round(bar, 1).tolist(), "sum", round(float(np.sum(bar)), 1))
print("V-cycle kernel total (us):", round(float(t[-1] - t[0]), 1))
