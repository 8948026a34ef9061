#!/usr/bin/env python
"""A/B of the element kernels (no torch): for each order, runs hdg_assemble with the thread-per-element kernel
(HDG_ELEM_V1=1) and the 4-lanes-per-element kernel in separate child processes, prints ms per pass, elements/s, the
HBM-roofline fraction, and the max relative difference of the assembled values / rhs / [K_e|b_e] between the two.

  python tools/ab_elem.py [--orders 2,3,4] [--nx 2000 --ny 1000] [--reps 20]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ALG_BYTES = {1: 840, 2: 2088, 3: 4200, 4: 7392}
QD = {1: 2, 2: 4, 3: 6, 4: 9}


def child(order, nx, ny, reps, dump):
    import ctypes as C

    import numpy as np

    import hdg_b200 as hdg
    from hdg_b200 import _lib as L
    lib = L.load()
    from hdg_b200.api import _Context
    ctx = _Context(order, QD[order], 1.0, 1)
    h = ctx.h
    L.check(lib.hdg_set_rectangle_mesh(h, nx, ny, 0.0, 0.0, 2.0, 1.0), h)
    ms = C.c_double()
    ts = []
    for _ in range(reps + 3):
        L.check(lib.hdg_assemble(h), h)
        L.check(lib.hdg_last_phase_ms(h, b"element_kernel", C.byref(ms)), h)
        ts.append(ms.value)
    ts = sorted(ts[3:])
    sz = ctx.sizes()
    out = {"order": order, "ncell": int(sz.ncell), "ms_med": ts[len(ts) // 2], "ms_min": ts[0]}
    if dump:
        rhs = np.empty(sz.ndof)
        L.check(lib.hdg_get_rhs(h, L.f64p(rhs)), h)
        nz = np.empty(sz.nnz)
        L.check(lib.hdg_get_values(h, L.f64p(nz)), h)
        Ke = np.empty((sz.m, sz.t))
        be = np.empty(sz.m)
        loc = []
        for cell in (1, 2, 33, int(sz.ncell) // 2, int(sz.ncell)):
            L.check(lib.hdg_get_local(h, cell, L.f64p(Ke), L.f64p(be)), h)
            loc.append(np.concatenate([Ke.ravel(), be]))
        np.savez(dump, rhs=rhs, nz=nz, loc=np.concatenate(loc))
    print("RESULT " + json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--orders", default="2,3,4")
    ap.add_argument("--nx", type=int, default=2000)
    ap.add_argument("--ny", type=int, default=1000)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--child", type=int, default=0)
    ap.add_argument("--dump", default="")
    ap.add_argument("--peak", type=float, default=6454.0)
    ap.add_argument("--libs", default="", help="comma-separated variant names: hdiscontinuousgalerkin.jl_b200/variants/lib_NAME.so")
    a = ap.parse_args()
    if a.child:
        child(a.child, a.nx, a.ny, a.reps, a.dump)
        return
    import numpy as np
    for k in [int(x) for x in a.orders.split(",")]:
        nx, ny = (a.nx, a.ny) if k < 4 else (a.nx // 2, a.ny // 2)
        res = {}
        arms = [("v1", {"HDG_ELEM_V1": "1"}), ("quad", {})]
        for nm in [x for x in a.libs.split(",") if x]:
            arms.append((nm, {"HDG_B200_LIB": os.path.join(ROOT, "hdiscontinuousgalerkin.jl_b200", "variants", f"lib_{nm}.so")}))
        for tag, env in arms:
            e = dict(os.environ)
            e.pop("HDG_ELEM_V1", None)
            e.pop("HDG_B200_LIB", None)
            e.update(env)
            dump = f"/tmp/ab_{k}_{tag}.npz"
            p = subprocess.run([sys.executable, __file__, "--child", str(k), "--nx", str(nx), "--ny", str(ny), "--reps", str(a.reps),
                                "--dump", dump], env=e, capture_output=True, text=True)
            line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            if not line:
                print(f"k={k} {tag}: FAILED rc={p.returncode}\n{p.stdout[-2000:]}\n{p.stderr[-2000:]}")
                continue
            r = json.loads(line[0][7:])
            r["el_per_s"] = r["ncell"] / (r["ms_med"] * 1e-3)
            r["hbm_frac"] = ALG_BYTES[k] * r["el_per_s"] / (a.peak * 1e9)
            res[tag] = r
            print(f"k={k} {tag:5s} {r['ncell']} cells  {r['ms_med']:.3f} ms (min {r['ms_min']:.3f})  {r['el_per_s']:.3e} el/s  HBM frac {r['hbm_frac']:.3f}", flush=True)
        if "v1" in res and "quad" in res:
            A, B = np.load(f"/tmp/ab_{k}_v1.npz"), np.load(f"/tmp/ab_{k}_quad.npz")
            for name in ("rhs", "nz", "loc"):
                d = np.abs(A[name] - B[name]).max() / max(np.abs(A[name]).max(), 1e-300)
                print(f"    max |v1 - quad| / max|v1|  {name}: {d:.3e}  finite={bool(np.isfinite(B[name]).all())}")
            print(f"    speed-up quad vs v1: {res['v1']['ms_med'] / res['quad']['ms_med']:.2f}x", flush=True)
        for nm in [x for x in a.libs.split(",") if x]:          # variant libraries against the default build
            if nm in res and "quad" in res:
                A, B = np.load(f"/tmp/ab_{k}_quad.npz"), np.load(f"/tmp/ab_{k}_{nm}.npz")
                d = max(np.abs(A[name] - B[name]).max() / max(np.abs(A[name]).max(), 1e-300) for name in ("rhs", "nz", "loc"))
                print(f"    {nm}: max rel diff vs default {d:.3e}{' (bit-identical)' if d == 0.0 else ''}; default/{nm} time {res['quad']['ms_med'] / res[nm]['ms_med']:.3f}", flush=True)
        for tag, _ in arms:                                     # the dumps are GBs at 4 M cells
            try:
                os.remove(f"/tmp/ab_{k}_{tag}.npz")
            except OSError:
                pass


if __name__ == "__main__":
    main()
