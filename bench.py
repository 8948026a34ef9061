#!/usr/bin/env python
"""bench.py - headline benchmark of the HDG hot path (BASELINE.json metric:
"HDG elements/sec (assemble+condense+scatter) and trace PCG solve time vs k").

headline workload (N GPUs): BASELINE config C2 per GPU - Poisson HDG k=1 on a 1M-element structured triangle mesh
         (rectangle_mesh 1000x500 on [0,2]x[0,1] per GPU, quad_degree 2, tau 1, f = 2 pi^2 sin sin); under torchrun every
         rank owns a strip of ONE global mesh 1000 x (500 N) (weak scaling, no data-path collective in assembly).
A "step" = one full pass of hdg_assemble over the mesh (memset + fused element kernel: local blocks, static condensation,
scatter into the block trace matrix + rhs, K_e/b_e stored).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (C port of the Julia driver, all host
                                                           # threads, the same 1M-element mesh per step)

The driver keeps only SCALAR entries of the top-level objects (roofline, e2e, cpu_baseline, config), so everything the metric
names travels as flat keys there:
  roofline.k{2,3,4}_*      order sweep on one GPU: elements/s, kernel ms, HBM fraction, multigrid-PCG solve (C3 4M k=2, 4M k=3, 1M k=4)
  roofline.pcg_* / mg_*    trace solve of the headline workload: Jacobi-PCG (iterations, ms/iteration, roofline against the CSR
                           byte model AND against the ncu-measured traffic), multigrid-PCG; recover_*, err2
  roofline.c3_* / c4_* / c5_*  (N > 1) strong scaling of the BASELINE configs that are defined on several GPUs: C3 on 2/4, C4 on 8,
                           the C5 sweep on every N - assembly elements/s and multigrid-PCG solve time of the communicating phase
  config.parity_*          (N > 1) N-GPU solution against the 1-GPU solution of the same global problem (Jacobi and multigrid)
  cpu_baseline.*           the C port per phase: assembly on 1 core and on all cores, PCG per iteration, recovery, sparse direct solve
The full nested records stay under "detail" (visible in the stdout tail only).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

QD_FOR_ORDER = {1: 2, 2: 4, 3: 6, 4: 9}
# algorithmic figures per element, SURVEY.md section 8(d) / BASELINE.md section 3
ALG_BYTES = {1: 840, 2: 2088, 3: 4200, 4: 7392}
ALG_FLOPS = {1: 3090, 2: 18648, 3: 73896, 4: 268350}
RECOVER_BYTES = {1: 600, 2: 1656, 3: 3456, 4: 6240}
# the meshes the configs are DEFINED on (total, all GPUs)
C2_MESH = (1, 1000, 500)         # k, nx, ny on [0,2]x[0,1]
C3_MESH = (2, 2000, 1000)
C4_MESH = (3, 4000, 2000)
C5_MESH = {1: 1155, 2: 943, 3: 816, 4: 730}      # square meshes on [0,1]^2, ~8M trace dofs


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.idx = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------------
# CPU side: the C port of the reference's path (oracle/, permitted here only as the timed baseline)
# ------------------------------------------------------------------------------------------------------------------------
_CPU_CACHE = {}


def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import hdg_oracle as orc      # checker / CPU baseline only
    import hdg_oracle_c as occ
    return orc, occ


def cpu_mesh_tables(order, qd, nx, ny):
    orc, occ = _oracle()
    key = (order, qd, nx, ny)
    if key not in _CPU_CACHE:
        _CPU_CACHE[key] = (occ.rectangle_mesh(nx, ny, (0.0, 0.0), (2.0, 1.0)), orc.build_tables(order, qd))
    return _CPU_CACHE[key]


def cpu_port_elements_per_s(order, qd, sample, nthreads, passes=1, keep=False):
    """Time the C port of the reference's doassemble on a mesh (host cores).  Only the C call is timed; the mesh (C restatement
    of rectangle_mesh) and the tables are built once and cached."""
    orc, occ = _oracle()
    mesh, tab = cpu_mesh_tables(order, qd, *sample)
    if "warm" not in _CPU_CACHE:
        occ.doassemble(occ.rectangle_mesh(8, 4), tab, nthreads=nthreads, keep_local=True)   # warm the library / thread pool
        _CPU_CACHE["warm"] = True
    out = None
    t0 = time.perf_counter()
    for _ in range(passes):
        out = occ.doassemble(mesh, tab, nthreads=nthreads, keep_local=True)
    dt = (time.perf_counter() - t0) / passes
    return mesh.ncells / dt, dt, mesh.ncells, (out if keep else None)


def host_threads():
    """All host threads this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers, which would turn the
    "all cores" arm into a single-threaded one, so the CPU affinity mask decides, not the OpenMP default (the count is
    handed to the C port explicitly: `num_threads` clauses)."""
    try:
        return max(len(os.sched_getaffinity(0)), 1)
    except Exception:
        return max(os.cpu_count() or 1, 1)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  Julia cannot run here, so this is
    the loop-faithful C port (oracle/hdg_oracle.c) with all host threads, on the SAME mesh as the GPU arm's step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    order = args.order
    qd = args.quad_degree or QD_FOR_ORDER[order]
    cores = host_threads()
    full = {1: (1000, 500), 2: (2000, 1000), 3: (2000, 1000), 4: (1000, 500)}[order]
    sample = (args.nx, args.ny) if args.nx and args.ny else (full if order == 1 else ((500, 250) if order == 2 else (250, 125)))
    for _ in range(args.warmup):
        cpu_port_elements_per_s(order, qd, (40, 20), cores)
    t_tot, n_tot = 0.0, 0
    for _ in range(args.steps):
        eps, dt, ncell, _ = cpu_port_elements_per_s(order, qd, sample, cores)
        t_tot += dt
        n_tot += ncell
    value = n_tot / t_tot
    same = sample == full
    desc = (f"rectangle_mesh {sample[0]}x{sample[1]} ({2*sample[0]*sample[1]} elements) per step"
            + (" = the GPU arm's mesh" if same else " (bounded sample of the GPU arm's mesh)") + ", same k/quad_degree/tau/f")
    out = {
        "impl": "reference", "metric": "HDG elements/sec (assemble+condense+scatter) and trace PCG solve time vs k", "value": value,
        "unit": "elements/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_tot / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C2: Poisson HDG k={order} quad_degree={qd} tau=1, {desc}", "elements_per_step": 2 * sample[0] * sample[1]},
        "cpu_baseline": {"value": value, "unit": "elements/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def cpu_phases(order, qd, budget_s=45.0):
    """C port per phase on this box's host cores (phases timed separately): doassemble on 1 core (bounded sample) and
    on all cores (the full C2 mesh), Jacobi-PCG per iteration on the full trace system (1 core / all cores, a fixed number of
    iterations), get_u_sigma!, and a sparse direct solve (scipy SuperLU in place of the reference's UMFPACK) on the sample."""
    import scipy.sparse.linalg as spla
    orc, occ = _oracle()
    cores = host_threads()
    sample = (400, 200) if order == 1 else ((200, 100) if order == 2 else (100, 50))
    full = (1000, 500) if order == 1 else sample
    v1, dt1, n1, asm1 = cpu_port_elements_per_s(order, qd, sample, 1, passes=2 if order == 1 else 1, keep=True)
    vN, dtN, nN, asmN = cpu_port_elements_per_s(order, qd, full, cores, passes=3, keep=True)
    out = {"value": v1, "unit": "elements/s", "cores": 1, "kind": "port",
           "sample": f"C port of doassemble (oracle/hdg_oracle.c), rectangle_mesh {sample[0]}x{sample[1]} = {n1} elements, "
                     f"{dt1:.2f} s/pass on one thread (the reference is single-threaded)",
           "all_cores_value": vN, "all_cores": cores, "all_cores_s_per_pass": dtN, "all_cores_elements": nN}
    try:
        mesh, tab = cpu_mesh_tables(order, qd, *full)
        K, rhs, Ke, be = asmN
        nt = tab.nt
        bf = mesh.boundary_faces_sorted()
        dofs = (nt * (bf[:, None] - 1) + np.arange(1, nt + 1)[None, :]).ravel()
        K2, f2, md, dset = occ.apply_dirichlet_homogeneous(K, rhs, dofs)
        nit = 12
        t0 = time.perf_counter(); x, it, rel = occ.pcg(K2, f2, dset, rtol=1e-30, maxit=nit, nthreads=1); tp1 = (time.perf_counter() - t0) / max(it, 1)
        t0 = time.perf_counter(); x, it, rel = occ.pcg(K2, f2, dset, rtol=1e-30, maxit=4 * nit, nthreads=cores); tpN = (time.perf_counter() - t0) / max(it, 1)
        out.update({"pcg_ms_per_iter_1core": 1e3 * tp1, "pcg_ms_per_iter_all_cores": 1e3 * tpN, "pcg_dofs": int(K2.shape[0])})
        t0 = time.perf_counter(); occ.recover(mesh, tab, x, Ke, be, nthreads=1); tr1 = time.perf_counter() - t0
        t0 = time.perf_counter(); occ.recover(mesh, tab, x, Ke, be, nthreads=cores); trN = time.perf_counter() - t0
        out.update({"recover_ms_1core": 1e3 * tr1, "recover_ms_all_cores": 1e3 * trN, "recover_elements": int(mesh.ncells)})
        # direct solve like the reference's K \ b, on the bounded sample (<= 0.5 M dofs)
        meshs, tabs = cpu_mesh_tables(order, qd, *sample)
        Ks, rhss, _, _ = asm1
        bfs = meshs.boundary_faces_sorted()
        dofss = (nt * (bfs[:, None] - 1) + np.arange(1, nt + 1)[None, :]).ravel()
        K2s, f2s, _, dsets = occ.apply_dirichlet_homogeneous(Ks, rhss, dofss)
        t0 = time.perf_counter(); xs = spla.spsolve(K2s.tocsc(), f2s); tsp = time.perf_counter() - t0
        t0 = time.perf_counter(); xp, itp, relp = occ.pcg(K2s, f2s, dsets, rtol=1e-12, maxit=200000, nthreads=cores); tpc = time.perf_counter() - t0
        out.update({"spsolve_s": tsp, "spsolve_dofs": int(K2s.shape[0]), "pcg_full_solve_s_all_cores": tpc, "pcg_full_solve_iters": int(itp),
                    "pcg_vs_spsolve_maxrel": float(np.abs(xp - xs).max() / np.abs(xs).max())})
    except Exception as ex:      # the headline baseline above stays
        out["phases_error"] = str(ex)[:160]
    return out


# ------------------------------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------------------------------
class Env:
    """torch / distributed plumbing shared by the legs."""

    def __init__(self):
        import torch
        import hdg_b200 as hdg
        self.torch, self.hdg = torch, hdg
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        self.peak, self.peak_src = measured_peaks()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def maxf(self, v):
        if self.dist is None:
            return float(v)
        t = self.torch.tensor([float(v)], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sumf(self, v):
        if self.dist is None:
            return float(v)
        t = self.torch.tensor([float(v)], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def context(self, order, qd, local_solver=0, comm=True):
        ctx = self.hdg._Context(order, qd, 1.0, 1, self.local_rank, local_solver)
        if comm and self.world > 1:
            ctx.comm_init(self.dist, device=self.dev)
        return ctx


def phase_ms(hdg, ctx, name):
    v = C.c_double()
    hdg.check(ctx.lib.hdg_last_phase_ms(ctx.h, name.encode(), C.byref(v)), ctx.h)
    return v.value


def solve_leg(env, ctx, precond, rtol, maxit, repeat=1):
    """hdg_solve with the given preconditioner on the applied system; best device time of `repeat` solves."""
    hdg, lib = env.hdg, ctx.lib
    hdg.check(lib.hdg_set_preconditioner(ctx.h, precond), ctx.h)
    info = hdg.api.SolveInfo()
    best = None
    for _ in range(repeat):
        st = lib.hdg_solve(ctx.h, rtol, maxit, C.byref(info))
        if st not in (0, 7):
            hdg.check(st, ctx.h)
        ms = env.maxf(info.solve_ms)
        best = ms if best is None else min(best, ms)
    return {"iterations": int(info.iterations), "converged": bool(info.converged), "relres": float(info.relres),
            "solve_ms": best, "ms_per_iter": best / max(info.iterations, 1)}


def workload(env, order, qd, nx, ny_total, lx, ly, steps, warmup, solves=(), rtol=1e-12, maxit=40000, perturb=0.0, local_solver=0,
             sampler=None, recover=False):
    """One workload on all ranks: mesh (strips of quad rows when world > 1), `steps` timed hdg_assemble passes, the element
    kernel alone, then the requested solves ("jacobi", "block", "mg") on the assembled system, recovery and err2."""
    hdg, torch = env.hdg, env.torch
    ctx = env.context(order, qd, local_solver)
    lib = ctx.lib
    hdg.check(lib.hdg_set_rectangle_mesh(ctx.h, nx, ny_total, 0.0, 0.0, lx, ly), ctx.h)
    if perturb > 0:
        hdg.check(lib.hdg_perturb_nodes(ctx.h, perturb, 12345), ctx.h)
    s = ctx.sizes()
    ncell = int(s.ncell)
    ncell_total = 2 * nx * ny_total
    stream = torch.cuda.ExternalStream(int(lib.hdg_stream(ctx.h)), device=env.dev)
    for _ in range(max(warmup, 3)):
        hdg.check(lib.hdg_assemble_async(ctx.h), ctx.h)
    hdg.check(lib.hdg_sync(ctx.h), ctx.h)
    hdg.check(lib.hdg_assemble(ctx.h), ctx.h)        # checked variant once: raises on bad geometry / singular cells
    launches0 = lib.hdg_launch_count(ctx.h)
    env.barrier()
    if sampler is not None:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(steps):
            hdg.check(lib.hdg_assemble_async(ctx.h), ctx.h)
        e1.record(stream)
    env.barrier()
    clocks = sampler.stop() if sampler is not None else None
    ms_step = env.maxf(e0.elapsed_time(e1)) / steps
    launches = int(lib.hdg_launch_count(ctx.h) - launches0)
    kern = []
    for _ in range(min(steps, 20)):
        hdg.check(lib.hdg_assemble_async(ctx.h), ctx.h)
        hdg.check(lib.hdg_sync(ctx.h), ctx.h)
        kern.append(phase_ms(hdg, ctx, "element_kernel"))
    kern_ms = float(np.mean(kern))
    out = {"order": order, "quad_degree": qd, "nx": nx, "ny": ny_total, "elements": ncell_total, "elements_per_gpu": ncell,
           "ndof_per_gpu": int(s.ndof), "nnz_per_gpu": int(s.nnz), "ms_per_step": ms_step, "elements_per_s": ncell_total / (ms_step * 1e-3),
           "kernel_ms": kern_ms, "kernel_GBps": ALG_BYTES[order] * ncell / (kern_ms * 1e-3) / 1e9,
           "frac": ALG_BYTES[order] * ncell / (kern_ms * 1e-3) / 1e9 / env.peak, "launches": launches, "clocks": clocks}
    if solves:
        hdg.check(lib.hdg_apply_dirichlet(ctx.h, None), ctx.h)
        for name in solves:
            pid = {"jacobi": 0, "block": 1, "mg": 2}[name]
            try:
                out[name] = solve_leg(env, ctx, pid, rtol, maxit, repeat=2 if name == "mg" else 1)
            except Exception as ex:
                out[name] = {"error": str(ex)[:200]}
        hdg.check(lib.hdg_set_preconditioner(ctx.h, 0), ctx.h)
        if recover:
            rec = []
            for _ in range(4):
                hdg.check(lib.hdg_recover(ctx.h), ctx.h)
                rec.append(phase_ms(hdg, ctx, "recover"))
            e2 = C.c_double()
            hdg.check(lib.hdg_errornorm(ctx.h, 1, C.byref(e2)), ctx.h)
            out["recover_ms"] = env.maxf(float(np.mean(rec[1:])))
            out["recover_frac"] = RECOVER_BYTES[order] * ncell / (out["recover_ms"] * 1e-3) / 1e9 / env.peak
            out["err2"] = e2.value
    return out, ctx


def parity_leg(env):
    """N-GPU solution of a small global problem against the 1-GPU solution of the same problem (every rank solves the whole
    problem once more on its own GPU without the communicator and compares its owned trace rows): Jacobi-PCG and multigrid-PCG,
    k = 1 and k = 3.  Returns the worst relative difference, whether the iteration counts agree, and err2 agreement."""
    hdg, world = env.hdg, env.world
    worst, iters_ok, err_ok, cases = 0.0, True, True, []
    for order, qd, precond in ((1, 2, 0), (1, 2, 2), (3, 6, 0), (3, 6, 2)):
        nx, ny = 256, 128 * world      # level 0 of the vertex hierarchy is distributed, the rest replicated
        res = []
        for multi in (True, False):
            ctx = env.context(order, qd, comm=multi)
            lib = ctx.lib
            hdg.check(lib.hdg_set_rectangle_mesh(ctx.h, nx, ny, 0.0, 0.0, 2.0, float(world)), ctx.h)
            hdg.check(lib.hdg_assemble(ctx.h), ctx.h)
            hdg.check(lib.hdg_apply_dirichlet(ctx.h, None), ctx.h)
            hdg.check(lib.hdg_set_preconditioner(ctx.h, precond), ctx.h)
            info = hdg.api.SolveInfo()
            hdg.check(lib.hdg_solve(ctx.h, 1e-13, 100000, C.byref(info)), ctx.h)
            hdg.check(lib.hdg_recover(ctx.h), ctx.h)
            e2 = C.c_double()
            hdg.check(lib.hdg_errornorm(ctx.h, 1, C.byref(e2)), ctx.h)
            s = ctx.sizes()
            x = np.empty(s.ndof)
            hdg.check(lib.hdg_get_trace(ctx.h, hdg.api.f64p(x)), ctx.h)
            part = ctx.partition()
            res.append((x, info.iterations, e2.value, part))
            ctx.close()
        (xm, itm, em, pm), (x1, it1, e1, _) = res
        nt = order + 1
        own = x1[pm["face_begin"] * nt: pm["face_end"] * nt]
        rel = env.maxf(np.abs(xm - own).max()) / env.maxf(np.abs(x1).max())
        worst = max(worst, rel)
        iters_ok = iters_ok and abs(itm - it1) <= max(1, it1 // 200)      # the dot products are summed in another order: +-0.5 % of a few thousand Jacobi iterations
        err_ok = err_ok and abs(em - e1) <= 1e-9 * abs(e1) + 1e-20
        cases.append({"order": order, "precond": precond, "max_rel": rel, "iters": [itm, it1], "err2": [em, e1]})
    return {"max_rel": worst, "iters_equal": bool(iters_ok), "err2_equal": bool(err_ok), "cases": cases,
            "what": f"256 x {128*world} mesh on {world} GPUs vs the same mesh on one GPU; Jacobi and multigrid PCG, k=1 and k=3, rtol 1e-13"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--order", type=int, default=1)
    ap.add_argument("--quad-degree", type=int, default=0)
    ap.add_argument("--nx", type=int, default=0)
    ap.add_argument("--ny", type=int, default=0)
    ap.add_argument("--perturb", type=float, default=0.0, help="jitter interior nodes by this fraction of h")
    ap.add_argument("--rtol", type=float, default=1e-12)
    ap.add_argument("--maxit", type=int, default=40000)
    ap.add_argument("--no-pcg", action="store_true")
    ap.add_argument("--pcg", default="all", choices=["all", "mg"], help="mg: time only the multigrid-preconditioned solve (large problems)")
    ap.add_argument("--ly", type=float, default=0.0, help="height of the global domain [0,lx]x[0,ly] (default: number of GPUs)")
    ap.add_argument("--lx", type=float, default=2.0)
    ap.add_argument("--strong", action="store_true", help="--nx/--ny name the GLOBAL mesh, split over the ranks (default: per-GPU mesh, weak scaling)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the order sweep / strong-scaling configs / multi-GPU parity leg")
    ap.add_argument("--local-solver", type=int, default=0, help="1: literal quadrature + dense LU element kernel")
    ap.add_argument("--e2e-faces", action="store_true", help="also upload mesh.faces in the e2e step (it is rebuilt on the device otherwise)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)

    env = Env()
    hdg, torch, dist = env.hdg, env.torch, env.dist
    world, rank, local_rank = env.world, env.rank, env.local_rank
    order = args.order
    qd = args.quad_degree or QD_FOR_ORDER[order]
    default_mesh = {1: (1000, 500), 2: (2000, 1000), 3: (2000, 1000), 4: (1000, 500)}
    custom = bool(args.nx and args.ny)
    nx, ny = (args.nx, args.ny) if custom else default_mesh[order]
    ny_total = ny if args.strong else ny * world
    ly = args.ly if args.ly > 0 else (1.0 if args.strong else float(world))
    W, K = max(args.warmup, 3), args.steps
    headline_is_default = (order == 1 and not custom and not args.local_solver and args.perturb == 0)

    # ---------------- headline: K steps of assemble+condense+scatter, device-resident ----------------
    solves = () if args.no_pcg else (("mg",) if args.pcg == "mg" else (("jacobi", "mg") if order == 1 else ("jacobi", "block", "mg")))
    sampler = ClockSampler(local_rank) if rank == 0 else None
    head, ctx = workload(env, order, qd, nx, ny_total, args.lx, ly, K, W, solves, args.rtol, args.maxit, args.perturb, args.local_solver,
                         sampler=sampler, recover=not args.no_pcg)
    lib = ctx.lib
    s = ctx.sizes()
    ncell = head["elements_per_gpu"]
    kname = f"element_schur_kernel<{order}>" if (order == 1 or os.environ.get("HDG_ELEM_V1")) else f"element_quad_kernel<{order}>"
    if args.local_solver:
        kname = "element_lu_kernel"
    roofline = {"bound": "hbm", "kernel": kname, "achieved": head["kernel_GBps"], "peak": env.peak, "unit": "GB/s",
                "frac": head["frac"], "traffic": None, "alg_bytes_per_element": ALG_BYTES[order],
                "alg_flops_per_element": ALG_FLOPS[order], "kernel_ms": head["kernel_ms"],
                "gflops_alg": ALG_FLOPS[order] * ncell / (head["kernel_ms"] * 1e-3) / 1e9, "peak_source": env.peak_src}
    fp64 = C.c_double()
    hdg.check(lib.hdg_measure_fp64_peak(ctx.h, C.byref(fp64)), ctx.h)
    roofline["fp64_peak_tflops"] = fp64.value          # DFMA microbenchmark of this run (MEASURED_PEAKS.json holds no FP64 figure)
    roofline["fp64_frac_alg"] = roofline["gflops_alg"] / 1e3 / max(fp64.value, 1e-9)
    tj = {}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tj = json.load(fh)
        tr, tn = tj.get(f"element_k{order}"), tj.get(f"element_k{order}_elements")
        # ncu capture of one launch at tn elements; scaled linearly when this run's launch covers a different count
        roofline["traffic"] = (tr * ncell / tn if tr and tn else None)
    except Exception:
        pass
    detail = {"headline": {k: v for k, v in head.items() if k != "clocks"}}

    # trace solve of the headline workload, flat
    if not args.no_pcg:
        nnz_own = int(s.nnz) if world == 1 else (order + 1) ** 2 * (5 * int(s.nface) - 2 * 2 * (nx + ny))
        bytes_iter = 12 * nnz_own + 116 * int(s.ndof)              # CSR byte model of one Jacobi-PCG iteration
        j = head.get("jacobi")
        if j and "error" not in j:
            roofline.update({"pcg_iters": j["iterations"], "pcg_converged": j["converged"], "pcg_ms_per_iter": j["ms_per_iter"],
                             "pcg_solve_s": j["solve_ms"] * 1e-3, "pcg_alg_bytes_per_iter": bytes_iter,
                             "pcg_frac_alg": bytes_iter / (j["ms_per_iter"] * 1e-3) / 1e9 / env.peak})
            if tj.get(f"pcg_k{order}_iteration"):
                trf = tj[f"pcg_k{order}_iteration"] * int(s.ndof) / tj[f"pcg_k{order}_dofs"]
                roofline.update({"pcg_traffic_per_iter": trf, "pcg_frac_traffic": trf / (j["ms_per_iter"] * 1e-3) / 1e9 / env.peak})
        b = head.get("block")
        if b and "error" not in b:
            roofline.update({"blockjacobi_iters": b["iterations"], "blockjacobi_solve_s": b["solve_ms"] * 1e-3})
        m = head.get("mg")
        if m and "error" not in m:
            roofline.update({"mg_iters": m["iterations"], "mg_converged": m["converged"], "mg_solve_ms": m["solve_ms"], "mg_ms_per_iter": m["ms_per_iter"]})
        elif m:
            roofline["mg_error"] = m["error"][:100]
        if "recover_ms" in head:
            roofline.update({"recover_ms": head["recover_ms"], "recover_frac": head["recover_frac"], "err2": head["err2"]})

    # ---------------- end to end through the C ABI with HOST buffers ----------------
    e2e = None
    if not args.no_e2e:
        nyl = ny_total // world if args.strong else ny
        ncell_l = 2 * nx * nyl
        nnode_s, nface_s, nbf_s = (nx + 1) * (nyl + 1), 3 * nx * nyl + nx + nyl, 2 * (nx + nyl)
        cells = torch.empty((ncell_l, 6), dtype=torch.int64).pin_memory().numpy()
        nodes = torch.empty((nnode_s, 2), dtype=torch.float64).pin_memory().numpy()
        faces = torch.empty((4, nface_s), dtype=torch.int64).pin_memory().numpy()      # column-major nface x 4
        bfaces = torch.empty((nbf_s,), dtype=torch.int64).pin_memory().numpy()
        rhs_out = torch.empty((nface_s * (order + 1),), dtype=torch.float64).pin_memory().numpy()
        # host copy of this rank's strip as a self-contained mesh in the Julia layouts (what the ccall shim passes)
        ctxh = env.context(order, qd, comm=False)
        hdg.check(lib.hdg_set_rectangle_mesh(ctxh.h, nx, nyl, 0.0, float(rank), args.lx, float(rank + 1)), ctxh.h)
        hdg.check(lib.hdg_get_mesh(ctxh.h, hdg.api.i64p(cells), hdg.api.f64p(nodes), hdg.api.i64p(faces), hdg.api.i64p(bfaces)), ctxh.h)
        ctxh.close()
        ctx2 = env.context(order, qd, args.local_solver, comm=False)
        parts = [0.0, 0.0, 0.0]      # seconds in hdg_set_mesh / hdg_assemble / hdg_get_rhs (each call returns synchronised)

        def e2e_step():
            t_a = time.perf_counter()
            hdg.check(lib.hdg_set_mesh(ctx2.h, hdg.api.i64p(cells), ncell_l, hdg.api.f64p(nodes), nnode_s,
                                       hdg.api.i64p(faces) if args.e2e_faces else None, nface_s, hdg.api.i64p(bfaces), nbf_s), ctx2.h)
            t_b = time.perf_counter()
            hdg.check(lib.hdg_assemble(ctx2.h), ctx2.h)
            t_c = time.perf_counter()
            hdg.check(lib.hdg_get_rhs(ctx2.h, hdg.api.f64p(rhs_out)), ctx2.h)
            t_d = time.perf_counter()
            parts[0] += t_b - t_a
            parts[1] += t_c - t_b
            parts[2] += t_d - t_c

        ke2e = max(3, min(K, 10))
        for _ in range(2):
            e2e_step()
        env.barrier()
        parts[:] = [0.0, 0.0, 0.0]
        t0 = time.perf_counter()
        for _ in range(ke2e):
            e2e_step()
        torch.cuda.synchronize()
        dt = env.maxf((time.perf_counter() - t0) / ke2e)
        pm = [1e3 * p_ / ke2e for p_ in parts]
        h2d = cells.nbytes + nodes.nbytes + (faces.nbytes if args.e2e_faces else 0) + bfaces.nbytes
        e2e = {"value": ncell_l * world / dt, "unit": "elements/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(rhs_out.nbytes), "ms_per_step": dt * 1e3, "steps": ke2e,
               "set_mesh_ms": pm[0], "assemble_ms": pm[1], "get_rhs_ms": pm[2],
               "h2d_GBps": h2d / max(pm[0], 1e-9) / 1e6, "d2h_GBps": rhs_out.nbytes / max(pm[2], 1e-9) / 1e6,
               "what": "hdg_set_mesh(pinned host arrays in the Julia layouts: cells, nodes, boundary set"
                       + (", faces" if args.e2e_faces else "; mesh.faces is rebuilt on the device") + ") + hdg_assemble + hdg_get_rhs(host)"}
        # the whole driver through the C ABI with host buffers: mesh arrays in, multigrid-PCG solve, recovery,
        # u_h / sigma_h m_values and err2 out - examples/poisson2D_HDG.jl:37-218 end to end (each rank its own strip problem)
        if not args.no_pcg:
            try:
                nb = (order + 1) * (order + 2) // 2
                sig_out = torch.empty((2 * nb, ncell_l), dtype=torch.float64).pin_memory().numpy()
                u_out = torch.empty((nb, ncell_l), dtype=torch.float64).pin_memory().numpy()
                hdg.check(lib.hdg_set_preconditioner(ctx2.h, 2), ctx2.h)
                info_d = hdg.api.SolveInfo()
                err_d = C.c_double()

                dparts = [0.0] * 6      # seconds per call of the driver (every call returns synchronised)

                def driver_step():
                    t = [time.perf_counter()]
                    e2e_step(); t.append(time.perf_counter())
                    hdg.check(lib.hdg_apply_dirichlet(ctx2.h, None), ctx2.h); t.append(time.perf_counter())
                    hdg.check(lib.hdg_solve(ctx2.h, args.rtol, args.maxit, C.byref(info_d)), ctx2.h); t.append(time.perf_counter())
                    hdg.check(lib.hdg_recover(ctx2.h), ctx2.h); t.append(time.perf_counter())
                    hdg.check(lib.hdg_get_mvalues(ctx2.h, hdg.api.f64p(sig_out), hdg.api.f64p(u_out), None), ctx2.h); t.append(time.perf_counter())
                    hdg.check(lib.hdg_errornorm(ctx2.h, 1, C.byref(err_d)), ctx2.h); t.append(time.perf_counter())
                    for i in range(6):
                        dparts[i] += t[i + 1] - t[i]

                driver_step()
                env.barrier()
                dparts[:] = [0.0] * 6
                t0 = time.perf_counter()
                for _ in range(3):
                    driver_step()
                dtd = env.maxf((time.perf_counter() - t0) / 3)
                e2e.update({"driver_ms_per_step": dtd * 1e3, "driver_elements_per_s": ncell_l * world / dtd, "driver_pcg_iterations": int(info_d.iterations),
                            "driver_err2": err_d.value, "driver_solve_device_ms": float(info_d.solve_ms),
                            **{f"driver_{nm}_ms": 1e3 * dparts[i] / 3 for i, nm in enumerate(("e2e_step", "apply", "solve", "recover", "get_mvalues", "errornorm"))}, "driver_d2h_bytes_per_step": int(rhs_out.nbytes + sig_out.nbytes + u_out.nbytes),
                            "driver_what": "e2e step + hdg_apply_dirichlet + hdg_solve (multigrid PCG) + hdg_recover + hdg_get_mvalues(host) + hdg_errornorm; "
                                           "several GPUs: every rank its own strip problem"})
            except Exception as ex:      # keep the headline line even if this extra leg fails
                e2e["driver_error"] = str(ex)[:200]
        ctx2.close()
        del cells, nodes, faces, bfaces, rhs_out
    ctx.close()

    config = {"workload": (f"{'C2' if headline_is_default else 'custom'}: Poisson HDG k={order} quad_degree={qd} tau=1, rectangle_mesh {nx}x{ny_total // world if args.strong else ny} per GPU "
                           f"({ncell} elements, {int(s.ndof)} trace dofs per GPU) - global mesh {nx}x{ny_total} on [0,{args.lx:g}]x[0,{ly:g}]"),
              "parallelism": f"strips of quad rows, {world} rank(s); no data-path collective in assembly; PCG halo + dot products over peer memory (NVLink)",
              "l2": f"inputs+outputs per step {ALG_BYTES[order]*ncell/1e6:.0f} MB > 126 MB L2 (no explicit flush needed)",
              "perturb": args.perturb, "elements_per_gpu": ncell}

    # ---------------- order sweep (one GPU) / strong-scaling configs and parity (several GPUs) ----------------
    if headline_is_default and not args.no_sweep and not args.no_pcg:
        if world == 1:
            for k, (mx, my) in ((2, (2000, 1000)), (3, (2000, 1000)), (4, (1000, 500))):
                try:
                    r, cx = workload(env, k, QD_FOR_ORDER[k], mx, my, 2.0, 1.0, 10, 3, ("mg",), args.rtol, args.maxit, recover=True)
                    cx.close()
                    detail[f"k{k}"] = {kk: vv for kk, vv in r.items() if kk != "clocks"}
                    roofline.update({f"k{k}_elements": r["elements"], f"k{k}_elements_per_s": r["elements_per_s"], f"k{k}_kernel_ms": r["kernel_ms"],
                                     f"k{k}_frac": r["frac"], f"k{k}_recover_frac": r.get("recover_frac"), f"k{k}_err2": r.get("err2")})
                    if "error" not in r["mg"]:
                        roofline.update({f"k{k}_mg_iters": r["mg"]["iterations"], f"k{k}_mg_solve_ms": r["mg"]["solve_ms"]})
                except Exception as ex:
                    roofline[f"k{k}_error"] = str(ex)[:100]
        else:
            try:
                par = parity_leg(env)
                detail["parity"] = par
                config.update({"parity_max_rel": par["max_rel"], "parity_iters_equal": par["iters_equal"], "parity_err2_equal": par["err2_equal"],
                               "parity_what": par["what"]})
            except Exception as ex:
                config["parity_error"] = str(ex)[:120]
        # BASELINE configs defined on several GPUs (strong scaling: the mesh is the config's, split over the ranks)
        strong = []
        if world in (2, 4):
            strong.append(("c3", C3_MESH[0], C3_MESH[1], C3_MESH[2], 2.0, 1.0))
        if world == 8:
            strong.append(("c4", C4_MESH[0], C4_MESH[1], C4_MESH[2], 2.0, 1.0))
        for k, nsq in C5_MESH.items():
            strong.append((f"c5_k{k}", k, nsq, nsq, 1.0, 1.0))
        for name, k, mx, my, lx_, ly_ in strong:
            try:
                r, cx = workload(env, k, QD_FOR_ORDER[k], mx, my, lx_, ly_, 10, 3, ("mg",), args.rtol, args.maxit, recover=True)
                cx.close()
                detail[name] = {kk: vv for kk, vv in r.items() if kk != "clocks"}
                roofline.update({f"{name}_elements": r["elements"], f"{name}_elements_per_s": r["elements_per_s"], f"{name}_frac": r["frac"],
                                 f"{name}_err2": r.get("err2")})
                if "error" not in r["mg"]:
                    roofline.update({f"{name}_mg_iters": r["mg"]["iterations"], f"{name}_mg_solve_ms": r["mg"]["solve_ms"]})
                else:
                    roofline[f"{name}_mg_error"] = r["mg"]["error"][:100]
            except Exception as ex:
                roofline[f"{name}_error"] = str(ex)[:100]

    # ---------------- CPU baseline on the box's host cores (rank 0, N=1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_phases(order, qd)

    if rank == 0:
        out = {
            "metric": "HDG elements/sec (assemble+condense+scatter) and trace PCG solve time vs k", "value": head["elements_per_s"], "unit": "elements/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": head["launches"],
            "clocks": head["clocks"], "detail": detail,
        }
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
