"""hdg_set_mesh on several GPUs: set-up time of an unstructured-style mesh handed over as arrays, per-rank upload,
ghost counts, and parity of the distributed solve with the same problem on one GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 \
        tools/partition_bench.py [nx ny] [--order k] [--shuffle-blocks B]

The mesh is rectangle_mesh(nx, ny) with jittered interior nodes (no grid structure left in the coordinates), passed through
hdg_set_mesh like a parse_mesh_triangle mesh: every rank holds the whole arrays on the host and uploads its part.  With
--shuffle-blocks B the cells are permuted in blocks of B (the numbering keeps locality inside a block only) and the faces are
renumbered by first encounter on the device (hdg_number_faces).  Prints one JSON line on rank 0."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hdg_b200 as hdg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("nx", type=int, nargs="?", default=2000)
    ap.add_argument("ny", type=int, nargs="?", default=1000)
    ap.add_argument("--order", type=int, default=1)
    ap.add_argument("--shuffle-blocks", type=int, default=0)
    ap.add_argument("--morton", action="store_true", help="renumber the (shuffled) mesh along the Morton curve on the device first (hdg_order_cells)")
    ap.add_argument("--no-solve", action="store_true")
    args = ap.parse_args()
    rank, world, lr = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    order, qd = args.order, {1: 2, 2: 4, 3: 6, 4: 9}[args.order]
    # ---- the mesh as host arrays (every rank builds the same one)
    gen = hdg._Context(1, 2, 1.0, 1, lr)
    hdg.check(gen.lib.hdg_set_rectangle_mesh(gen.h, args.nx, args.ny, 0.0, 0.0, 2.0, 1.0), gen.h)
    base = gen.download_mesh()
    gen.close()
    rng = np.random.default_rng(7)
    nodes = base.nodes.copy()
    faces0 = np.asarray(base.faces)
    bnodes = np.unique(faces0[faces0[:, 3] == 0, :2]) - 1
    interior = np.ones(nodes.shape[0], bool)
    interior[bnodes] = False
    h = min(2.0 / args.nx, 1.0 / args.ny)
    nodes[interior] += rng.uniform(-0.2 * h, 0.2 * h, size=(int(interior.sum()), 2))
    if args.shuffle_blocks > 0:
        nb = (base.cells.shape[0] + args.shuffle_blocks - 1) // args.shuffle_blocks
        perm = np.concatenate([np.arange(b * args.shuffle_blocks, min((b + 1) * args.shuffle_blocks, base.cells.shape[0])) for b in rng.permutation(nb)])
        cells, faces = hdg.api.number_faces_gpu(base.cells[perm, :3], nodes)
    else:
        cells, faces = np.ascontiguousarray(base.cells), np.asfortranarray(faces0)
    renumber_ms = None
    if args.morton:
        for _ in range(2):      # second call: without the one-off allocations / module load
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rm = hdg.renumber_mesh(hdg.PolygonalMesh(cells, nodes, faces, {"boundary": set()}))
            renumber_ms = 1e3 * (time.perf_counter() - t0)
        cells, faces = rm.mesh.cells, np.asarray(rm.mesh.faces)
    cells = np.ascontiguousarray(cells, dtype=np.int64)
    faces = np.asfortranarray(faces, dtype=np.int64)
    bf = (np.flatnonzero(faces[:, 3] == 0) + 1).astype(np.int64)
    ncell, nface, nnode = cells.shape[0], faces.shape[0], nodes.shape[0]
    # pinned copies (what a shim with registered arrays would pass)
    pc = torch.from_numpy(cells).pin_memory().numpy()
    pf = torch.from_numpy(np.ascontiguousarray(faces.T)).pin_memory().numpy()      # 4 x nface C-order == nface x 4 column-major
    pn = torch.from_numpy(nodes).pin_memory().numpy()
    pb = torch.from_numpy(bf).pin_memory().numpy()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def set_mesh(ctx):
        hdg.check(ctx.lib.hdg_set_mesh(ctx.h, hdg.api.i64p(pc), ncell, hdg.api.f64p(pn), nnode, hdg.api.i64p(pf), nface, hdg.api.i64p(pb), pb.size), ctx.h)

    ctx = hdg._Context(order, qd, 1.0, 1, lr)
    if world > 1:
        ctx.comm_init(dist, device=torch.device("cuda", lr))
    times = []
    for rep in range(4):
        barrier()
        t0 = time.perf_counter()
        set_mesh(ctx)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        times.append(dt)
    part = ctx.partition()
    own_c, own_f = part["cell_end"] - part["cell_begin"], part["face_end"] - part["face_begin"]
    h2d = 48 * (own_c + part["ghost_cells"]) + 32 * own_f + 16 * nnode + 8 * pb.size
    out = {"mesh": f"jittered rectangle_mesh {args.nx}x{args.ny}" + (f", cells shuffled in blocks of {args.shuffle_blocks}" if args.shuffle_blocks else ", natural order")
           + (", renumbered along the Morton curve (hdg_order_cells + hdg_number_faces)" if args.morton else ""),
           "renumber_ms": renumber_ms,
           "ncell": ncell, "nface": nface, "n_gpus": world, "order": order,
           "set_mesh_ms_first": 1e3 * times[0], "set_mesh_ms": 1e3 * float(np.min(times[1:])),
           "rank0_owned_cells": own_c, "rank0_ghost_cells": part["ghost_cells"], "rank0_ghost_faces": part["ghost_faces"],
           "rank0_h2d_bytes": int(h2d), "whole_mesh_bytes": int(48 * ncell + 32 * nface + 16 * nnode + 8 * pb.size),
           "rank0_h2d_fraction": h2d / (48 * ncell + 32 * nface + 16 * nnode + 8 * pb.size)}
    if not args.no_solve:
        def solve(cx):
            lib = cx.lib
            hdg.check(lib.hdg_assemble(cx.h), cx.h)
            hdg.check(lib.hdg_apply_dirichlet(cx.h, None), cx.h)
            hdg.check(lib.hdg_set_preconditioner(cx.h, 1), cx.h)
            info = hdg.api.SolveInfo()
            hdg.check(lib.hdg_solve(cx.h, 1e-12, 200000, C.byref(info)), cx.h)
            hdg.check(lib.hdg_recover(cx.h), cx.h)
            e = C.c_double()
            hdg.check(lib.hdg_errornorm(cx.h, 1, C.byref(e)), cx.h)
            x = np.empty(cx.sizes().ndof)
            hdg.check(lib.hdg_get_trace(cx.h, hdg.api.f64p(x)), cx.h)
            return x, info, e.value
        xm, im, em = solve(ctx)
        ref = hdg._Context(order, qd, 1.0, 1, lr)      # the same problem on ONE GPU (every rank, its own device)
        set_mesh(ref)
        x1, i1, e1 = solve(ref)
        ref.close()
        nt = order + 1
        own = x1[part["face_begin"] * nt: part["face_end"] * nt]
        d = float(np.abs(xm - own).max())
        if world > 1:
            t = torch.tensor([d], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            d = float(t.item())
        out.update({"solve_iterations": [int(im.iterations), int(i1.iterations)], "solve_ms": [float(im.solve_ms), float(i1.solve_ms)],
                    "uhat_max_rel_vs_1gpu": d / float(np.abs(x1).max()), "err2": [em, e1]})
    ctx.close()
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
