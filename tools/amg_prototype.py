#!/usr/bin/env python
"""CPU prototype (scipy) of an ALGEBRAIC hierarchy for the P1-vertex space of the multigrid preconditioner, for meshes that
are not the triangulation of rectangle_mesh (SURVEY 8(f) rank 1 x rank 2; DESIGN.md "next").  Not part of the product: it
picks the algorithm the device code will implement and records the iteration counts to expect.

  A_c = P'AP                      the vertex operator (general sparse, one row per mesh vertex, vertex-degree + 1 entries)
  aggregation                     MIS(2) roots with hashed weights (the parallel scheme of Bell, Dalton & Olson: every step
                                  is a max-propagation over graph neighbours, i.e. gather kernels, no sequential sweep),
                                  remaining vertices join the aggregate of their strongest already-aggregated neighbour
  tentative prolongation          piecewise constant over aggregates
  smoothed prolongation           (I - w D^-1 A) P_tent, w = 2/3         (smoothed aggregation)
  Galerkin coarse operators, damped-Jacobi V(1,1), dense solve on the coarsest level

  python tools/amg_prototype.py --n 60 --k 1            jittered + randomly renumbered mesh, ~2 n^2 cells
  python tools/amg_prototype.py --delaunay 4000 --k 2   Delaunay mesh of 4000 random points in the unit square
"""
import argparse
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import hdg_oracle as orc  # noqa: E402
import hdg_oracle_c as occ  # noqa: E402
from mg_prototype import block_jacobi, pcg, prolongation  # noqa: E402


def number_faces(tri):
    """first-encounter face numbering (vectorised; same as api.number_faces)"""
    tri = np.asarray(tri, dtype=np.int64)
    nc = tri.shape[0]
    k1, k2 = np.array([1, 2, 0]), np.array([2, 0, 1])
    v1, v2 = tri[:, k1].reshape(-1), tri[:, k2].reshape(-1)
    lo, hi = np.minimum(v1, v2), np.maximum(v1, v2)
    key = lo * (tri.max() + 1) + hi
    order = np.argsort(key, kind="stable")
    sk = key[order]
    start = np.ones(sk.size, bool)
    start[1:] = sk[1:] != sk[:-1]
    end = np.ones(sk.size, bool)
    end[:-1] = start[1:]
    grp = np.cumsum(start) - 1
    first_pos, last_pos = order[start], order[end]
    by_first = np.argsort(first_pos, kind="stable")
    rank = np.empty(first_pos.size, np.int64)
    rank[by_first] = np.arange(first_pos.size)
    fop = np.empty(sk.size, np.int64)
    fop[order] = rank[grp]
    fp, lp = first_pos[by_first], last_pos[by_first]
    faces = np.zeros((fp.size, 4), np.int64)
    faces[:, 0], faces[:, 1] = v1[fp], v2[fp]
    faces[:, 2] = fp // 3 + 1
    faces[:, 3] = np.where(lp != fp, lp // 3 + 1, 0)
    return (fop + 1).reshape(nc, 3), faces


def make_mesh(args):
    rng = np.random.default_rng(args.seed)
    if args.delaunay:
        from scipy.spatial import Delaunay
        m = int(np.sqrt(args.delaunay))
        edge = np.linspace(0.0, 1.0, m + 1)
        bpts = np.concatenate([np.c_[edge, 0 * edge], np.c_[edge, 0 * edge + 1], np.c_[0 * edge[1:-1], edge[1:-1]], np.c_[0 * edge[1:-1] + 1, edge[1:-1]]])
        if args.lattice:      # well-shaped triangles: a jittered hexagonal lattice instead of uniformly random points
            hx = 1.0 / m
            gx, gy = np.meshgrid(np.arange(1, m) * hx, np.arange(1, int(m / 0.866)) * hx * 0.866)
            gx = gx + 0.5 * hx * (np.arange(gx.shape[0])[:, None] % 2)
            ip = np.c_[gx.ravel(), gy.ravel()]
            ip = ip[(ip[:, 0] > 0.4 * hx) & (ip[:, 0] < 1 - 0.4 * hx) & (ip[:, 1] > 0.4 * hx) & (ip[:, 1] < 1 - 0.4 * hx)]
            ip += rng.uniform(-0.15 * hx, 0.15 * hx, size=ip.shape)
        else:
            ip = rng.uniform(0.02, 0.98, size=(args.delaunay, 2))
        pts = np.concatenate([bpts, ip])
        tri = Delaunay(pts).simplices.astype(np.int64) + 1
        nodes = pts
    else:
        base = orc.rectangle_mesh(args.n, args.n)
        nodes = base.nodes.copy()
        h = 1.0 / args.n
        interior = (nodes[:, 0] > 1e-12) & (nodes[:, 0] < 1 - 1e-12) & (nodes[:, 1] > 1e-12) & (nodes[:, 1] < 1 - 1e-12)
        nodes[interior] += rng.uniform(-0.25 * h, 0.25 * h, size=(interior.sum(), 2))
        perm = rng.permutation(nodes.shape[0])              # random vertex renumbering: no grid structure left
        inv = np.empty_like(perm)
        inv[perm] = np.arange(perm.size)
        tri = (inv[base.cells - 1] + 1)[rng.permutation(base.ncells)]
        nodes = nodes[perm]
    p = nodes[tri - 1]
    a, b = p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]
    flip = (a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]) < 0
    tri[flip] = tri[flip][:, [0, 2, 1]]
    cf, faces = number_faces(tri)
    bset = set((np.flatnonzero(faces[:, 3] == 0) + 1).tolist())
    return orc.Mesh(nodes=nodes, cells=tri, cell_faces=cf, faces=faces, facesets={"boundary": bset})


def build_system(mesh, k, qd):
    tab = orc.build_tables(k, qd)
    K, rhs, _, _ = occ.doassemble(mesh, tab, 1.0, None, occ.max_threads(), keep_local=False)
    nt = tab.nt
    bf = mesh.boundary_faces_sorted()
    dofs = (np.repeat(bf, nt) - 1) * nt + np.tile(np.arange(nt), bf.size) + 1
    Kc, b, m = orc.apply_dirichlet(K, rhs, dofs, np.zeros(dofs.size))
    isbc = np.zeros(K.shape[0], bool)
    isbc[dofs - 1] = True
    D = sp.diags(np.where(isbc, 1.0, -1.0))
    return tab, (D @ Kc).tocsr(), D @ b, isbc


# ---- MIS(2) aggregation, written as the gather kernels the device would run ------------------------------------------------
def mis2_aggregate(A, free, seed=1):
    """A: CSR strength graph (off-diagonal pattern), free: mask of vertices that carry an unknown.
    Returns agg (aggregate id per vertex, -1 for fixed vertices) and the number of aggregates."""
    n = A.shape[0]
    G = A.copy().tocsr()
    G.setdiag(0)
    G.eliminate_zeros()
    indptr, indices = G.indptr, G.indices
    # hashed weights (deterministic, no ties): state 1 = undecided, 2 = root, 0 = removed
    w = (np.arange(n, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed)) >> np.uint64(11)
    key = (w.astype(np.int64) << 1)
    state = np.where(free, 1, 0).astype(np.int64)
    row = np.repeat(np.arange(n), np.diff(indptr))

    def nbr_max(val):          # max over the closed neighbourhood (one gather kernel)
        out = val.copy()
        np.maximum.at(out, row, val[indices])
        return out

    while (state == 1).any():
        t = np.where(state == 1, key + 1, np.where(state == 2, np.iinfo(np.int64).max, -1))   # roots dominate, removed vertices never win
        t1 = nbr_max(t)
        t2 = nbr_max(t1)                                    # distance-2 maximum
        new_root = (state == 1) & (t2 == key + 1)
        state[new_root] = 2
        # everything within distance 2 of a root is removed from the candidate set
        r0 = np.where(state == 2, 1, 0)
        r2 = nbr_max(nbr_max(r0))
        state[(state == 1) & (r2 == 1)] = 0
    roots = np.flatnonzero(state == 2)
    agg = np.full(n, -1, np.int64)
    agg[roots] = np.arange(roots.size)
    # pass 1: neighbours of a root join it; pass 2+: join the aggregate of the strongest aggregated neighbour
    absA = abs(G).tocsr()
    for _ in range(4):
        todo = free & (agg < 0)
        if not todo.any():
            break
        best = np.full(n, -1.0)
        choice = np.full(n, -1, np.int64)
        val = absA.data
        cand = agg[indices]
        ok = cand >= 0
        # strongest aggregated neighbour per row (gather)
        for r_, c_, v_ in zip(row[ok], cand[ok], val[ok]):
            if todo[r_] and v_ > best[r_]:
                best[r_] = v_
                choice[r_] = c_
        agg[todo & (choice >= 0)] = choice[todo & (choice >= 0)]
    left = free & (agg < 0)                                  # isolated leftovers: singletons
    agg[left] = roots.size + np.arange(left.sum())
    return agg, roots.size + int(left.sum())


class SAMG:
    def __init__(self, Ac, free, omega_p=2.0 / 3.0, omega=0.8, nu=1, min_pts=64, smooth_p=True):
        self.levels = []
        A = Ac.tocsr()
        self.nu, self.omega = nu, omega
        while True:
            d = A.diagonal()
            lev = dict(A=A, dinv=np.where(free, 1.0 / np.where(d != 0, d, 1.0), 0.0))
            self.levels.append(lev)
            if free.sum() <= min_pts or len(self.levels) > 20:
                break
            agg, nagg = mis2_aggregate(A, free)
            rows = np.flatnonzero(agg >= 0)
            T = sp.coo_matrix((np.ones(rows.size), (rows, agg[rows])), shape=(A.shape[0], nagg)).tocsr()
            if smooth_p:
                Df = sp.diags(lev["dinv"])
                Pl = (T - omega_p * (Df @ (A @ T))).tocsr()
            else:
                Pl = T
            lev["P"] = Pl
            A = (Pl.T @ A @ Pl).tocsr()
            free = np.ones(nagg, bool)
        last = self.levels[-1]["A"].tocsc()
        self.coarse = spla.splu(last + sp.diags((last.diagonal() == 0).astype(float)).tocsc())
        self.sizes = [int(l["A"].shape[0]) for l in self.levels]
        self.nnz_row = [l["A"].nnz / max(l["A"].shape[0], 1) for l in self.levels]

    def vcycle(self, r, l=0):
        lev = self.levels[l]
        if l == len(self.levels) - 1:
            return self.coarse.solve(r)
        A, dinv = lev["A"], lev["dinv"]
        x = self.omega * dinv * r
        for _ in range(self.nu - 1):
            x += self.omega * dinv * (r - A @ x)
        x += lev["P"] @ self.vcycle(lev["P"].T @ (r - A @ x), l + 1)
        for _ in range(self.nu):
            x += self.omega * dinv * (r - A @ x)
        return x


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=40)
    ap.add_argument("--delaunay", type=int, default=0)
    ap.add_argument("--k", type=int, default=1)
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--unsmoothed", action="store_true")
    ap.add_argument("--lattice", action="store_true", help="with --delaunay: jittered hexagonal lattice (well-shaped triangles)")
    ap.add_argument("--nu", type=int, default=1)
    a = ap.parse_args()
    qd = {1: 2, 2: 4, 3: 6, 4: 9}[a.k]
    t0 = time.time()
    mesh = make_mesh(a)
    tab, A, b, isbc = build_system(mesh, a.k, qd)
    nt = tab.nt
    print(f"{mesh.ncells} cells, {mesh.nnodes} vertices, k={a.k}: {A.shape[0]} trace dofs, assembled in {time.time() - t0:.1f} s")
    P, bnode = prolongation(mesh, nt, isbc)
    Ac = (P.T @ A @ P).tocsr() + sp.diags(bnode.astype(float))
    bj = block_jacobi(A, nt)
    x0, it_bj = pcg(A, b, bj, maxit=20000)
    print(f"  block-Jacobi PCG: {it_bj} iterations")
    exact = spla.splu(Ac.tocsc())
    x2, it2 = pcg(A, b, lambda r: bj(r) + P @ exact.solve(P.T @ r))
    print(f"  additive, exact vertex solve: {it2} iterations")
    t0 = time.time()
    mg = SAMG(Ac, ~bnode, smooth_p=not a.unsmoothed, nu=a.nu)
    print(f"  hierarchy in {time.time() - t0:.1f} s: sizes {mg.sizes}, nnz/row {[round(x, 1) for x in mg.nnz_row]}")
    x3, it3 = pcg(A, b, lambda r: bj(r) + P @ mg.vcycle(P.T @ r))
    print(f"  additive, {'unsmoothed' if a.unsmoothed else 'smoothed'}-aggregation V({a.nu},{a.nu}): {it3} iterations, "
          f"|x - x_bj| / |x_bj| = {np.linalg.norm(x3 - x0) / np.linalg.norm(x0):.2e}")


if __name__ == "__main__":
    main()
