#!/usr/bin/env python
"""CPU prototype (scipy) of the P1-vertex multigrid preconditioner for the HDG trace system - SURVEY 8(f) rank 1.
Used to pick the algorithm that hdg_set_preconditioner(ctx, 2) implements on the device; not part of the product.

  A  = sign-fixed condensed trace matrix after apply! (SPD)
  P  = trace coefficients of the P1 function with given vertex values (modes 0 and 1 of the Legendre trace basis)
  Ac = P'AP on the vertex grid, coarsened geometrically with Galerkin products (structured rectangle_mesh)
  M^-1 r = omega * Binv r  +  P * Vcycle(P' r)             (additive)  or the symmetric multiplicative variant
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
import hdg_oracle as orc  # noqa: E402
import hdg_oracle_c as occ  # noqa: E402


def build_system(nx, ny, k, qd, UR=(2.0, 1.0)):
    mesh = orc.rectangle_mesh(nx, ny, (0.0, 0.0), UR)
    tab = orc.build_tables(k, qd)
    K, rhs, _, _ = occ.doassemble(mesh, tab, 1.0, None, occ.max_threads(), keep_local=False)
    nt = tab.nt
    bf = mesh.boundary_faces_sorted()
    dofs = (np.repeat(bf, nt) - 1) * nt + np.tile(np.arange(nt), bf.size) + 1
    Kc, b, m = orc.apply_dirichlet(K, rhs, dofs, np.zeros(dofs.size))
    isbc = np.zeros(K.shape[0], bool)
    isbc[dofs - 1] = True
    D = sp.diags(np.where(isbc, 1.0, -1.0))
    A = (D @ Kc).tocsr()
    b = D @ b
    return mesh, tab, A, b, isbc


def prolongation(mesh, nt, isbc):
    """P: ndof x nnode.  Face f with vertices lo < hi (canonical trace direction lo -> hi): the linear function with vertex
    values (a, b) has Legendre coefficients  mode 0: (a+b)/2,  mode 1: (b-a)/(2 sqrt 3)."""
    nf = mesh.nfaces
    v1, v2 = mesh.faces[:, 0] - 1, mesh.faces[:, 1] - 1
    lo, hi = np.minimum(v1, v2), np.maximum(v1, v2)
    rows, cols, vals = [], [], []
    f = np.arange(nf)
    rows += [f * nt, f * nt]; cols += [lo, hi]; vals += [np.full(nf, 0.5), np.full(nf, 0.5)]
    if nt > 1:
        c = 1.0 / (2.0 * np.sqrt(3.0))
        rows += [f * nt + 1, f * nt + 1]; cols += [lo, hi]; vals += [np.full(nf, -c), np.full(nf, c)]
    P = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nf * nt, mesh.nnodes)).tocsr()
    # boundary vertices carry no coarse dof, Dirichlet trace rows get no correction
    bnode = np.zeros(mesh.nnodes, bool)
    bfaces = mesh.boundary_faces_sorted() - 1
    bnode[v1[bfaces]] = True; bnode[v2[bfaces]] = True
    P = sp.diags(np.where(isbc, 0.0, 1.0)) @ P @ sp.diags(np.where(bnode, 0.0, 1.0))
    return P.tocsr(), bnode


def grid_interp(px, py, free):
    """P1 interpolation on the structured triangulation (diagonal from (i+1,j) to (i,j+1)) from the coarse grid of the
    points with even indices.  px, py = number of points per direction; free = mask (py, px) of free points."""
    cx, cy = (px + 1) // 2, (py + 1) // 2
    rows, cols, vals = [], [], []
    for j in range(py):
        for i in range(px):
            if not free[j, i]:
                continue
            I, J, a, b = i // 2, j // 2, i % 2, j % 2
            if a == 0 and b == 0:
                src = [(I, J, 1.0)]
            elif a == 1 and b == 0:
                src = [(I, J, 0.5), (I + 1, J, 0.5)]
            elif a == 0 and b == 1:
                src = [(I, J, 0.5), (I, J + 1, 0.5)]
            else:
                src = [(I + 1, J, 0.5), (I, J + 1, 0.5)]
            for (ii, jj, w) in src:
                if ii < cx and jj < cy:
                    rows.append(j * px + i); cols.append(jj * cx + ii); vals.append(w)
    Pg = sp.coo_matrix((vals, (rows, cols)), shape=(px * py, cx * cy)).tocsr()
    return Pg, cx, cy


class VertexMG:
    def __init__(self, Ac, px, py, free, nu=2, omega=0.8, min_pts=200, cheb=0):
        self.levels = []
        A = Ac.tocsr()
        while True:
            dinv = 1.0 / A.diagonal()
            lev = dict(A=A, dinv=dinv, px=px, py=py, free=free)
            self.levels.append(lev)
            if px * py <= min_pts or min(px, py) <= 3:
                break
            Pg, cx, cy = grid_interp(px, py, free)
            # coarse free mask: a coarse point is free if its fine twin is
            cfree = free[::2, ::2].copy()
            Pg = Pg @ sp.diags(cfree.ravel().astype(float))
            Acoarse = (Pg.T @ A @ Pg).tocsr()
            # fixed points: identity rows
            d = Acoarse.diagonal()
            fix = (~cfree.ravel()) | (d == 0)
            Acoarse = Acoarse + sp.diags(fix.astype(float))
            lev["P"] = Pg
            A, px, py, free = Acoarse.tocsr(), cx, cy, (cfree & ~fix.reshape(cfree.shape))
        self.nu, self.omega = nu, omega
        self.coarse = spla.splu(self.levels[-1]["A"].tocsc())
        self.max_stencil = max((np.diff(l["A"].indptr).max() for l in self.levels))

    def vcycle(self, r, l=0):
        lev = self.levels[l]
        if l == len(self.levels) - 1:
            return self.coarse.solve(r)
        A, dinv = lev["A"], lev["dinv"]
        x = np.zeros_like(r)
        for _ in range(self.nu):
            x += self.omega * dinv * (r - A @ x)
        rc = lev["P"].T @ (r - A @ x)
        x += lev["P"] @ self.vcycle(rc, l + 1)
        for _ in range(self.nu):
            x += self.omega * dinv * (r - A @ x)
        return x


def pcg(A, b, M, rtol=1e-12, maxit=5000):
    x = np.zeros_like(b)
    r = b.copy()
    z = M(r)
    p = z.copy()
    rz = r @ z
    bn = np.linalg.norm(b)
    for it in range(1, maxit + 1):
        Ap = A @ p
        al = rz / (p @ Ap)
        x += al * p
        r -= al * Ap
        if np.linalg.norm(r) <= rtol * bn:
            return x, it
        z = M(r)
        rz2 = r @ z
        p = z + (rz2 / rz) * p
        rz = rz2
    return x, maxit


def block_jacobi(A, nt):
    n = A.shape[0]
    nb = n // nt
    Ab = A.tobsr(blocksize=(nt, nt))
    inv = np.zeros((nb, nt, nt))
    for i in range(nb):
        for p in range(Ab.indptr[i], Ab.indptr[i + 1]):
            if Ab.indices[p] == i:
                inv[i] = np.linalg.inv(Ab.data[p])
    return lambda r: np.einsum("bij,bj->bi", inv, r.reshape(nb, nt)).ravel()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=64)
    ap.add_argument("--ny", type=int, default=32)
    ap.add_argument("--k", type=int, default=1)
    ap.add_argument("--omega", type=float, default=0.8)
    ap.add_argument("--womega", type=float, default=1.0, help="weight of the trace-level smoother term")
    ap.add_argument("--nu", type=int, default=2)
    a = ap.parse_args()
    qd = {1: 2, 2: 4, 3: 6, 4: 9}[a.k]
    t0 = time.time()
    mesh, tab, A, b, isbc = build_system(a.nx, a.ny, a.k, qd)
    nt = tab.nt
    print(f"mesh {a.nx}x{a.ny} k={a.k}: ndof {A.shape[0]}  assembled in {time.time() - t0:.1f}s")
    P, bnode = prolongation(mesh, nt, isbc)
    Ac = (P.T @ A @ P).tocsr() + sp.diags(bnode.astype(float))
    px, py = a.nx + 1, a.ny + 1
    free = (~bnode).reshape(py, px)
    dinv = 1.0 / A.diagonal()
    jac = lambda r: dinv * r
    bj = block_jacobi(A, nt)
    x0, it = pcg(A, b, jac)
    print(f"  Jacobi PCG          : {it} iterations")
    x1, it = pcg(A, b, bj)
    print(f"  block-Jacobi PCG    : {it} iterations")
    exact = spla.splu(Ac.tocsc())
    for name, sm in (("jacobi", jac), ("block-jacobi", bj)):
        M2 = lambda r: a.womega * sm(r) + P @ exact.solve(P.T @ r)
        x2, it = pcg(A, b, M2)
        print(f"  additive 2-level ({name:12s}, exact coarse): {it} iterations, |x-xj|/|xj| = {np.linalg.norm(x2 - x0) / np.linalg.norm(x0):.2e}")
    mg = VertexMG(Ac, px, py, free, nu=a.nu, omega=a.omega)
    print(f"  vertex MG: {len(mg.levels)} levels, sizes {[l['px'] * l['py'] for l in mg.levels]}, max stencil {mg.max_stencil}")
    for name, sm in (("jacobi", jac), ("block-jacobi", bj)):
        M3 = lambda r: a.womega * sm(r) + P @ mg.vcycle(P.T @ r)
        x3, it = pcg(A, b, M3)
        print(f"  additive multilevel ({name:12s}, V({a.nu},{a.nu}))  : {it} iterations, |x-xj|/|xj| = {np.linalg.norm(x3 - x0) / np.linalg.norm(x0):.2e}")

        def M4(r, sm=sm):
            z = sm(r)
            z = z + P @ mg.vcycle(P.T @ (r - A @ z))
            return z + sm(r - A @ z)
        x4, it = pcg(A, b, M4)
        print(f"  multiplicative      ({name:12s}, V({a.nu},{a.nu}))  : {it} iterations (3 SpMV each)")


if __name__ == "__main__":
    main()
