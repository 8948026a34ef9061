#!/usr/bin/env python
"""Host emulation of csrc/hdg_mg_general.cuh: the kernel bodies are
compiled for the CPU (-DHDG_HOST_EMU), run in serial loops over their index, and checked against scipy:
  * the ELL vertex operator equals P'AP,
  * restrict -> m Chebyshev steps -> prolong equals the numpy formulation of tools/cheb_prototype.py,
  * PCG with the emulated preconditioner converges in the same number of iterations.
    python tools/check_mg_general.py [--n 24] [--delaunay 3000 --lattice] [--k 2]
"""
import argparse
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import amg_prototype as ap  # noqa: E402
from mg_prototype import block_jacobi, pcg, prolongation  # noqa: E402

SRC = r'''
#define HDG_HOST_EMU 1
#include "hdg_mg_general.cuh"
using namespace hdg;
extern "C" {
int emu_maxval() { return MGX_MAXVAL; }
double emu_build(int64_t nv, int NT, const double* Kd, const double* Ko, const int32_t* kcol, const uint8_t* isbc,
                 const int32_t* facenode, const int32_t* vcnt, const int32_t* vface, int32_t* nbr, double* diag, double* val, double* dinv) {
    for (int64_t v = 0; v < nv; ++v) mgx_neighbours_row(v, vcnt, vface, facenode, nbr);
    for (int64_t v = 0; v < nv; ++v) mgx_operator_row(v, NT, Kd, Ko, kcol, isbc, facenode, vcnt, vface, nbr, diag, val);
    double lmax = 0.0;
    for (int64_t v = 0; v < nv; ++v) { double g = mgx_dinv_row(v, vcnt, diag, val, dinv); if (g > lmax) lmax = g; }
    return lmax;
}
void emu_apply(int64_t nv, int64_t nface, int NT, int m, double lmin, double lmax, const uint8_t* isbc, const int32_t* facenode,
               const int32_t* vcnt, const int32_t* vface, const int32_t* nbr, const double* diag, const double* val, const double* dinv,
               const double* r, double* z, double* x, double* res, double* d) {
    for (int64_t v = 0; v < nv; ++v) { res[v] = mgx_restrict_row(v, NT, vcnt, vface, r); x[v] = 0.0; }
    const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
    double rho = 1.0 / sigma;
    for (int64_t v = 0; v < nv; ++v) d[v] = dinv[v] * res[v] / theta;
    for (int i = 0; i < m; ++i) {
        for (int64_t v = 0; v < nv; ++v) mgx_cheb_residual_row(v, vcnt, nbr, diag, val, d, x, res);   // reads the old d of the neighbours: d is not written here
        const double rho_new = 1.0 / (2.0 * sigma - rho);
        for (int64_t v = 0; v < nv; ++v) mgx_cheb_direction_row(v, dinv, res, rho_new * rho, 2.0 * rho_new / delta, d);
        rho = rho_new;
    }
    for (int64_t f = 0; f < nface; ++f) mgx_prolong_face(f, NT, facenode, isbc, x, z);
}
}
'''


def build_lib():
    tmp = tempfile.mkdtemp()
    src = os.path.join(tmp, "emu.cpp")
    open(src, "w").write(SRC)
    so = os.path.join(tmp, "emu.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "hdiscontinuousgalerkin.jl_b200", "csrc"),
                           src, "-o", so])
    return C.CDLL(so)


def block_ell(mesh, Kc, nt):
    """Kd / Ko / kcol of csrc/hdg_internal.h from the CSC matrix: slots 0,1 = the other two faces of the face's first cell in
    cyclic order, slots 2,3 = the same for the second cell; blocks column-major."""
    nf = mesh.nfaces
    Kr = Kc.tocsr()
    Kd = np.zeros((nf, nt * nt))
    Ko = np.zeros((nf, 4, nt * nt))
    kcol = -np.ones((nf, 4), np.int32)
    dense = lambda f, g: Kr[f * nt:(f + 1) * nt, g * nt:(g + 1) * nt].toarray()
    for f in range(nf):
        Kd[f] = dense(f, f).T.ravel()                   # blk[b*nt + a] = K[a][b]
        for side in (0, 1):
            cell = mesh.faces[f, 2 + side]
            if cell == 0:
                continue
            cf = list(mesh.cell_faces[cell - 1] - 1)
            l = cf.index(f)
            for s in (0, 1):
                g = cf[(l + 1 + s) % 3]
                kcol[f, 2 * side + s] = g
                Ko[f, 2 * side + s] = dense(f, g).T.ravel()
    return Kd.ravel(), Ko.ravel(), kcol.ravel()


def adjacency(mesh, isbc_face, maxval):
    nv, nf = mesh.nnodes, mesh.nfaces
    lists = [[] for _ in range(nv)]
    for f in range(nf):
        v1, v2 = mesh.faces[f, 0] - 1, mesh.faces[f, 1] - 1
        lo, hi = min(v1, v2), max(v1, v2)
        lists[lo].append(f)
        lists[hi].append(f | 0x80000000)
    vcnt = np.zeros(nv, np.int32)
    vface = np.zeros((nv, maxval), np.uint32)
    for v in range(nv):
        es = sorted(lists[v], key=lambda e: e & 0x7fffffff)
        assert len(es) <= maxval
        vface[v, :len(es)] = es
        fixed = any(isbc_face[e & 0x7fffffff] for e in es)
        vcnt[v] = -1 if (fixed or not es) else len(es)
    return vcnt, vface.view(np.int32).ravel()


def main():
    ap_ = argparse.ArgumentParser()
    ap_.add_argument("--n", type=int, default=20)
    ap_.add_argument("--delaunay", type=int, default=0)
    ap_.add_argument("--lattice", action="store_true")
    ap_.add_argument("--k", type=int, default=1)
    ap_.add_argument("--seed", type=int, default=3)
    ap_.add_argument("--m", type=int, default=16)
    ap_.add_argument("--alpha", type=float, default=100.0)
    a = ap_.parse_args()
    lib = build_lib()
    lib.emu_build.restype = C.c_double
    MAXVAL = lib.emu_maxval()
    mesh = ap.make_mesh(a)
    qd = {1: 2, 2: 4, 3: 6, 4: 9}[a.k]
    tab, A, b, isbc = ap.build_system(mesh, a.k, qd)
    nt = tab.nt
    nv, nf = mesh.nnodes, mesh.nfaces
    isbc_face = np.ascontiguousarray(isbc.reshape(nf, nt)[:, 0].astype(np.uint8))
    Kc = (sp.diags(np.where(isbc, 1.0, -1.0)) @ A).tocsc()               # back to the applied K (A = D K)
    Kd, Ko, kcol = block_ell(mesh, Kc, nt)
    facenode = np.ascontiguousarray((mesh.faces[:, :2] - 1).astype(np.int32)).ravel()
    vcnt, vface = adjacency(mesh, isbc_face, MAXVAL)
    nbr = np.empty(nv * MAXVAL, np.int32)
    diag, dinv = np.empty(nv), np.empty(nv)
    val = np.empty(nv * MAXVAL)
    p = lambda x, t: x.ctypes.data_as(C.POINTER(t))
    lmax = lib.emu_build(C.c_int64(nv), C.c_int(nt), p(Kd, C.c_double), p(Ko, C.c_double), p(kcol, C.c_int32), p(isbc_face, C.c_uint8),
                         p(facenode, C.c_int32), p(vcnt, C.c_int32), p(vface, C.c_int32), p(nbr, C.c_int32), p(diag, C.c_double),
                         p(val, C.c_double), p(dinv, C.c_double))
    # ---- 1. the ELL operator equals P'AP on the free vertices
    P, bnode = prolongation(mesh, nt, isbc)
    Ac = (P.T @ A @ P).tocsr()
    rows = np.repeat(np.arange(nv), MAXVAL)
    ok = nbr >= 0
    Aell = sp.coo_matrix((val[ok], (rows[ok], nbr[ok])), shape=(nv, nv)).tocsr() + sp.diags(diag)
    err = abs(Aell - Ac).max() / abs(Ac).max()
    assert np.array_equal(vcnt < 0, bnode), "fixed vertices differ"
    print(f"{mesh.ncells} cells, k={a.k}: |A_ell - P'AP| / |P'AP| = {err:.2e}   Gershgorin lmax = {lmax:.4f}")
    assert err < 1e-13
    # ---- 2. restrict -> Chebyshev -> prolong against numpy
    lmin = lmax / a.alpha
    dj = np.where(bnode, 0.0, 1.0 / np.where(Ac.diagonal() != 0, Ac.diagonal(), 1.0))

    def cheb_np(r):
        th, de = (lmax + lmin) / 2, (lmax - lmin) / 2
        x = np.zeros_like(r); res = r.copy(); sig = th / de; rho = 1 / sig
        d = dj * res / th
        for _ in range(a.m):
            x += d
            res -= Ac @ d
            rn = 1 / (2 * sig - rho)
            d = rn * rho * d + 2 * rn / de * (dj * res)
            rho = rn
        return x

    xb, rb, db = np.empty(nv), np.empty(nv), np.empty(nv)

    def vertex_term(r):
        z = np.zeros_like(r)
        lib.emu_apply(C.c_int64(nv), C.c_int64(nf), C.c_int(nt), C.c_int(a.m), C.c_double(lmin), C.c_double(lmax), p(isbc_face, C.c_uint8),
                      p(facenode, C.c_int32), p(vcnt, C.c_int32), p(vface, C.c_int32), p(nbr, C.c_int32), p(diag, C.c_double), p(val, C.c_double),
                      p(dinv, C.c_double), p(np.ascontiguousarray(r), C.c_double), p(z, C.c_double), p(xb, C.c_double), p(rb, C.c_double), p(db, C.c_double))
        return z

    r = np.random.default_rng(1).standard_normal(A.shape[0])
    r[isbc] = 0.0
    z_emu, z_np = vertex_term(r), P @ cheb_np(P.T @ r)
    e2 = np.abs(z_emu - z_np).max() / np.abs(z_np).max()
    print(f"   restrict -> Chebyshev({a.m}) -> prolong: emulated kernels vs numpy {e2:.2e}")
    assert e2 < 1e-11
    # ---- 3. PCG with the emulated preconditioner
    bj = block_jacobi(A, nt)
    x0, it0 = pcg(A, b, bj, maxit=20000)
    x1, it1 = pcg(A, b, lambda rr: bj(rr) + vertex_term(rr))
    x2, it2 = pcg(A, b, lambda rr: bj(rr) + P @ cheb_np(P.T @ rr))
    print(f"   PCG iterations: block-Jacobi {it0}, + emulated vertex term {it1} (numpy formulation {it2}); |x - x_bj| / |x_bj| = {np.linalg.norm(x1 - x0) / np.linalg.norm(x0):.1e}")
    assert abs(it1 - it2) <= 1 and np.linalg.norm(x1 - x0) < 1e-9 * np.linalg.norm(x0)
    print("OK")


if __name__ == "__main__":
    main()
