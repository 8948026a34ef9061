"""Full-size parity: the CUDA path through the C ABI against the C oracle port at BASELINE.json's own sizes.

C2 = k=1 on the 1M-element mesh (1000 x 500), C3 per-GPU size = k=2 on the 4M-element mesh (2000 x 1000).  The small-mesh tests
in test_gpu_parity.py compare every cell with the numpy oracle; here the same quantities are compared where the 32-cell tile
layout, the 64-bit offsets and the grid-stride loops are actually exercised:

  mesh ids / coordinates      bit-exact   vs oracle/hdg_oracle.c hdg_c_rectangle_mesh (sequential first-encounter numbering,
                                          itself bit-identical to the numpy oracle: tests/test_host_logic.py)
  colptr / rowval             bit-exact   vs the port's sparse(I,J,V)
  nzval, rhs                  1e-10       (max-norm, relative to the largest entry)
  K_e, b_e of sampled cells   1e-10
  K, b after apply!, meandiag 1e-10 / 1e-13
  u_hat                       backward error of the GPU solution in the ORACLE's system  ||b - K u|| / || |K||u| + |b| || <= 2e-14, and
                              (C2) max-norm distance to the port's own Jacobi-PCG solution (see the tolerance note there)
  sigma_h, u_h of ALL cells   1e-10       vs the port's get_u_sigma! applied to the port's K_e, b_e (so every K_e, b_e is covered)
  err2                        1e-9        vs the port's errornorm
"""
import ctypes as C
import os

import numpy as np
import pytest

import hdg_b200 as hdg
import hdg_oracle as orc
import hdg_oracle_c as occ

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def _threads():
    try:
        return max(len(os.sched_getaffinity(0)), 1)
    except Exception:
        return os.cpu_count() or 1


def _maxrel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def _full_size(order, qd, nx, ny, cpu_pcg):
    nth = _threads()
    LL, UR = (0.0, 0.0), (2.0, 1.0)
    ctx = hdg._Context(order, qd)
    lib = ctx.lib
    hdg.check(lib.hdg_set_rectangle_mesh(ctx.h, nx, ny, LL[0], LL[1], UR[0], UR[1]), ctx.h)
    s = ctx.sizes()
    n, nt, m, t = s.n, s.nt, s.m, s.t
    # ---- mesh: ids and coordinates bit for bit
    mo = occ.rectangle_mesh(nx, ny, LL, UR)
    gm = ctx.download_mesh()
    assert np.array_equal(gm.cells[:, :3], mo.cells) and np.array_equal(gm.cells[:, 3:], mo.cell_faces)
    assert np.array_equal(np.asarray(gm.faces), mo.faces)
    assert np.array_equal(gm.nodes, mo.nodes)
    assert gm.facesets["boundary"] == mo.facesets["boundary"]
    del gm
    # ---- doassemble
    tab = orc.build_tables(order, qd)
    Ko, rhs_o, Ke_o, be_o = occ.doassemble(mo, tab, nthreads=nth)
    hdg.check(lib.hdg_assemble(ctx.h), ctx.h)
    assert s.nnz == Ko.nnz and s.ndof == Ko.shape[0]
    colptr = np.empty(s.ndof + 1, np.int64)
    rowval = np.empty(s.nnz, np.int64)
    hdg.check(lib.hdg_get_pattern(ctx.h, hdg.api.i64p(colptr), hdg.api.i64p(rowval)), ctx.h)
    assert np.array_equal(colptr - 1, Ko.indptr) and np.array_equal(rowval - 1, Ko.indices)      # bit-exact pattern
    del colptr, rowval
    nz = np.empty(s.nnz)
    hdg.check(lib.hdg_get_values(ctx.h, hdg.api.f64p(nz)), ctx.h)
    assert _maxrel(nz, Ko.data) < RTOL
    rhs = np.empty(s.ndof)
    hdg.check(lib.hdg_get_rhs(ctx.h, hdg.api.f64p(rhs)), ctx.h)
    assert _maxrel(rhs, rhs_o) < RTOL
    # ---- K_e, b_e of sampled cells (first / last tile, tile boundaries, random)
    rng = np.random.default_rng(5)
    sample = np.unique(np.r_[0, 1, 31, 32, 33, s.ncell - 33, s.ncell - 32, s.ncell - 1, rng.integers(0, s.ncell, 48)])
    Ke, be = np.empty((m, t), order="F"), np.empty(m)
    for c0 in sample:
        hdg.check(lib.hdg_get_local(ctx.h, int(c0) + 1, hdg.api.f64p(Ke), hdg.api.f64p(be)), ctx.h)
        assert _maxrel(Ke, Ke_o[c0]) < RTOL and _maxrel(be, be_o[c0]) < RTOL
    # ---- apply!
    bf = mo.boundary_faces_sorted()
    dofs = (nt * (bf[:, None] - 1) + np.arange(1, nt + 1)[None, :]).ravel()
    K2, f2, md, dset = occ.apply_dirichlet_homogeneous(Ko, rhs_o, dofs)
    del Ko
    hdg.check(lib.hdg_apply_dirichlet(ctx.h, None), ctx.h)
    mg = C.c_double()
    hdg.check(lib.hdg_get_meandiag(ctx.h, C.byref(mg)), ctx.h)
    assert abs(mg.value - md) < 1e-13 * md
    hdg.check(lib.hdg_get_values(ctx.h, hdg.api.f64p(nz)), ctx.h)
    assert _maxrel(nz, K2.data) < RTOL
    hdg.check(lib.hdg_get_rhs(ctx.h, hdg.api.f64p(rhs)), ctx.h)
    assert _maxrel(rhs, f2) < RTOL
    del nz
    # ---- K \ b: multigrid-PCG on the device; the solution must solve the ORACLE's system
    hdg.check(lib.hdg_set_preconditioner(ctx.h, 2), ctx.h)
    info = hdg.api.SolveInfo()
    hdg.check(lib.hdg_solve(ctx.h, 1e-13, 2000, C.byref(info)), ctx.h)
    assert info.converged and info.iterations <= 80
    x = np.empty(s.ndof)
    hdg.check(lib.hdg_get_trace(ctx.h, hdg.api.f64p(x)), ctx.h)
    # normwise backward error: ||K|| ||x|| is ~1e6 ||b|| here (K ~ 16, x ~ 1, b ~ 4e-5), so ||b - K x|| / ||b|| of ANY computed
    # solution sits at ~1e-10..1e-9 (sparse direct solve 8e-11, the port's PCG 1.1e-9); the meaningful scale is |K||x| + |b|
    # (direct solve: 1.0e-16, the port's PCG at rtol 1e-13: 1.4e-15)
    res = f2 - K2 @ x
    assert np.linalg.norm(res) / np.linalg.norm(abs(K2) @ np.abs(x) + np.abs(f2)) < 2e-14
    if cpu_pcg:
        # the port's own Jacobi-PCG (same sign-fixed system, same stopping rule, rtol 1e-13).  The condition number of this
        # system is ~h^-2 ~ 1e6, so unit roundoff alone moves ANY computed solution by ~1e-10 (the port's PCG is 3.7e-12 away
        # from a sparse direct solve, the multigrid-PCG solution 1.8e-10 from the port's): the 1e-10 bar of the north-star is
        # checked against the direct solve on the small meshes (test_gpu_parity.py), here the bar is 10 x cond x eps
        xo, it, rel = occ.pcg(K2, f2, dset, rtol=1e-13, maxit=100000, nthreads=nth)
        assert rel <= 1e-13
        assert _maxrel(x, xo) < 1e-9
    # ---- get_u_sigma! for every cell: the port's recovery on the port's K_e, b_e with the same trace
    hdg.check(lib.hdg_recover(ctx.h), ctx.h)
    sig = np.empty((2 * n, s.ncell))
    u = np.empty((n, s.ncell))
    hdg.check(lib.hdg_get_mvalues(ctx.h, hdg.api.f64p(sig), hdg.api.f64p(u), None), ctx.h)      # column-major ncell x nb
    sig_o, u_o = occ.recover(mo, tab, x, Ke_o, be_o, nthreads=nth)
    assert _maxrel(sig.T, sig_o) < RTOL and _maxrel(u.T, u_o) < RTOL
    # ---- errornorm
    e2 = C.c_double()
    hdg.check(lib.hdg_errornorm(ctx.h, 1, C.byref(e2)), ctx.h)
    e2_o = occ.errornorm(mo, tab, u_o, nthreads=nth)
    # (u_h - u_ex)^2 of differences ~ sqrt(err2) carries an absolute rounding error ~ eps |u| sqrt(err2): 1.5e-7 relative at err2 = 2.6e-19 (C3)
    assert abs(e2.value - e2_o) < 1e-9 * e2_o + 1e-15 * np.sqrt(e2_o)
    ctx.close()
    return info.iterations, e2.value


def test_c2_full_size_against_the_oracle_port():
    """BASELINE config C2: k=1, 1000 x 500 = 1M elements, 3 003 000 trace dofs, 30 006 000 stored entries."""
    it, e2 = _full_size(1, 2, 1000, 500, cpu_pcg=True)
    assert 1e-11 < e2 < 3e-11            # discretisation error of this mesh (h^4): 1.83e-11


def test_c3_per_gpu_size_against_the_oracle_port():
    """BASELINE config C3 at its single-GPU size: k=2, 2000 x 1000 = 4M elements, 18 009 000 trace dofs, 270 M stored entries
    (the element_quad_kernel<2> path: 125 000 tiles, 64-bit offsets into the 5.8 GB [K_e | b_e] array)."""
    _full_size(2, 4, 2000, 1000, cpu_pcg=False)
