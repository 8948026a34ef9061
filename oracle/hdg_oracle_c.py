"""ctypes wrapper of oracle/hdg_oracle.c (C restatement of the reference's CPU path).

TEST INFRASTRUCTURE / CPU BASELINE ONLY - see the header of hdg_oracle.c.  Tables come from the
numpy oracle (hdg_oracle.build_tables)."""
import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libhdg_oracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.check_call(["make", "-C", _HERE])
        _lib = C.CDLL(_SO)
        _lib.hdg_c_doassemble.restype = C.c_int
        _lib.hdg_c_pcg.restype = C.c_int
        _lib.hdg_c_max_threads.restype = C.c_int
        _lib.hdg_c_rectangle_mesh.restype = C.c_int64
        _lib.hdg_c_errornorm.restype = C.c_double
        _lib.hdg_c_recover.restype = None
    return _lib


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def max_threads():
    return load().hdg_c_max_threads()


def doassemble(mesh, tab, tau=1.0, fq=None, nthreads=1, keep_local=True):
    """C doassemble on an oracle Mesh + Tables.  Returns (K csc, rhs, K_e (ncell,m,t), b_e (ncell,m))."""
    lib = load()
    n, nt, nq, nfq = tab.n, tab.nt, tab.nq, tab.nfq
    m, t = 3 * n, 3 * nt
    nc, nf = mesh.ncells, mesh.nfaces
    ndof = nf * nt
    cells = np.ascontiguousarray(mesh.cells, dtype=np.int64)
    cfaces = np.ascontiguousarray(mesh.cell_faces, dtype=np.int64)
    nodes = np.ascontiguousarray(mesh.nodes, dtype=np.float64)
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (tab.N, tab.dN, tab.E, tab.T, tab.M, tab.qw, tab.fw)]
    Ke = np.empty((nc, m, t)) if keep_local else None
    be = np.empty((nc, m)) if keep_local else None
    rhs = np.empty(ndof)
    colptr = np.empty(ndof + 1, np.int64)
    rowval = np.empty(nc * t * t, np.int64)
    nzval = np.empty(nc * t * t)
    nnz = C.c_int64()
    fqa = np.ascontiguousarray(fq, dtype=np.float64) if fq is not None else None
    st = lib.hdg_c_doassemble(C.c_int(n), C.c_int(nt), C.c_int(nq), C.c_int(nfq), *[_p(a) for a in arrs],
                              C.c_int64(nc), C.c_int64(nf), _p(cells, C.c_int64), _p(cfaces, C.c_int64), _p(nodes),
                              C.c_double(tau), _p(fqa), C.c_int(nthreads), _p(Ke), _p(be), _p(rhs),
                              _p(colptr, C.c_int64), _p(rowval, C.c_int64), _p(nzval), C.byref(nnz))
    if st != 0:
        raise RuntimeError(f"hdg_c_doassemble failed with status {st}")
    K = sp.csc_matrix((nzval[:nnz.value], rowval[:nnz.value], colptr), shape=(ndof, ndof))
    return K, rhs, Ke, be


def pcg(K, b, isbc, rtol=1e-12, maxit=100000, nthreads=1):
    lib = load()
    K = K.tocsc()
    n = K.shape[0]
    x = np.empty(n)
    rel = C.c_double()
    colptr = K.indptr.astype(np.int64)
    rowval = K.indices.astype(np.int64)
    data = np.ascontiguousarray(K.data, dtype=np.float64)
    bb = np.ascontiguousarray(b, dtype=np.float64)
    mask = np.ascontiguousarray(isbc, dtype=np.uint8)
    it = lib.hdg_c_pcg(C.c_int64(n), _p(colptr, C.c_int64), _p(rowval, C.c_int64), _p(data), _p(bb),
                       _p(mask, C.c_uint8), C.c_double(rtol), C.c_int(maxit), C.c_int(nthreads), _p(x), C.byref(rel))
    return x, it, rel.value


def rectangle_mesh(nx, ny, LL=(0.0, 0.0), UR=(1.0, 1.0)):
    """C restatement of rectangle_mesh (src/generate_mesh.jl:101-143) for sizes where the numpy oracle's Python loops take
    minutes; same outputs as hdg_oracle.rectangle_mesh except that only the "boundary" face set is built (faces with one cell,
    src/generate_mesh.jl:60-89).  Checked against the numpy oracle in tests/test_host_logic.py."""
    import hdg_oracle as orc
    lib = load()
    nc, nn, nf = 2 * nx * ny, (nx + 1) * (ny + 1), 3 * nx * ny + nx + ny
    cells = np.empty((nc, 3), np.int64)
    cfaces = np.empty((nc, 3), np.int64)
    nodes = np.empty((nn, 2))
    faces = np.zeros((nf, 4), np.int64)
    got = lib.hdg_c_rectangle_mesh(C.c_int64(nx), C.c_int64(ny), C.c_double(LL[0]), C.c_double(LL[1]), C.c_double(UR[0]),
                                   C.c_double(UR[1]), _p(cells, C.c_int64), _p(cfaces, C.c_int64), _p(nodes), _p(faces, C.c_int64))
    assert got == nf, (got, nf)
    bnd = np.flatnonzero(faces[:, 3] == 0) + 1
    return orc.Mesh(cells, cfaces, nodes, faces, {"boundary": set(bnd.tolist())})


def recover(mesh, tab, uhat, Ke, be, nthreads=1):
    """get_u_sigma! (examples/poisson2D_HDG.jl:197-212) in C: returns sigma (ncell,2n), u (ncell,n)."""
    lib = load()
    n, nt = tab.n, tab.nt
    nc = mesh.ncells
    sig = np.empty((nc, 2 * n))
    u = np.empty((nc, n))
    cf = np.ascontiguousarray(mesh.cell_faces, dtype=np.int64)
    lib.hdg_c_recover(C.c_int(n), C.c_int(nt), C.c_int64(nc), _p(cf, C.c_int64), _p(np.ascontiguousarray(Ke)),
                      _p(np.ascontiguousarray(be)), _p(np.ascontiguousarray(uhat, dtype=np.float64)), C.c_int(nthreads), _p(sig), _p(u))
    return sig, u


def errornorm(mesh, tab, u_vals, nthreads=1):
    """errornorm(u_h, u_ex = sin pi x sin pi y), src/DiscreteFunctions.jl:97-120, in C."""
    lib = load()
    cells = np.ascontiguousarray(mesh.cells, dtype=np.int64)
    nodes = np.ascontiguousarray(mesh.nodes, dtype=np.float64)
    return lib.hdg_c_errornorm(C.c_int(tab.n), C.c_int(tab.nq), _p(np.ascontiguousarray(tab.N, dtype=np.float64)),
                               _p(np.ascontiguousarray(tab.M, dtype=np.float64)), _p(np.ascontiguousarray(tab.qw, dtype=np.float64)),
                               C.c_int64(mesh.ncells), _p(cells, C.c_int64), _p(nodes),
                               _p(np.ascontiguousarray(u_vals, dtype=np.float64)), C.c_int(nthreads))


def apply_dirichlet_homogeneous(K, rhs, dofs):
    """apply!(K,f,dbc) for g = 0 (src/boundary.jl:121-158), vectorised for the full-size parity tests: meandiag over all dofs
    before modification, prescribed rows and columns zeroed in place (pattern unchanged), K[d,d] = m, f[d] = 0.  Checked
    against hdg_oracle.apply_dirichlet in tests/test_host_logic.py.  Returns (K, rhs, m, dirichlet mask)."""
    K = K.copy().tocsc()
    rhs = rhs.copy()
    m = float(np.abs(K.diagonal()).mean())
    dset = np.zeros(K.shape[0], bool)
    dset[np.asarray(dofs) - 1] = True
    col_of = np.repeat(np.arange(K.shape[1]), np.diff(K.indptr))
    K.data[dset[col_of] | dset[K.indices]] = 0.0
    K.data[(K.indices == col_of) & dset[col_of]] = m
    rhs[dset] = 0.0
    return K, rhs, m, dset
