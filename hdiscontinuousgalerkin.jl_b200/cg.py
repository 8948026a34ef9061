"""Continuous-Galerkin side of the reference's exported API: DofHandler, sparsity pattern,
Dirichlet conditions on a DofHandler, reconstruct!, and the mesh matrix helpers used for plotting.

Two layers.  (1) Host-side integer logic (`DofHandler`, `create_sparsity_pattern`, `DirichletCG`, `reconstruct_`): the
sequential dictionary walk of the reference (src/dofhandler.jl:84-152) as a sort-based first-encounter ranking; identical to
the reference's goldens (test/test_handlers.jl:13-19) and to the loop-faithful oracle (tests/test_host_logic.py).
(2) The device path (`poisson2D_CG`, `CGDevice`): examples/poisson2D_CG.jl through the C ABI (hdg_cg_* in
include/hdg_b200.h, csrc/hdg_cg.cu) - dof numbering, sparsity pattern, element matrices + assemble!, Dirichlet + apply!, the
solve and errornorm all run on the GPU; there is no CPU fallback for it."""
from __future__ import annotations

import numpy as np

import ctypes as C

from ._lib import SolveInfo, check, f64p, i64p
from .api import PolygonalMesh, RefTetrahedron, _Context


class ContinuousLagrange:
    """ContinuousLagrange{dim,shape,order}, src/LagrangeFE.jl:1-31 (nodal basis; topology from get_nodal_points,
    src/shapes.jl:46-57)."""

    def __init__(self, dim, shape, order):
        if dim != 2 or shape is not RefTetrahedron:
            raise NotImplementedError("ContinuousLagrange{2,RefTetrahedron,k} only")
        if order < 1:
            raise ValueError("order >= 1")
        self.dim, self.shape, self.order = dim, shape, order
        self.topology = {0: 3, 1: 3 * (order - 1), 2: (order - 1) * (order - 2) // 2}

    def getnbasefunctions(self):
        return (self.order + 1) * (self.order + 2) // 2


class LagrangeField:
    """What DofHandler needs from a TrialFunction on a CG space: the element and the number of components."""

    def __init__(self, felem: ContinuousLagrange, mesh: PolygonalMesh, ncomponents=1):
        self.fe, self.mesh, self.ncomponents = felem, mesh, ncomponents
        self.m_values = np.zeros((mesh.getncells(), felem.getnbasefunctions() * ncomponents))

    def getnlocaldofs(self):
        return self.fe.getnbasefunctions() * self.ncomponents


def _first_encounter_rank(keys):
    """rank (0-based) of every key by the position of its first occurrence in `keys`."""
    uniq, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty(uniq.size, np.int64)
    rank[order] = np.arange(uniq.size)
    return rank[inv], first[order]


class DofHandler:
    """DofHandler([u_h], mesh), src/dofhandler.jl:1-19,84-152: `cell_dofs`, `cell_dofs_offset` (1-based)."""

    def __init__(self, variables, mesh: PolygonalMesh):
        if len(variables) != 1:
            raise NotImplementedError("one field (the reference's examples and tests use one)")
        self.variables, self.mesh = list(variables), mesh
        u = variables[0]
        k, ncomp = u.fe.order, u.ncomponents
        if k > 2:
            raise NotImplementedError("orders 1 and 2 (for order >= 3 the reference re-uses only the first edge dof, "
                                      "src/dofhandler.jl:121-124; see oracle/hdg_oracle.py distribute_dofs)")
        nc = mesh.getncells()
        nn = mesh.getnnodes()
        # entities in the order a cell meets them: its 3 vertices, then (order 2) its 3 faces
        ent = mesh.cells[:, :3].astype(np.int64)
        if k == 2:
            ent = np.hstack([ent, nn + mesh.cells[:, 3:6].astype(np.int64)])
        per_cell = ent.shape[1]
        rank, _ = _first_encounter_rank(ent.reshape(-1))
        # every entity carries one node (ncomp dofs); cells carry no interior dofs for k <= 2, so dof = ncomp*rank + d
        base = ncomp * rank.reshape(nc, per_cell)
        dofs = (base[:, :, None] + np.arange(ncomp)[None, None, :] + 1).reshape(nc, per_cell * ncomp)
        self.cell_dofs = dofs.reshape(-1)
        self.cell_dofs_offset = 1 + per_cell * ncomp * np.arange(nc + 1, dtype=np.int64)

    def ndofs(self):
        return int(self.cell_dofs.max())                                   # src/dofhandler.jl:41

    def ndofs_per_cell(self, cell=1):
        return int(self.cell_dofs_offset[cell] - self.cell_dofs_offset[cell - 1])   # :42

    def celldofs(self, i):
        """celldofs!(global_dofs, dh, i), :154-158 (i is 1-based)."""
        a = self.cell_dofs_offset[i - 1] - 1
        return self.cell_dofs[a:a + self.ndofs_per_cell(i)].copy()

    def dof_range(self, field):
        """dof_range(dh, u), :76-81: 1-based inclusive range as a Python range."""
        j = self.variables.index(field)
        off = sum(v.getnlocaldofs() for v in self.variables[:j])
        return range(off + 1, off + field.getnlocaldofs() + 1)


def ndofs(dh):
    return dh.ndofs()


def ndofs_per_cell(dh, cell=1):
    return dh.ndofs_per_cell(cell)


def dof_range(dh, field):
    return dh.dof_range(field)


def create_sparsity_pattern(dh: DofHandler):
    """create_sparsity_pattern(dh), src/dofhandler.jl:160-216: (colptr, rowval), 1-based Int64, rows ascending per column
    - the pattern of sparse(I, J, zeros) over all element couplings plus the diagonal."""
    n = dh.ndofs_per_cell()
    nd = dh.ndofs()
    g = dh.cell_dofs.reshape(-1, n) - 1
    I = np.concatenate([np.repeat(g[:, None, :], n, axis=1).reshape(-1), np.arange(nd)])     # I[e, j, i] = g[e, i]
    J = np.concatenate([np.repeat(g[:, :, None], n, axis=2).reshape(-1), np.arange(nd)])     # J[e, j, i] = g[e, j]
    key = np.unique(J * nd + I)
    col, row = key // nd, key % nd
    colptr = np.zeros(nd + 1, np.int64)
    np.cumsum(np.bincount(col, minlength=nd), out=colptr[1:])
    return colptr + 1, row + 1


class DirichletCG:
    """Dirichlet(u, dh, faceset, f), src/boundary.jl:44-96 for a nodal basis: sorted `prescribed_dofs` and `values`
    (f evaluated at the dof's node: vertex, or edge midpoint for order 2 - spatial_nodal_coordinate)."""
    _EDGE_NODES = ((2, 3), (3, 1), (1, 2))        # reference_edge_nodes, src/shapes.jl:24

    def __init__(self, u, dh: DofHandler, faceset, f):
        mesh = dh.mesh
        fset = sorted(mesh.getfaceset(faceset)) if isinstance(faceset, str) else sorted(faceset)
        k = u.fe.order
        ncomp = u.ncomponents
        if ncomp != 1:
            raise NotImplementedError("scalar fields")
        dofs, vals, seen = [], [], set()
        for fi in fset:                              # ascending face id, like the reference's 1:getnfaces loop
            if mesh.faces[fi - 1, 3] != 0:
                raise AssertionError(f"Face {fi} is not in boundary")          # :66
            cell = int(mesh.faces[fi - 1, 2])
            lidx = list(mesh.cells[cell - 1, 3:6]).index(fi)
            x = mesh.nodes[mesh.cells[cell - 1, :3] - 1]
            cd = dh.celldofs(cell)
            a, b = self._EDGE_NODES[lidx]
            cand = [(cd[a - 1], x[a - 1]), (cd[b - 1], x[b - 1])]
            if k == 2:
                cand.append((cd[3 + lidx], 0.5 * (x[a - 1] + x[b - 1])))
            for d, xd in cand:
                if int(d) not in seen:
                    seen.add(int(d))
                    dofs.append(int(d))
                    vals.append(float(f(xd)) if callable(f) else float(f[0]))
        p = np.argsort(np.array(dofs, dtype=np.int64), kind="stable")           # :94-95
        self.prescribed_dofs = np.array(dofs, dtype=np.int64)[p]
        self.values = np.array(vals, dtype=np.float64)[p]


def reconstruct_(field, u, dh: DofHandler):
    """reconstruct!(field, u, dh), src/dofhandler.jl:219-226: m_values[cell,:] = u[cell dofs of the field]."""
    n = dh.ndofs_per_cell()
    field.m_values[:, :] = np.asarray(u)[dh.cell_dofs.reshape(-1, n) - 1][:, [i - 1 for i in dh.dof_range(field)]]


def get_vertices_matrix(mesh: PolygonalMesh):
    """get_vertices_matrix(mesh), src/mesh.jl:56-62: nnode x 2."""
    return np.array(mesh.nodes, dtype=np.float64, copy=True)


def getcells_matrix(mesh: PolygonalMesh):
    """getcells_matrix(mesh), src/mesh.jl:63-69: ncell x 3 node ids (1-based)."""
    return np.array(mesh.cells[:, :3], dtype=np.int64, copy=True)


# ----------------------------------------------------------------------------------------------
# device path: examples/poisson2D_CG.jl through the C ABI
# ----------------------------------------------------------------------------------------------
class CGDevice:
    """DofHandler + create_sparsity_pattern + doassemble + apply! + K \\ b + errornorm of examples/poisson2D_CG.jl on the GPU
    (hdg_cg_*).  `mesh` is handed over like in the HDG driver; `order` is that of ContinuousLagrange{2,RefTetrahedron,order}."""

    def __init__(self, mesh: PolygonalMesh, order=1, device=-1):
        self.ctx = _Context(1, 2, 1.0, 1, device)        # the HDG tables of the context are not used by the CG side
        self.ctx.set_mesh(mesh)
        self.mesh, self.order = mesh, order
        n = C.c_int64()
        check(self.ctx.lib.hdg_cg_setup(self.ctx.h, int(order), C.byref(n)), self.ctx.h)
        sz = np.zeros(4, np.int64)
        check(self.ctx.lib.hdg_cg_get_sizes(self.ctx.h, i64p(sz)), self.ctx.h)
        self.ndofs, self.nnz, self.ndofs_per_cell, self.ncell = (int(v) for v in sz)

    def dofhandler(self):
        """(cell_dofs (ncell, ndofs_per_cell), colptr, rowval) - 1-based Int64 like dh.cell_dofs and the SparseMatrixCSC."""
        cd = np.empty((self.ncell, self.ndofs_per_cell), np.int64)
        cp = np.empty(self.ndofs + 1, np.int64)
        rv = np.empty(self.nnz, np.int64)
        check(self.ctx.lib.hdg_cg_get_dofhandler(self.ctx.h, i64p(cd), i64p(cp), i64p(rv)), self.ctx.h)
        return cd, cp, rv

    def doassemble(self):
        check(self.ctx.lib.hdg_cg_assemble(self.ctx.h), self.ctx.h)
        return self.system()

    def system(self):
        import scipy.sparse as sp
        cd, cp, rv = self.dofhandler()
        nz, b = np.empty(self.nnz), np.empty(self.ndofs)
        check(self.ctx.lib.hdg_cg_get_system(self.ctx.h, f64p(nz), f64p(b), None), self.ctx.h)
        return sp.csc_matrix((nz, rv - 1, cp - 1), shape=(self.ndofs, self.ndofs)), b

    def apply_(self):
        check(self.ctx.lib.hdg_cg_apply_dirichlet(self.ctx.h), self.ctx.h)
        m = C.c_double()
        check(self.ctx.lib.hdg_cg_get_meandiag(self.ctx.h, C.byref(m)), self.ctx.h)
        return m.value

    def solve(self, rtol=1e-13, maxit=100000):
        info = SolveInfo()
        check(self.ctx.lib.hdg_cg_solve(self.ctx.h, float(rtol), int(maxit), C.byref(info)), self.ctx.h)
        u = np.empty(self.ndofs)
        check(self.ctx.lib.hdg_cg_get_system(self.ctx.h, None, None, f64p(u)), self.ctx.h)
        return u, dict(iterations=info.iterations, converged=bool(info.converged), relres=info.relres, solve_ms=info.solve_ms)

    def errornorm(self):
        e = C.c_double()
        check(self.ctx.lib.hdg_cg_errornorm(self.ctx.h, C.byref(e)), self.ctx.h)
        return e.value

    def close(self):
        self.ctx.close()


def poisson2D_CG(mesh=None, order=1, rtol=1e-13):
    """examples/poisson2D_CG.jl end to end on the device.  Returns a dict (K and b before apply!, u, err2, info)."""
    from .api import TriangleCell, rectangle_mesh
    if mesh is None:
        mesh = rectangle_mesh(TriangleCell, (10, 10), (0.0, 0.0), (1.0, 1.0))
    dev = CGDevice(mesh, order)
    try:
        K, b = dev.doassemble()
        m = dev.apply_()
        u, info = dev.solve(rtol)
        return dict(K=K, b=b, u=u, err2=dev.errornorm(), info=info, meandiag=m, ndofs=dev.ndofs, dofhandler=dev.dofhandler())
    finally:
        dev.close()
