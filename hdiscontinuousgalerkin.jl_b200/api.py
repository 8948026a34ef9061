"""Host-side mirror of the reference's Julia API for the HDG Poisson path.

Julia is not available in this image, so the host layer above the C ABI is Python; names,
argument meaning and error behaviour follow the reference (src/HDiscontinuousGalerkin.jl:38-108
exports and the script-level functions of examples/poisson2D_HDG.jl) so that
`examples/poisson2D_hdg.py` and the parity tests read like the reference's own driver/tests.
All numerical work happens in libhdg_b200.so on the GPU - the objects here are thin handles
around one `hdg_context`.  The same calls expressed as Julia `ccall`s are in
julia/HDGB200.jl / INTEGRATION.md.
"""
from __future__ import annotations

import ctypes as C
import re

import numpy as np

from . import _lib
from ._lib import BadGeometryError, HDGError, Params, Sizes, SolveInfo, check, f64p, i64p

# ----------------------------------------------------------------------------------------------
# element / basis descriptors (src/basis.jl, src/FiniteElement.jl) - tables are built by the library
# ----------------------------------------------------------------------------------------------
TriangleCell = "TriangleCell"      # Cell{2,3,3}, src/mesh.jl:24
RefTetrahedron = "RefTetrahedron"  # src/shapes.jl:4


class Dubiner:
    """Dubiner{dim,RefTetrahedron,order}, src/basis.jl:52."""

    def __init__(self, dim=2, shape=RefTetrahedron, order=1):
        if dim != 2:
            raise ValueError("Dubiner basis is defined on the triangle (dim = 2)")
        self.dim, self.shape, self.order = dim, shape, int(order)

    def getnbasefunctions(self):
        return (self.order + 1) * (self.order + 2) // 2


class Legendre:
    """Legendre{dim,RefTetrahedron,order}, src/basis.jl:339."""

    def __init__(self, dim=1, shape=RefTetrahedron, order=1):
        self.dim, self.shape, self.order = dim, shape, int(order)

    def getnbasefunctions(self):
        return self.order + 1


def value(ip, j, xi):
    """value(ip, j, xi), src/basis.jl:65-86 (Dubiner) / :351-354 (Legendre): evaluated by the library's table builder."""
    x = np.atleast_1d(np.asarray(xi, dtype=np.float64)).copy()
    v = C.c_double()
    check(_lib.load().hdg_basis_value(0 if isinstance(ip, Dubiner) else 1, int(j), f64p(x), C.byref(v), None))
    return v.value


def gradient_value(ip, j, xi):
    """gradient_value(ip, j, xi), src/basis.jl:88-110."""
    x = np.atleast_1d(np.asarray(xi, dtype=np.float64)).copy()
    v = C.c_double()
    g = np.zeros(2)
    check(_lib.load().hdg_basis_value(0 if isinstance(ip, Dubiner) else 1, int(j), f64p(x), C.byref(v), f64p(g)))
    return g if isinstance(ip, Dubiner) else g[:1]


class GenericFiniteElement:
    """GenericFiniteElement(func_basis), src/FiniteElement.jl:8-27 (order-1 Lagrange geometry)."""

    def __init__(self, basis):
        self.basis = basis
        self.order = basis.order


# ----------------------------------------------------------------------------------------------
# mesh (src/mesh.jl, src/generate_mesh.jl, src/triangle_mesh.jl)
# ----------------------------------------------------------------------------------------------
class PolygonalMesh:
    """PolygonalMesh{2,3,3,2,Float64}, src/mesh.jl:43-49, in the Julia memory layouts the C ABI takes.

    cells : (ncell,6) int64 C-order  = Vector{Cell{2,3,3}} (3 node ids then 3 face ids, 1-based)
    nodes : (nnode,2) float64        = Vector{Node{2,Float64}}
    faces : (nface,4) int64 F-order  = Matrix{Int}  (v1 v2 cell1 cell2|0)
    """

    def __init__(self, cells, nodes, faces, facesets, rect=None):
        self.cells = np.ascontiguousarray(cells, dtype=np.int64)
        self.nodes = np.ascontiguousarray(nodes, dtype=np.float64)
        self.faces = np.asfortranarray(faces, dtype=np.int64)
        self.facesets = facesets
        self._rect = rect   # (nx, ny, LL, UR) when produced by rectangle_mesh

    # queries, src/mesh.jl:70-76
    def getncells(self):
        return self.cells.shape[0]

    def getnfaces(self):
        return self.faces.shape[0]

    def getnnodes(self):
        return self.nodes.shape[0]

    def getfaceset(self, name):
        return self.facesets[name]

    def getcells_matrix(self):      # src/mesh.jl:63-69
        return self.cells[:, :3].copy()

    def get_vertices_matrix(self):  # src/mesh.jl:56-62
        return self.nodes.copy()


def getncells(mesh):
    return mesh.getncells()


def getnfaces(mesh):
    return mesh.getnfaces()


def getnnodes(mesh):
    return mesh.getnnodes()


def getfaceset(mesh, name):
    return mesh.getfaceset(name)


# small mesh queries of src/mesh.jl (host-side, 1-based indices like the reference)
def n_faces_per_cell(mesh):
    return 3                                                    # src/mesh.jl:70


def n_nodes_per_cell(mesh):
    return 3                                                    # src/mesh.jl:71


def reference_edge_nodes():
    return ((2, 3), (3, 1), (1, 2))                             # src/shapes.jl:24


def getnodeset(mesh, name):
    """getnodeset(mesh, name), src/mesh.jl:76: the nodes of the faces of the face set of that name
    (_get_rectangular_boundary_sets, src/generate_mesh.jl:60-89, builds both from the same faces)."""
    f = np.array(sorted(mesh.facesets[name]), dtype=np.int64) - 1
    return set(np.unique(mesh.faces[f, :2]).tolist())


def getnodesets(mesh):
    return {k: getnodeset(mesh, k) for k in mesh.facesets}


def get_coordinates(item, mesh):
    """get_coordinates(cell, mesh) / get_coordinates(face::Int, mesh), src/mesh.jl:83-119: vertex coordinates of a cell
    (pass the 1-based cell index as ("cell", i), or a row of mesh.cells) or of a face (1-based face index)."""
    if isinstance(item, tuple) and item[0] == "cell":
        return mesh.nodes[mesh.cells[item[1] - 1, :3] - 1].copy()
    if np.ndim(item) == 0:
        return mesh.nodes[mesh.faces[int(item) - 1, :2] - 1].copy()
    return mesh.nodes[np.asarray(item)[:3] - 1].copy()


def get_cell_coordinates(cell_idx, mesh):
    return mesh.nodes[mesh.cells[cell_idx - 1, :3] - 1].copy()   # src/mesh.jl:101-107


def volume(coords):
    """volume(verts) of a polygon given by its vertices (shoelace formula, src/shapes.jl / src/mesh.jl:148-152)."""
    x = np.asarray(coords, dtype=np.float64)
    xn = np.roll(x, -1, axis=0)
    return 0.5 * abs(np.sum(x[:, 0] * xn[:, 1] - xn[:, 0] * x[:, 1]))


def cell_volume(mesh, cell_idx):
    return volume(get_cell_coordinates(cell_idx, mesh))         # src/mesh.jl:148-152


def cell_centroid(mesh, cell_idx):
    """src/mesh.jl:154-161."""
    x = get_cell_coordinates(cell_idx, mesh)
    xn = np.roll(x, -1, axis=0)
    cr = x[:, 0] * xn[:, 1] - xn[:, 0] * x[:, 1]
    ve = cell_volume(mesh, cell_idx)
    return np.array([np.sum(cr * (x[:, 0] + xn[:, 0])), np.sum(cr * (x[:, 1] + xn[:, 1]))]) / (6.0 * ve)


def cell_diameter(mesh, cell_idx):
    """cell_diameter(mesh, idx), src/mesh.jl:121-129: the longest face of the cell."""
    h = 0.0
    for f in mesh.cells[cell_idx - 1, 3:6]:
        a, b = mesh.nodes[mesh.faces[f - 1, 0] - 1], mesh.nodes[mesh.faces[f - 1, 1] - 1]
        h = max(h, float(np.linalg.norm(b - a)))
    return h


def face_orientation(mesh, cell_idx, face_idx):
    """src/mesh.jl:51-54 (1-based cell / local face)."""
    k1, k2 = ((2, 3), (3, 1), (1, 2))[face_idx - 1]
    nd = mesh.cells[cell_idx - 1]
    return bool(nd[k2 - 1] > nd[k1 - 1])


def rectangle_mesh(celltype, nel, LL, UR):
    """rectangle_mesh(TriangleCell,(nx,ny),LL,UR), src/generate_mesh.jl:101-143.

    Generated on the GPU (node coordinates with the reference's arithmetic, first-encounter face
    numbering in closed form) and downloaded into the Julia layouts."""
    if celltype != TriangleCell:
        raise NotImplementedError("only TriangleCell meshes are on the HDG path (RectangleCell is VEM-only)")
    nx, ny = int(nel[0]), int(nel[1])
    ctx = _Context(order=1)
    try:
        check(_lib.load().hdg_set_rectangle_mesh(ctx.h, nx, ny, float(LL[0]), float(LL[1]), float(UR[0]), float(UR[1])), ctx.h)
        mesh = ctx.download_mesh()
    finally:
        ctx.close()
    mesh._rect = (nx, ny, (float(LL[0]), float(LL[1])), (float(UR[0]), float(UR[1])))
    # named edge sets by exact coordinate compare, src/generate_mesh.jl:60-89
    b = np.array(sorted(mesh.facesets["boundary"]), dtype=np.int64) - 1
    a, c = mesh.nodes[mesh.faces[b, 0] - 1], mesh.nodes[mesh.faces[b, 1] - 1]
    sets = {"bottom": (a[:, 1] == LL[1]) & (c[:, 1] == LL[1])}
    sets["right"] = ~sets["bottom"] & (a[:, 0] == UR[0]) & (c[:, 0] == UR[0])
    sets["top"] = ~sets["bottom"] & ~sets["right"] & (a[:, 1] == UR[1]) & (c[:, 1] == UR[1])
    sets["left"] = ~sets["bottom"] & ~sets["right"] & ~sets["top"] & (a[:, 0] == LL[0]) & (c[:, 0] == LL[0])
    for k, msk in sets.items():
        mesh.facesets[k] = set((b[msk] + 1).tolist())
    return mesh


_NUM = re.compile(r"\b((\d*\.)?\d+)\b")


def _triangle_rows(path):
    rows, first = [], True
    with open(path) as fh:
        for ln in fh:
            if re.match(r"^\s*(?:#|$)", ln):
                continue
            if first:           # header row
                first = False
                continue
            rows.append([m.group(0) for m in _NUM.finditer(ln)])
    return rows


def number_faces(tri_nodes):
    """First-encounter face numbering of a triangle list (src/generate_mesh.jl:20-46,
    src/triangle_mesh.jl:66-101) without the sequential hash table: sort the 3*ncell edge keys,
    take the first encounter position of every distinct edge, rank the distinct edges by that
    position.  tri_nodes: (ncell,3) 1-based CCW.  Returns (cell_faces (ncell,3), faces (nface,4))."""
    tri = np.asarray(tri_nodes, dtype=np.int64)
    nc = tri.shape[0]
    k1 = np.array([1, 2, 0])
    k2 = np.array([2, 0, 1])
    v1 = tri[:, k1].reshape(-1)          # encounter position p = 3*cell + local face
    v2 = tri[:, k2].reshape(-1)
    lo, hi = np.minimum(v1, v2), np.maximum(v1, v2)
    key = lo * (tri.max() + 1) + hi
    order = np.argsort(key, kind="stable")          # stable: encounter positions ascend inside a group
    sk = key[order]
    start = np.ones(sk.size, bool)
    start[1:] = sk[1:] != sk[:-1]
    end = np.ones(sk.size, bool)
    end[:-1] = start[1:]
    grp = np.cumsum(start) - 1                       # group (distinct edge) of every sorted slot
    first_pos, last_pos = order[start], order[end]   # first / last encounter of every distinct edge
    if np.any(np.bincount(grp) > 2):
        raise ValueError("non-manifold mesh: an edge is shared by more than two cells")
    by_first = np.argsort(first_pos, kind="stable")  # face id = rank of the first encounter
    rank = np.empty(first_pos.size, np.int64)
    rank[by_first] = np.arange(first_pos.size)
    face_of_pos = np.empty(sk.size, np.int64)
    face_of_pos[order] = rank[grp]
    cell_faces = (face_of_pos + 1).reshape(nc, 3)
    fp, lp = first_pos[by_first], last_pos[by_first]
    faces = np.zeros((fp.size, 4), dtype=np.int64)
    faces[:, 0] = v1[fp]                             # (v1,v2) in the first cell's local direction
    faces[:, 1] = v2[fp]
    faces[:, 2] = fp // 3 + 1
    faces[:, 3] = np.where(lp != fp, lp // 3 + 1, 0)
    return cell_faces, faces


def number_faces_gpu(tri_nodes, nodes):
    """First-encounter face numbering on the device (hdg_number_faces): same result as `number_faces`, plus the
    counter-clockwise fix of _check_node_data.  Returns (cells (ncell,6), faces (nface,4) F-order)."""
    tri = np.ascontiguousarray(tri_nodes, dtype=np.int64)
    xy = np.ascontiguousarray(nodes, dtype=np.float64)
    ctx = _Context(order=1)
    try:
        nface = C.c_int64()
        check(ctx.lib.hdg_number_faces(ctx.h, i64p(tri), tri.shape[0], f64p(xy), xy.shape[0], None, None, 0, C.byref(nface)), ctx.h)
        cells = np.empty((tri.shape[0], 6), np.int64)
        faces = np.empty((nface.value, 4), np.int64, order="F")
        check(ctx.lib.hdg_number_faces(ctx.h, i64p(tri), tri.shape[0], f64p(xy), xy.shape[0], i64p(cells), i64p(faces),
                                       nface.value, C.byref(nface)), ctx.h)
    finally:
        ctx.close()
    return cells, faces


def order_cells_gpu(tri_nodes, nodes):
    """Morton (Z-order) permutation of the cells by centroid, computed and sorted on the device (hdg_order_cells).
    Returns perm (0-based): the cell that comes i-th along the curve is input cell perm[i]."""
    tri = np.ascontiguousarray(np.asarray(tri_nodes)[:, :3], dtype=np.int64)
    xy = np.ascontiguousarray(nodes, dtype=np.float64)
    perm = np.empty(tri.shape[0], np.int64)
    ctx = _Context(order=1)
    try:
        check(ctx.lib.hdg_order_cells(ctx.h, i64p(tri), tri.shape[0], f64p(xy), xy.shape[0], i64p(perm)), ctx.h)
    finally:
        ctx.close()
    return perm - 1


class RenumberedMesh:
    """A mesh with its cells reordered (and its faces renumbered by the reference's first-encounter rule for that cell order,
    src/triangle_mesh.jl:66-101) plus the maps back to the caller's numbering.

    mesh      : the renumbered PolygonalMesh (hand this to doassemble / hdg_set_mesh)
    cell_perm : new cell i (0-based) is the caller's cell cell_perm[i]
    face_new  : the caller's face f (0-based) is face face_new[f] of `mesh`
    Node ids are unchanged, so face orientations - and with them the sign conventions of the trace basis - are the same in both
    numberings: trace coefficients and cell-wise coefficients only move, they never change."""

    def __init__(self, mesh, cell_perm, face_new):
        self.mesh, self.cell_perm, self.face_new = mesh, cell_perm, face_new

    def trace_to_original(self, x, nt):
        """trace vector of `mesh` (face-major, nt coefficients per face) -> the caller's face numbering"""
        return np.asarray(x).reshape(-1, nt)[self.face_new].reshape(-1)

    def cells_to_original(self, m_values):
        """cell-wise values (ncell, ...) of `mesh` -> the caller's cell numbering"""
        out = np.empty_like(m_values)
        out[self.cell_perm] = m_values
        return out


def renumber_mesh(mesh, perm=None, number=None):
    """Reorder the cells of `mesh` along the Morton curve (or by a given 0-based permutation) and renumber its faces the way
    the reference would for that element order - the pre-processing step that makes the contiguous-range partition of
    hdg_set_mesh on several GPUs local for any input order.  Both steps run on the device (`number`: another implementation of
    the first-encounter numbering with the signature of number_faces_gpu, e.g. the numpy mirror in tests)."""
    if perm is None:
        perm = order_cells_gpu(mesh.cells[:, :3], mesh.nodes)
    perm = np.asarray(perm, dtype=np.int64)
    old = mesh.cells[perm]
    cells, faces = (number or number_faces_gpu)(old[:, :3], mesh.nodes)
    if not np.array_equal(cells[:, :3], old[:, :3]):
        raise ValueError("renumber_mesh expects counter-clockwise cells (a mesh built by this package or the reference)")
    face_new = np.empty(mesh.faces.shape[0], np.int64)
    face_new[old[:, 3:].reshape(-1) - 1] = cells[:, 3:].reshape(-1) - 1
    facesets = {k: set((face_new[np.fromiter(v, np.int64, len(v)) - 1] + 1).tolist()) for k, v in mesh.facesets.items()}
    return RenumberedMesh(PolygonalMesh(cells, mesh.nodes, faces, facesets), perm, face_new)


def parse_mesh_triangle(root_file):
    """parse_mesh_triangle(root_file), src/triangle_mesh.jl:115-126 (.node/.edge/.ele reader)."""
    nodes = np.array([[float(r[1]), float(r[2])] for r in _triangle_rows(root_file + ".node")])
    marks = {}
    for r in _triangle_rows(root_file + ".edge"):
        a, b, mk = int(r[1]), int(r[2]), int(r[3])
        marks.setdefault((min(a, b), max(a, b)), mk)
    tri = np.array([[int(r[1]), int(r[2]), int(r[3])] for r in _triangle_rows(root_file + ".ele")], dtype=np.int64)
    # _check_node_data (src/generate_mesh.jl:49-57): make every cell counter-clockwise
    p = nodes[tri - 1]
    a, b = p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]
    flip = (a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]) < 0
    tri[flip] = tri[flip][:, [0, 2, 1]]
    cell_faces, faces = number_faces(tri)
    boundary = {f + 1 for f in range(faces.shape[0])
                if marks.get((min(faces[f, 0], faces[f, 1]), max(faces[f, 0], faces[f, 1])), -1) > 0}
    return PolygonalMesh(np.hstack([tri, cell_faces]), nodes, faces, {"boundary": boundary})


def quad_face_base(i, j, nx):
    """Number of faces created before quad (i,j) (0-based) by the first-encounter numbering of
    rectangle_mesh (src/generate_mesh.jl:20-46,130-138)."""
    if j == 0:
        return 4 * i + (1 if i > 0 else 0)
    return 4 * nx + 1 + (j - 1) * (3 * nx + 1) + 3 * i + (1 if i > 0 else 0)


def strip_partition(nx, ny, rank, world):
    """Mesh partition used on `world` GPUs: contiguous strips of quad rows = contiguous cell-id and
    face-id ranges (0-based, half-open).  Mirrors mesh_rectangle in csrc/hdg_mesh.cu."""
    if ny < world:
        raise ValueError("need at least one quad row per rank")
    j0, j1 = ny * rank // world, ny * (rank + 1) // world
    nface = 3 * nx * ny + nx + ny
    f0 = quad_face_base(0, j0, nx)
    f1 = quad_face_base(0, j1, nx) if j1 < ny else nface
    return dict(j0=j0, j1=j1, cell_begin=2 * nx * j0, cell_end=2 * nx * j1, face_begin=f0, face_end=f1,
                ncell_global=2 * nx * ny, nface_global=nface, ghost_cells=nx if j1 < ny else 0,
                ghost_faces=(nx if j0 > 0 else 0) + (2 * nx if j1 < ny else 0))


def ref_table(order, quad_degree, name):
    """Reference table from the library's host-side builder (no device needed)."""
    lib = _lib.load()
    cnt = C.c_int64()
    check(lib.hdg_ref_table(int(order), int(quad_degree or 0), name.encode(), None, C.byref(cnt)), None)
    buf = np.empty(cnt.value)
    check(lib.hdg_ref_table(int(order), int(quad_degree or 0), name.encode(), f64p(buf), C.byref(cnt)), None)
    return buf


# ----------------------------------------------------------------------------------------------
# context handle
# ----------------------------------------------------------------------------------------------
class _Context:
    def __init__(self, order, quad_degree=0, tau=1.0, source_id=1, device=-1, local_solver=0):
        lib = _lib.load()
        prm = Params(int(order), int(quad_degree or 0), float(tau), int(source_id), int(device), int(local_solver), 0)
        h = C.c_void_p()
        check(lib.hdg_create(C.byref(prm), C.byref(h)), None)
        self.h = h
        self.lib = lib
        self.mesh = None

    def close(self):
        if self.h:
            self.lib.hdg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def comm_init(self, dist, device=None):
        """Join the NCCL communicator of libhdg_b200: rank 0 creates the ncclUniqueId, torch.distributed
        (any backend - it is only plumbing) broadcasts the 128 bytes, every rank calls hdg_comm_init.
        Must be called before the mesh is set."""
        import torch
        rank, world = dist.get_rank(), dist.get_world_size()
        buf = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            check(self.lib.hdg_comm_unique_id(buf.ctypes.data_as(C.POINTER(C.c_uint8))), None)
        t = torch.from_numpy(buf)
        if device is not None:
            t = t.to(device)
        dist.broadcast(t, src=0)
        buf = t.cpu().numpy().copy()
        check(self.lib.hdg_comm_init(self.h, rank, world, buf.ctypes.data_as(C.POINTER(C.c_uint8))), self.h)

    def partition(self):
        out = np.zeros(8, dtype=np.int64)
        check(self.lib.hdg_get_partition(self.h, i64p(out)), self.h)
        keys = ("cell_begin", "cell_end", "face_begin", "face_end", "ncell_global", "nface_global", "ghost_cells", "ghost_faces")
        return dict(zip(keys, out.tolist()))

    def sizes(self):
        s = Sizes()
        check(self.lib.hdg_get_sizes(self.h, C.byref(s)), self.h)
        return s

    def set_mesh(self, mesh):
        self._dirichlet_faces = None     # the library takes the mesh's "boundary" set again
        if mesh._rect is not None and not getattr(mesh, "_modified", False):
            nx, ny, LL, UR = mesh._rect
            check(self.lib.hdg_set_rectangle_mesh(self.h, nx, ny, LL[0], LL[1], UR[0], UR[1]), self.h)
        else:
            bf = np.array(sorted(mesh.facesets["boundary"]), dtype=np.int64)
            check(self.lib.hdg_set_mesh(self.h, i64p(mesh.cells), mesh.cells.shape[0], f64p(mesh.nodes),
                                        mesh.nodes.shape[0], i64p(mesh.faces), mesh.faces.shape[0],
                                        i64p(bf), bf.size), self.h)
        self.mesh = mesh

    def download_mesh(self):
        s = self.sizes()
        cells = np.empty((s.ncell, 6), np.int64)
        nodes = np.empty((s.nnode, 2), np.float64)
        faces = np.empty((s.nface, 4), np.int64, order="F")
        bf = np.empty(s.nbface, np.int64)
        check(self.lib.hdg_get_mesh(self.h, i64p(cells), f64p(nodes), i64p(faces), i64p(bf)), self.h)
        return PolygonalMesh(cells, nodes, faces, {"boundary": set(bf.tolist())})

    def table(self, name):
        cnt = C.c_int64()
        check(self.lib.hdg_get_table(self.h, name.encode(), None, C.byref(cnt)), self.h)
        buf = np.empty(cnt.value)
        check(self.lib.hdg_get_table(self.h, name.encode(), f64p(buf), C.byref(cnt)), self.h)
        return buf


# ----------------------------------------------------------------------------------------------
# function spaces (src/ScalarFunctionSpaces.jl, src/VectorFunctionSpaces.jl, src/TraceFunctionSpaces.jl)
# ----------------------------------------------------------------------------------------------
class ScalarFunctionSpace:
    """ScalarFunctionSpace(mesh, fe; quad_degree = order+1), src/ScalarFunctionSpaces.jl:24-29."""
    components = 1

    def __init__(self, mesh, felem, quad_degree=None):
        self.mesh, self.fe = mesh, felem
        self.quad_degree = felem.order + 1 if quad_degree is None else int(quad_degree)
        self._tabs = None

    def getnbasefunctions(self):
        return self.fe.basis.getnbasefunctions()

    getnlocaldofs = getnbasefunctions

    def _tables(self):
        if self._tabs is None:   # raises UnsupportedRuleError like src/quadrature.jl:24
            self._tabs = {k: ref_table(self.fe.order, self.quad_degree, k)
                          for k in ("qpoints", "qweights", "fpoints", "fweights", "N", "dNdxi", "E", "T")}
        return self._tabs

    def getnquadpoints(self):
        return self._tables()["qweights"].size

    # ---- per-cell data and accessors (src/ScalarFunctionSpaces.jl:101-161), 1-based indices like the reference.  The hot
    # path does not go through these (hdg_assemble forms the blocks on the device); they are the inspection / hand-written
    # loop interface of the reference, served from the SAME reference tables the kernels consume (hdg_ref_table).
    def _arrays(self):
        if getattr(self, "_arr", None) is None:
            t, n = self._tables(), self.fe.basis.getnbasefunctions()
            nq, nfq = t["qweights"].size, t["fweights"].size
            self._arr = dict(N=t["N"].reshape(nq, n).T, dN=t["dNdxi"].reshape(nq, n, 2).transpose(1, 0, 2),
                             E=t["E"].reshape(3, nfq, n).transpose(2, 1, 0), qp=t["qpoints"].reshape(nq, 2), qw=t["qweights"],
                             fw=t["fweights"], n=n, nq=nq, nfq=nfq)
        return self._arr

    def reinit_(self, x):
        """reinit!(fs, x): x = (3,2) vertex coordinates (or a row of mesh.cells).  det(J) <= 0 raises like :110."""
        x = np.asarray(x, dtype=np.float64)
        if x.ndim == 1:
            x = self.mesh.nodes[x[:3].astype(np.int64) - 1]
        a = self._arrays()
        J = np.stack([x[1] - x[0], x[2] - x[0]], axis=1)            # sum_j x_j (x) dM_j/dxi, P1 geometry
        detJ = J[0, 0] * J[1, 1] - J[0, 1] * J[1, 0]
        if not detJ > 0.0:
            raise BadGeometryError(2, f"det(J) is not positive: det(J) = {detJ}")
        self.detJ = detJ
        self.Jinv = np.array([[J[1, 1], -J[0, 1]], [-J[1, 0], J[0, 0]]]) / detJ
        self.dNdx = a["dN"] @ self.Jinv                              # dNdxi . Jinv, :116
        wn = np.array([[-(J[1, 0] - J[1, 1]), J[0, 0] - J[0, 1]], [-J[1, 1], J[0, 1]], [J[1, 0], -J[0, 0]]])   # src/shapes.jl:80-87
        self.detJf = np.linalg.norm(wn, axis=1)
        self.normals = wn / self.detJf[:, None]
        self._x = x

    def getdetJdV(self, q):
        return self.detJ * self._arrays()["qw"][q - 1]

    def shape_value(self, *a):
        """shape_value(fs, q, i)  or  shape_value(fs, face, q, i, orientation=True)  (:138, :153)."""
        A = self._arrays()
        if len(a) == 2:
            return A["N"][a[1] - 1, a[0] - 1]
        face, q, i = a[0], a[1], a[2]
        ori = a[3] if len(a) > 3 else True
        return A["E"][i - 1, (q - 1) if ori else (A["nfq"] - q), face - 1]

    def shape_gradient(self, q, i):
        return self.dNdx[i - 1, q - 1]

    def getnfacequadpoints(self):
        return self._arrays()["nfq"]

    def getfacedetJdS(self, face, q):
        return self.detJf[face - 1] * self._arrays()["fw"][q - 1]

    def get_normal(self, face):
        return self.normals[face - 1]

    def spatial_coordinate(self, q, x=None):
        """spatial_coordinate(fs, q, x): x_q = sum_g M[g,q] x_g (src/DiscreteFunctions.jl:15-24)."""
        x = self._x if x is None else np.asarray(x, dtype=np.float64)
        r, s_ = self._arrays()["qp"][q - 1]
        return (1.0 - r - s_) * x[0] + r * x[1] + s_ * x[2]

    def function_value(self, f, q, x=None):
        return f(self.spatial_coordinate(q, x))


class VectorFunctionSpace(ScalarFunctionSpace):
    """VectorFunctionSpace(mesh, fe; quad_degree), src/VectorFunctionSpaces.jl:10-18."""
    components = 2

    def getnbasefunctions(self):
        return 2 * self.fe.basis.getnbasefunctions()

    getnlocaldofs = getnbasefunctions

    def shape_value(self, *a):
        """Vector-valued: dof i lives in component (i-1) div n with scalar function (i-1) mod n (:28-34, :49-57)."""
        n = self._arrays()["n"]
        i = a[1] if len(a) == 2 else a[2]
        out = np.zeros(2)
        sc = (i - 1) % n + 1
        out[(i - 1) // n] = ScalarFunctionSpace.shape_value(self, *((a[0], sc) if len(a) == 2 else (a[0], a[1], sc) + tuple(a[3:])))
        return out

    def shape_divergence(self, q, i):
        n = self._arrays()["n"]
        return self.dNdx[(i - 1) % n, q - 1, (i - 1) // n]             # tr(shape_gradient), :45-48


class ScalarTraceFunctionSpace:
    """ScalarTraceFunctionSpace(Wh, fe), src/TraceFunctionSpaces.jl:11-28."""
    components = 1

    def __init__(self, psp, felem):
        self.fs, self.fe, self.mesh = psp, felem, psp.mesh

    def getnbasefunctions(self):
        return self.fe.basis.getnbasefunctions()

    def getnlocaldofs(self):
        return 3 * self.getnbasefunctions()

    # src/TraceFunctionSpaces.jl:31-57
    def getnquadpoints(self):
        return self.fs._arrays()["nfq"]

    def shape_value(self, q, i):
        nt = self.getnbasefunctions()
        return ref_table(self.fe.order, self.fs.quad_degree, "T").reshape(-1, nt)[q - 1, i - 1]

    def getfacedetJdS(self, face, q):
        return self.fs.getfacedetJdS(face, q)


def getnbasefunctions(x):
    return x.getnbasefunctions()


def getnlocaldofs(x):
    return x.getnlocaldofs()


class TrialFunction:
    """TrialFunction(fs) with NaN-filled m_values, src/DiscreteFunctions.jl:27-54."""

    def __init__(self, fs):
        self.fs = fs
        nc = fs.mesh.getncells()
        if isinstance(fs, ScalarTraceFunctionSpace):
            self.m_values = np.full((nc, fs.getnbasefunctions(), 3), np.nan, order="F")
        else:
            self.m_values = np.full((nc, fs.getnbasefunctions()), np.nan, order="F")
        self.components = fs.components
        self._ctx = None

    def getnbasefunctions(self):
        return self.fs.getnbasefunctions()


class Dirichlet:
    """Dirichlet(u_hat, mesh, faceset, f) for trace spaces, src/boundary.jl:7-42: prescribed dofs
    face*nt-nt+i over the set in ascending face order, values = the reference's face projection of f,
    restated as it is written there:  values[k] = N with N += qr_weights[q] f(x_q) T_i(q) accumulated over the
    face's quadrature points - the accumulator N is NOT reset between the nt dofs of a face (:27) and the
    weights are the first nfq CELL weights (fs.fs.qr_weights, :33), not the face weights.  For f == 0 (the
    only case the reference's own driver uses) the values are exactly 0.  `corrected=True` gives the L2
    projection onto the orthonormal Legendre trace basis instead (face weights, accumulator per dof)."""

    # reference edges of the triangle, src/shapes.jl:19-23
    _REF_EDGES = np.array([[[1.0, 0.0], [0.0, 1.0]], [[0.0, 1.0], [0.0, 0.0]], [[0.0, 0.0], [1.0, 0.0]]])

    def __init__(self, u, mesh, faceset, f, corrected=False):
        fs = u.fs
        if not isinstance(fs, ScalarTraceFunctionSpace):
            raise NotImplementedError("only trace-space Dirichlet conditions are on the HDG path")
        nt = fs.getnbasefunctions()
        faces = sorted(mesh.getfaceset(faceset)) if isinstance(faceset, str) else sorted(faceset)
        for fi in faces:
            if mesh.faces[fi - 1, 3] != 0:
                raise AssertionError(f"Face {fi} is not in boundary")   # src/boundary.jl:22
        self.faces = np.array(faces, dtype=np.int64)
        self.prescribed_dofs = (self.faces[:, None] * nt - nt + np.arange(1, nt + 1)[None, :]).reshape(-1)
        order = fs.fe.order
        qd = fs.fs.quad_degree
        T = np.asarray(ref_table(order, qd, "T")).reshape(-1, nt).T        # T[i, q]
        s_pts = np.asarray(ref_table(order, qd, "fpoints"))
        w_cell = np.asarray(ref_table(order, qd, "qweights"))
        w_face = np.asarray(ref_table(order, qd, "fweights"))
        nfq = s_pts.size
        vals = np.zeros(self.prescribed_dofs.size)
        k = 0
        for fi in faces:
            cell = int(mesh.faces[fi - 1, 2])
            lidx = list(mesh.cells[cell - 1, 3:6]).index(fi)
            ori = face_orientation(mesh, cell, lidx + 1)
            x = mesh.nodes[mesh.cells[cell - 1, :3] - 1]                  # get_coordinates(cell, mesh)
            e1, e2 = self._REF_EDGES[lidx]
            acc = 0.0
            for i in range(nt):
                if corrected:
                    acc = 0.0
                for q in range(nfq):
                    qo = q if ori else nfq - 1 - q                        # spatial_coordinate, src/TraceFunctionSpaces.jl:47-57
                    eta = (1.0 - s_pts[qo]) * e1 + s_pts[qo] * e2
                    xq = (1.0 - eta[0] - eta[1]) * x[0] + eta[0] * x[1] + eta[1] * x[2]
                    w = w_face[q] if corrected else w_cell[q]
                    acc += w * float(f(xq)) * T[i, q]
                vals[k] = acc
                k += 1
        self.values = vals


# ----------------------------------------------------------------------------------------------
# the hot path: doassemble / apply! / solve / get_uσ! / errornorm
# ----------------------------------------------------------------------------------------------
def poisson_source(x):
    """f of examples/poisson2D_HDG.jl:55; passing this exact function selects the built-in
    device evaluation (source_id = 1) instead of sampling on the host."""
    return 2 * np.pi ** 2 * np.sin(np.pi * x[0]) * np.sin(np.pi * x[1])


def poisson_exact(x):
    """u_ex of examples/poisson2D_HDG.jl:216."""
    return np.sin(np.pi * x[0]) * np.sin(np.pi * x[1])


class TraceMatrix:
    """Handle of the assembled trace matrix K (SparseMatrixCSC in the reference)."""

    def __init__(self, ctx):
        self._ctx = ctx

    @property
    def shape(self):
        s = self._ctx.sizes()
        return (s.ndof, s.ndof)

    def pattern(self):
        """(colptr, rowval) of sparse(I,J,V), Int64 1-based (src/assembler.jl:47-49)."""
        s = self._ctx.sizes()
        colptr = np.empty(s.ndof + 1, np.int64)
        rowval = np.empty(s.nnz, np.int64)
        check(self._ctx.lib.hdg_get_pattern(self._ctx.h, i64p(colptr), i64p(rowval)), self._ctx.h)
        return colptr, rowval

    def nzval(self):
        s = self._ctx.sizes()
        v = np.empty(s.nnz)
        check(self._ctx.lib.hdg_get_values(self._ctx.h, f64p(v)), self._ctx.h)
        return v

    def to_scipy(self):
        import scipy.sparse as sp
        colptr, rowval = self.pattern()
        n = colptr.size - 1
        return sp.csc_matrix((self.nzval(), rowval - 1, colptr - 1), shape=(n, n))


class DeviceVector:
    def __init__(self, ctx, getter):
        self._ctx, self._getter = ctx, getter

    def to_numpy(self):
        s = self._ctx.sizes()
        v = np.empty(s.ndof)
        check(getattr(self._ctx.lib, self._getter)(self._ctx.h, f64p(v)), self._ctx.h)
        return v

    def __array__(self, dtype=None, copy=None):
        return self.to_numpy()


class LocalSolvers:
    """K_element / b_element (examples/poisson2D_HDG.jl:74-75), resident on the device."""

    def __init__(self, ctx, which):
        self._ctx, self._which = ctx, which

    def __len__(self):
        return self._ctx.sizes().ncell

    def __getitem__(self, cell0):
        s = self._ctx.sizes()
        Ke = np.empty((s.m, s.t), order="F")
        be = np.empty(s.m)
        check(self._ctx.lib.hdg_get_local(self._ctx.h, int(cell0) + 1, f64p(Ke), f64p(be)), self._ctx.h)
        return Ke if self._which == "K" else be


def doassemble(Vh, Wh, Mh, tau=1.0, f=poisson_source, device=-1, local_solver=0):
    """doassemble(Vh,Wh,Mh,tau) of examples/poisson2D_HDG.jl:58-186 -> (K, rhs, K_element, b_element)."""
    if Vh.quad_degree != Wh.quad_degree or Vh.fe.order != Wh.fe.order or Mh.fe.order != Wh.fe.order:
        raise ValueError("Vh, Wh, Mh must share order and quad_degree")
    mesh = Wh.mesh
    builtin = f is poisson_source
    ctx = _Context(Wh.fe.order, Wh.quad_degree, tau, 1 if builtin else 0, device, local_solver)
    ctx.set_mesh(mesh)
    if not builtin:
        # function_value(f, Wh, cell, q) sampled on the host: x_q = sum_g M[g,q] x_g  (src/DiscreteFunctions.jl:6-24)
        qp = Wh._tables()["qpoints"].reshape(-1, 2)
        M = np.stack([1 - qp[:, 0] - qp[:, 1], qp[:, 0], qp[:, 1]], axis=0)          # (3,nq)
        xc = mesh.nodes[mesh.cells[:, :3] - 1]                                        # (ncell,3,2)
        xq = np.einsum("gq,cgd->cqd", M, xc)
        fq = np.ascontiguousarray(np.vectorize(lambda a, b: f((a, b)))(xq[..., 0], xq[..., 1]), dtype=np.float64)
        check(ctx.lib.hdg_set_source_values(ctx.h, f64p(fq)), ctx.h)
    check(ctx.lib.hdg_assemble(ctx.h), ctx.h)
    return TraceMatrix(ctx), DeviceVector(ctx, "hdg_get_rhs"), LocalSolvers(ctx, "K"), LocalSolvers(ctx, "b")


def apply_(K, b, dbc):
    """apply!(K, b, dbc), src/boundary.jl:121-158 (in place on the device)."""
    ctx = K._ctx
    s = ctx.sizes()
    mine = np.array(sorted(ctx.mesh.facesets["boundary"]), dtype=np.int64)
    if getattr(ctx, "_dirichlet_faces", None) is not None:
        mine = ctx._dirichlet_faces
    if dbc.faces.size != s.nbface or not np.array_equal(dbc.faces, mine):
        # another named face set ("bottom", "left", ... src/generate_mesh.jl:60-89): the rest of the boundary stays natural
        faces = np.ascontiguousarray(dbc.faces, dtype=np.int64)
        check(ctx.lib.hdg_set_dirichlet_faces(ctx.h, i64p(faces), faces.size), ctx.h)
        ctx._dirichlet_faces = faces
    vals = dbc.values if np.any(dbc.values != 0) else None
    check(ctx.lib.hdg_apply_dirichlet(ctx.h, f64p(vals)), ctx.h)


def meandiag(K):
    m = C.c_double()
    check(K._ctx.lib.hdg_get_meandiag(K._ctx.h, C.byref(m)), K._ctx.h)
    return m.value


def solve(K, b, rtol=1e-13, maxit=200000, precond="jacobi"):
    """u_hat = K \\ b (examples/poisson2D_HDG.jl:195) by Jacobi-PCG on the sign-fixed system
    (precond="block" uses the nt x nt face-diagonal blocks, precond="mg" adds the P1-vertex multigrid V-cycle).  Returns (DeviceVector, info dict)."""
    ctx = K._ctx
    check(ctx.lib.hdg_set_preconditioner(ctx.h, {"jacobi": 0, "block": 1, "mg": 2}[precond]), ctx.h)
    info = SolveInfo()
    check(ctx.lib.hdg_solve(ctx.h, float(rtol), int(maxit), C.byref(info)), ctx.h)
    d = dict(iterations=info.iterations, converged=bool(info.converged), relres=info.relres,
             bnorm=info.bnorm, solve_ms=info.solve_ms)
    return DeviceVector(ctx, "hdg_get_trace"), d


def get_usigma_(sigma_h, u_h, uhat_h, uhat, K_e, b_e, mesh):
    """get_uσ!(σ_h,u_h,û_h,û,K_e,b_e,mesh), examples/poisson2D_HDG.jl:197-212."""
    ctx = K_e._ctx
    if isinstance(uhat, np.ndarray):
        check(ctx.lib.hdg_set_trace(ctx.h, f64p(np.ascontiguousarray(uhat, dtype=np.float64))), ctx.h)
    check(ctx.lib.hdg_recover(ctx.h), ctx.h)
    check(ctx.lib.hdg_get_mvalues(ctx.h, f64p(sigma_h.m_values), f64p(u_h.m_values), f64p(uhat_h.m_values)), ctx.h)
    u_h._ctx = sigma_h._ctx = uhat_h._ctx = ctx


def errornorm(u_h, u_ex=poisson_exact, norm_type="L2"):
    """errornorm(u_h,u_ex): squared L2 error, src/DiscreteFunctions.jl:97-120."""
    if norm_type != "L2":
        raise ValueError(f"Norm {norm_type} not available")
    ctx = u_h._ctx
    if ctx is None:
        raise HDGError(1, "errornorm before get_usigma_")
    e = C.c_double()
    if u_ex is poisson_exact:
        check(ctx.lib.hdg_errornorm(ctx.h, 1, C.byref(e)), ctx.h)
        return e.value
    # any other u_ex: sampled on the host at x_q = sum_g M[g,q] x_g like the source (function_value,
    # src/DiscreteFunctions.jl:6-24, :108), the quadrature sum runs on the device
    mesh = u_h.fs.mesh
    qp = np.asarray(ref_table(u_h.fs.fe.order, u_h.fs.quad_degree, "qpoints")).reshape(-1, 2)
    M = np.stack([1 - qp[:, 0] - qp[:, 1], qp[:, 0], qp[:, 1]], axis=0)
    xq = np.einsum("gq,cgd->cqd", M, mesh.nodes[mesh.cells[:, :3] - 1])
    uq = np.ascontiguousarray(np.vectorize(lambda a, b: u_ex((a, b)))(xq[..., 0], xq[..., 1]), dtype=np.float64)
    check(ctx.lib.hdg_errornorm_values(ctx.h, f64p(uq), C.byref(e)), ctx.h)
    return e.value


def nodal_avg(u_h):
    """nodal_avg(u_h), src/DiscreteFunctions.jl:81-95 (u_h must have been filled by get_usigma_)."""
    ctx = u_h._ctx
    if ctx is None:
        raise HDGError(1, "nodal_avg before get_usigma_")
    out = np.empty(u_h.fs.mesh.getnnodes())
    check(ctx.lib.hdg_nodal_avg(ctx.h, f64p(out)), ctx.h)
    return out


def poisson2D_HDG(mesh=None, order=1, quad_degree=None, tau=1.0, rtol=1e-13, maxit=200000, precond="jacobi"):
    """The driver examples/poisson2D_HDG.jl:37-218 end to end.  Returns a dict of results.
    precond: "jacobi" (default), "block", or "mg" (rectangle_mesh triangulations), see `solve`."""
    if mesh is None:
        mesh = rectangle_mesh(TriangleCell, (10, 10), (0.0, 0.0), (1.0, 1.0))
    fe = GenericFiniteElement(Dubiner(2, RefTetrahedron, order))
    Wh = ScalarFunctionSpace(mesh, fe, quad_degree)
    Vh = VectorFunctionSpace(mesh, fe, quad_degree)
    Mh = ScalarTraceFunctionSpace(Wh, GenericFiniteElement(Legendre(1, RefTetrahedron, order)))
    uhat_h, sigma_h, u_h = TrialFunction(Mh), TrialFunction(Vh), TrialFunction(Wh)
    dbc = Dirichlet(uhat_h, mesh, "boundary", lambda x: 0)
    K, b, K_e, b_e = doassemble(Vh, Wh, Mh, tau)
    apply_(K, b, dbc)
    uhat, info = solve(K, b, rtol, maxit, precond)
    get_usigma_(sigma_h, u_h, uhat_h, uhat, K_e, b_e, mesh)
    err2 = errornorm(u_h, poisson_exact)
    return dict(K=K, b=b, K_e=K_e, b_e=b_e, uhat=uhat, sigma_h=sigma_h, u_h=u_h, uhat_h=uhat_h,
                err2=err2, info=info, dbc=dbc, mesh=mesh)
