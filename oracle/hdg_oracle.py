"""CPU oracle for the HDG Poisson hot path of Paulms/HDiscontinuousGalerkin.jl.

TEST INFRASTRUCTURE ONLY.  This module is a loop-faithful numpy restatement of
the reference algorithm (Julia, cannot run here: no `julia` binary, see
DESIGN.md).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
cpu_baseline / `--impl reference` legs may import it, and only as the checker
(or as the CPU baseline being timed) - never on the product path.

Pinning: every golden vector of the reference's own tests for this path is
checked in tests/test_oracle_goldens.py (mesh numbering test/test_mesh.jl:17-43,
quadrature test/test_quadrature.jl:6-25, bases test/test_basis.jl:6-12,62-93,
geometry test/test_ScalarFuncSp.jl:15-32, local blocks Ae/Be/Ce/Ee/He
test/test_FunctionSpace.jl:49-72,125-126,176-178, error bounds :243 and
examples/poisson2D_HDG.jl:218).  Fe, be, K_e, At, K, u_hat have NO golden in the
reference ("parity unpinned" for those): they are pinned only through
the two scalar error bounds.

All `file:line` citations are relative to /root/reference.
Indices in this module are 0-based internally; arrays that mirror reference
integer data (`cells`, `cell_faces`, `faces`, Dirichlet dofs, CSC pattern) hold
the reference's 1-based values so they can be compared bit for bit.
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

REF_EDGE_NODES = ((1, 2), (2, 0), (0, 1))  # src/mesh.jl:26 ((2,3),(3,1),(1,2)) 0-based


# --------------------------------------------------------------------------
# Mesh  (src/mesh.jl:17-54, src/generate_mesh.jl:1-57,101-143, src/triangle_mesh.jl)
# --------------------------------------------------------------------------
@dataclass
class Mesh:
    cells: np.ndarray       # (ncell,3) int64, 1-based node ids   src/mesh.jl:17-20
    cell_faces: np.ndarray  # (ncell,3) int64, 1-based face ids
    nodes: np.ndarray       # (nnode,2) float64                   src/mesh.jl:5-7
    faces: np.ndarray       # (nface,4) int64: v1 v2 cell1 cell2/0  src/mesh.jl:46
    facesets: dict = field(default_factory=dict)  # name -> set of 1-based face ids

    @property
    def ncells(self):
        return self.cells.shape[0]

    @property
    def nfaces(self):
        return self.faces.shape[0]

    @property
    def nnodes(self):
        return self.nodes.shape[0]

    def boundary_faces_sorted(self):
        return np.array(sorted(self.facesets["boundary"]), dtype=np.int64)


def _check_node_data(nodes, n1, n2, n3):
    """src/generate_mesh.jl:49-57 - make the triangle counter-clockwise."""
    a = nodes[n2 - 1] - nodes[n1 - 1]
    b = nodes[n3 - 1] - nodes[n1 - 1]
    if a[0] * b[1] - a[1] * b[0] < 0:
        return (n1, n3, n2)
    return (n1, n2, n3)


def _build_cells_sequential(tri_nodes, nodes, nface_alloc):
    """First-encounter face numbering, src/generate_mesh.jl:20-46 (same loop in
    src/triangle_mesh.jl:66-101).  `tri_nodes`: iterable of CCW 1-based triples."""
    facesdict = {}
    faces = np.zeros((nface_alloc, 4), dtype=np.int64)
    cells, cell_faces = [], []
    face_idx = 0
    for n_el, el_nodes in enumerate(tri_nodes, start=1):
        el_faces = [0, 0, 0]
        for i, (k1, k2) in enumerate(REF_EDGE_NODES):
            v1, v2 = el_nodes[k1], el_nodes[k2]
            key = (min(v1, v2), max(v1, v2))
            fid = facesdict.get(key)
            if fid is not None:
                el_faces[i] = fid
                if n_el != faces[fid - 1, 3]:
                    faces[fid - 1, 3] = n_el
            else:
                face_idx += 1
                facesdict[key] = face_idx
                faces[face_idx - 1] = (v1, v2, n_el, 0)
                el_faces[i] = face_idx
        cells.append(el_nodes)
        cell_faces.append(el_faces)
    return (np.array(cells, dtype=np.int64), np.array(cell_faces, dtype=np.int64),
            faces[:face_idx].copy())


def rectangle_mesh(nx, ny, LL=(0.0, 0.0), UR=(1.0, 1.0)):
    """rectangle_mesh(TriangleCell,(nx,ny),LL,UR), src/generate_mesh.jl:101-143."""
    LL = np.asarray(LL, float)
    UR = np.asarray(UR, float)
    LR = np.array([UR[0], LL[1]])
    UL = np.array([LL[0], UR[1]])
    nnx, nny = nx + 1, ny + 1
    # _generate_2d_nodes!  src/generate_mesh.jl:1-18 (called with node counts)
    nodes = np.empty((nnx * nny, 2))
    p = 0
    for i in range(nny):
        rb = i / (nny - 1)
        x0 = LL[0] * (1 - rb) + rb * UL[0]
        x1 = LR[0] * (1 - rb) + rb * UR[0]
        y0 = LL[1] * (1 - rb) + rb * UL[1]
        y1 = LR[1] * (1 - rb) + rb * UR[1]
        for j in range(nnx):
            r = j / (nnx - 1)
            nodes[p, 0] = x0 * (1 - r) + r * x1
            nodes[p, 1] = y0 * (1 - r) + r * y1
            p += 1
    na = lambda i, j: i + (j - 1) * nnx  # node_array[i,j], 1-based, column-major reshape

    def tris():
        for j in range(1, ny + 1):
            for i in range(1, nx + 1):
                yield _check_node_data(nodes, na(i, j), na(i + 1, j), na(i, j + 1))        # lower-left
                yield _check_node_data(nodes, na(i + 1, j), na(i + 1, j + 1), na(i, j + 1))  # upper-right

    cells, cell_faces, faces = _build_cells_sequential(tris(), nodes, nx + ny + 3 * nx * ny)
    # _get_rectangular_boundary_sets  src/generate_mesh.jl:60-89
    sets = {"bottom": set(), "right": set(), "top": set(), "left": set()}
    for k in range(faces.shape[0]):
        if faces[k, 3] == 0:
            a, b = nodes[faces[k, 0] - 1], nodes[faces[k, 1] - 1]
            if a[1] == LL[1] and b[1] == LL[1]:
                sets["bottom"].add(k + 1)
            elif a[0] == UR[0] and b[0] == UR[0]:
                sets["right"].add(k + 1)
            elif a[1] == UR[1] and b[1] == UR[1]:
                sets["top"].add(k + 1)
            elif a[0] == LL[0] and b[0] == LL[0]:
                sets["left"].add(k + 1)
            else:
                raise ValueError(f"Face {k+1} belongs to one cell but is not in boundary")
    sets["boundary"] = sets["bottom"] | sets["right"] | sets["left"] | sets["top"]
    return Mesh(cells, cell_faces, nodes, faces, sets)


_NUM = re.compile(r"\b((\d*\.)?\d+)\b")  # src/triangle_mesh.jl:3


def _triangle_rows(path):
    """Data rows of a Triangle file: skip blank / '#' lines and the header row
    (src/triangle_mesh.jl:8-24)."""
    rows, first = [], True
    with open(path) as fh:
        for ln in fh:
            if re.match(r"^\s*(?:#|$)", ln):
                continue
            if first:
                first = False
                continue
            rows.append([m.group(0) for m in _NUM.finditer(ln)])
    return rows


def parse_mesh_triangle(root):
    """parse_mesh_triangle, src/triangle_mesh.jl:115-126."""
    nodes = np.array([[float(r[1]), float(r[2])] for r in _triangle_rows(root + ".node")])
    faces_ref = {}
    for r in _triangle_rows(root + ".edge"):          # :27-46, first marker wins
        a, b, mark = int(r[1]), int(r[2]), int(r[3])
        faces_ref.setdefault((min(a, b), max(a, b)), mark)
    tri = [_check_node_data(nodes, int(r[1]), int(r[2]), int(r[3]))
           for r in _triangle_rows(root + ".ele")]
    cells, cell_faces, faces = _build_cells_sequential(tri, nodes, len(faces_ref))
    boundary = set()
    for f in range(faces.shape[0]):                    # :80-96 marker > 0
        v1, v2 = faces[f, 0], faces[f, 1]
        if faces_ref.get((min(v1, v2), max(v1, v2)), -1) > 0:
            boundary.add(f + 1)
    return Mesh(cells, cell_faces, nodes, faces, {"boundary": boundary})


def face_orientation(mesh, cell, lface):
    """src/mesh.jl:51-54 (0-based cell / local face)."""
    k1, k2 = REF_EDGE_NODES[lface]
    return bool(mesh.cells[cell, k2] > mesh.cells[cell, k1])


# --------------------------------------------------------------------------
# Quadrature (src/quadrature.jl, src/StrangQuad.jl, src/GrundmannMoellerQuad.jl)
# --------------------------------------------------------------------------
def strang(order):
    """src/StrangQuad.jl:1-61.  Returns (points (nq,2), weights (nq,))."""
    if order in (0, 1):
        p = [(1 / 3, 1 / 3)]
        w = [0.5]
    elif order == 2:
        p = [(1 / 6, 1 / 6), (1 / 6, 2 / 3), (2 / 3, 1 / 6)]
        w = [1 / 6] * 3
    elif order == 3:
        a, b, c = 0.659027622374092, 0.231933368553031, 0.109039009072877
        p = [(a, b), (a, c), (b, a), (b, c), (c, a), (c, b)]
        w = [1 / 12] * 6
    elif order == 4:
        a, b = 0.816847572980459, 0.091576213509771
        c, d = 0.108103018168070, 0.445948490915965
        p = [(a, b), (b, a), (b, b), (c, d), (d, c), (d, d)]
        w = [0.109951743655322 * 0.5] * 3 + [0.223381589678011 * 0.5] * 3
    elif order == 5:
        a, b = 0.79742698535308720, 0.10128650732345633
        c, d = 0.05971587178976981, 0.47014206410511505
        t = 0.33333333333333333
        p = [(t, t), (a, b), (b, a), (b, b), (c, d), (d, c), (d, d)]
        w = [0.225 * 0.5] + [0.12593918054482717 * 0.5] * 3 + [0.13239415278850616 * 0.5] * 3
    elif order == 6:
        a, b = 0.873821971016996, 0.063089014491502
        c, d = 0.501426509658179, 0.249286745170910
        e, f, g = 0.636502499121399, 0.310352451033785, 0.053145049844816
        p = [(a, b), (b, a), (b, b), (c, d), (d, c), (d, d),
             (e, f), (e, g), (f, e), (f, g), (g, e), (g, f)]
        w = ([0.050844906370207 * 0.5] * 3 + [0.116786275726379 * 0.5] * 3
             + [0.082851075618374 * 0.5] * 6)
    else:
        raise ValueError(f"Strang rule of order {order} not available")
    return np.array(p, float), np.array(w, float)


def _all_exponentials(n, k):
    """src/GrundmannMoellerQuad.jl:29-49: all k-tuples of non-negative ints summing to n,
    in the reference's enumeration order."""
    a = [0] * k
    a[0] = n
    out = [tuple(a)]
    t, h = n, 0
    while a[k - 1] != n:
        if 1 < t:
            h = 0
        h += 1
        t = a[h - 1]
        a[h - 1] = 0
        a[0] = t - 1
        a[h] += 1
        out.append(tuple(a))
    return out


def grundmann_moeller(s, n=2):
    """src/GrundmannMoellerQuad.jl:10-27."""
    d = 2 * s + 1
    pts, wts = [], []
    for i in range(s + 1):
        w = ((-1) ** i * 2.0 ** (-2 * s) * float(d + n - 2 * i) ** d) / (
            math.factorial(i) * math.factorial(d + n - i))
        for p in _all_exponentials(s - i, n + 1):
            pts.append([(2 * p[j + 1] + 1) / (d + n - 2 * i) for j in range(n)])
            wts.append(w)
    wts = np.array(wts)
    return np.array(pts, float), wts / np.sum(2 * wts)


def default_quad_2d(order):
    """src/quadrature.jl:17-26."""
    if order <= 6:
        return strang(order)
    if order % 2 == 1:
        return grundmann_moeller((order - 1) // 2)
    raise ValueError(f"Quadrature rule of order {order} not available")


def gauss_legendre_01(npts):
    """src/quadrature.jl:29-39: `npts`-point Gauss-Legendre mapped to (0,1).
    FastGaussQuadrature.gausslegendre (third party, absent) == numpy leggauss to rounding."""
    x, w = np.polynomial.legendre.leggauss(npts)
    return (x + 1.0) / 2.0, 0.5 * w


REF_EDGES = np.array([[[1.0, 0.0], [0.0, 1.0]],
                      [[0.0, 1.0], [0.0, 0.0]],
                      [[0.0, 0.0], [1.0, 0.0]]])  # src/shapes.jl:19-23


def face_quad_points(s):
    """src/quadrature.jl:60-75: eta[l,p,:] = (1-s_p) e_l^1 + s_p e_l^2."""
    return (1.0 - s)[None, :, None] * REF_EDGES[:, None, 0, :] + s[None, :, None] * REF_EDGES[:, None, 1, :]


# --------------------------------------------------------------------------
# Bases (src/basis.jl)
# --------------------------------------------------------------------------
def jacobi(x, p, alpha, beta):
    """src/basis.jl:132-150 three-term recursion."""
    a = 1.0
    b = ((2 + alpha + beta) * x + alpha - beta) / 2
    if p <= 0:
        return a
    if p == 1:
        return b
    for n in range(2, p + 1):
        a1 = 2 * n * (n + alpha + beta) * (2 * n - 2 + alpha + beta)
        a2 = (2 * n - 1 + alpha + beta) * (alpha + beta) * (alpha - beta)
        a3 = (2 * n - 2 + alpha + beta) * (2 * n - 1 + alpha + beta) * (2 * n + alpha + beta)
        a4 = 2 * (n - 1 + alpha) * (n - 1 + beta) * (2 * n + alpha + beta)
        a, b = b, ((a2 + a3 * x) * b - a4 * a) / a1
    return b


def djacobi(x, n, alpha, beta):
    """src/basis.jl:157-162."""
    if n <= 0:
        return 0.0
    return (n + alpha + beta + 1) / 2 * jacobi(x, n - 1, alpha + 1, beta + 1)


def dubiner_nm(j):
    """Degree pair (n,m) of basis function j (1-based), src/basis.jl:211-218."""
    t = -1.5 + 0.5 * math.sqrt(1 + 8 * j)
    d = math.ceil(t - 1e-12)
    n = (d + 1) * (d + 2) // 2 - j
    return n, d - n


def dubiner_value(j, r, s):
    """Dubiner basis j at (r,s): src/basis.jl:169-181 (== closed forms :65-86, pinned to eps
    by test/test_basis.jl:6-12)."""
    n, m = dubiner_nm(j)
    xi = -1.0 if abs(r) < np.finfo(float).eps else 2 * r / (1 - s) - 1
    eta = 2 * s - 1
    P = jacobi(xi, n, 0, 0) * jacobi(eta, m, 2 * n + 1, 0) * ((1 - eta) / 2) ** n
    N = math.sqrt(2 / ((2 * n + 1) * (m + n + 1)))
    return 2 * P / N


def dubiner_grad(j, r, s):
    """src/basis.jl:188-204."""
    n, m = dubiner_nm(j)
    xi = -1.0 if abs(r) < np.finfo(float).eps else 2 * r / (1 - s) - 1
    k = 2 * n + 1
    eta = 2 * s - 1
    dpn = djacobi(xi, n, 0, 0)
    pn = jacobi(xi, n, 0, 0)
    pm = jacobi(eta, m, k, 0)
    dpm = djacobi(eta, m, k, 0)
    hn = ((1 - eta) / 2) ** n
    px = 2 / (1 - eta) * dpn * pm * hn
    N = math.sqrt(2 / ((2 * n + 1) * (m + n + 1)))
    hn1 = ((1 - eta) / 2) ** (n - 1) if n > 0 else 0.0
    py = (2 * r / (1 - s) ** 2 * dpn * hn - n * hn1 * pn) * pm + 2 * pn * hn * dpm
    return np.array([4 * px / N, 2 * py / N])


def legendre_value(k, x):
    """Orthonormal Legendre on (0,1), k 1-based: src/basis.jl:351-354."""
    return math.sqrt(2 * (k - 1) + 1) * jacobi(2 * x - 1, k - 1, 0.0, 0.0)


# --------------------------------------------------------------------------
# Reference tables (src/ScalarFunctionSpaces.jl:31-99, src/TraceFunctionSpaces.jl:11-28)
# --------------------------------------------------------------------------
@dataclass
class Tables:
    order: int
    quad_degree: int
    n: int          # scalar dofs per cell
    nt: int         # trace dofs per face
    nq: int
    nfq: int
    qpts: np.ndarray    # (nq,2)
    qw: np.ndarray      # (nq,)  sum = 0.5
    fs: np.ndarray      # (nfq,) GL points on (0,1)
    fw: np.ndarray      # (nfq,) GL weights, sum = 1
    N: np.ndarray       # (n,nq)
    dN: np.ndarray      # (n,nq,2)   dN/dxi
    M: np.ndarray       # (3,nq)     P1 geometry map
    dM: np.ndarray      # (3,2)
    E: np.ndarray       # (n,nfq,3)  cell basis at face points
    T: np.ndarray       # (nt,nfq)   trace basis at GL points
    L: np.ndarray       # (3,nfq,3)  geometry map at face points [g,p,l]


def build_tables(order, quad_degree=None):
    """ScalarFunctionSpace / VectorFunctionSpace / ScalarTraceFunctionSpace tables.
    Default quad_degree = order+1, src/ScalarFunctionSpaces.jl:24-25."""
    if quad_degree is None:
        quad_degree = order + 1
    n = (order + 1) * (order + 2) // 2
    nt = order + 1
    qpts, qw = default_quad_2d(quad_degree)
    fs_, fw = gauss_legendre_01(quad_degree)
    nq, nfq = len(qw), len(fw)
    N = np.empty((n, nq))
    dN = np.empty((n, nq, 2))
    for q in range(nq):
        for i in range(n):
            N[i, q] = dubiner_value(i + 1, qpts[q, 0], qpts[q, 1])
            dN[i, q] = dubiner_grad(i + 1, qpts[q, 0], qpts[q, 1])
    geo = lambda r, s: np.array([1 - r - s, r, s])   # Lagrange{2,.,1}, src/basis.jl:41-46
    M = np.stack([geo(*qpts[q]) for q in range(nq)], axis=1)
    dM = np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]])
    eta = face_quad_points(fs_)
    E = np.empty((n, nfq, 3))
    L = np.empty((3, nfq, 3))
    for l in range(3):
        for p in range(nfq):
            L[:, p, l] = geo(*eta[l, p])
            for i in range(n):
                E[i, p, l] = dubiner_value(i + 1, eta[l, p, 0], eta[l, p, 1])
    T = np.array([[legendre_value(j + 1, fs_[p]) for p in range(nfq)] for j in range(nt)])
    return Tables(order, quad_degree, n, nt, nq, nfq, qpts, qw, fs_, fw, N, dN, M, dM, E, T, L)


# --------------------------------------------------------------------------
# Per-cell geometry: reinit!  (src/ScalarFunctionSpaces.jl:101-132, src/shapes.jl:80-87)
# --------------------------------------------------------------------------
@dataclass
class CellGeom:
    J: np.ndarray
    detJ: float
    Jinv: np.ndarray
    detJf: np.ndarray   # (3,)
    normals: np.ndarray  # (3,2)


def reinit(tab, x):
    """x: (3,2) vertex coordinates."""
    J = np.zeros((2, 2))
    for j in range(3):
        J += np.outer(x[j], tab.dM[j])
    detJ = J[0, 0] * J[1, 1] - J[0, 1] * J[1, 0]
    if not detJ > 0.0:
        raise ValueError(f"det(J) is not positive: det(J) = {detJ}")
    Jinv = np.array([[J[1, 1], -J[0, 1]], [-J[1, 0], J[0, 0]]]) / detJ
    wn = np.array([[-(J[1, 0] - J[1, 1]), J[0, 0] - J[0, 1]],
                   [-J[1, 1], J[0, 1]],
                   [J[1, 0], -J[0, 0]]])
    detJf = np.sqrt(wn[:, 0] ** 2 + wn[:, 1] ** 2)
    if not np.all(detJf > 0.0):
        raise ValueError("det(Jf) is not positive")
    return CellGeom(J, detJ, Jinv, detJf, wn / detJf[:, None])


# --------------------------------------------------------------------------
# Local blocks + condensation  (examples/poisson2D_HDG.jl:80-174)
# --------------------------------------------------------------------------
def source_poisson(x):
    """f of examples/poisson2D_HDG.jl:55."""
    return 2 * math.pi ** 2 * math.sin(math.pi * x[0]) * math.sin(math.pi * x[1])


def exact_poisson(x):
    """u_ex of examples/poisson2D_HDG.jl:216."""
    return math.sin(math.pi * x[0]) * math.sin(math.pi * x[1])


def local_blocks(tab, x, orient, f=source_poisson, tau=1.0, fq=None):
    """The quadrature loops of examples/poisson2D_HDG.jl:88-153 for one cell.
    x (3,2) vertices, orient (3,) bools.  Returns dict A,B,C,E,F,H,be and geometry."""
    n, nt, nq, nfq = tab.n, tab.nt, tab.nq, tab.nfq
    nv = 2 * n
    g = reinit(tab, x)
    dNdx = np.einsum("iqa,ab->iqb", tab.dN, g.Jinv)      # dNdxi . Jinv  (:114-116)
    A = np.zeros((nv, nv)); B = np.zeros((nv, n)); C = np.zeros((n, n))
    E = np.zeros((nv, 3 * nt)); F = np.zeros((n, 3 * nt)); H = np.zeros((3 * nt, 3 * nt))
    be = np.zeros(n)
    for q in range(nq):                                   # :88-104
        dO = g.detJ * tab.qw[q]
        for c in range(2):
            sl = slice(c * n, (c + 1) * n)
            A[sl, sl] += np.outer(tab.N[:, q], tab.N[:, q]) * dO
            B[sl, :] += np.outer(dNdx[:, q, c], tab.N[:, q]) * dO
    for q in range(nq):                                   # :106-114
        dO = g.detJ * tab.qw[q]
        fh = fq[q] if fq is not None else f(tab.M[:, q] @ x)
        be += fh * tab.N[:, q] * dO
    for l in range(3):                                    # :116-153
        for q in range(nfq):
            dS = g.detJf[l] * tab.fw[q]
            qo = q if orient[l] else nfq - 1 - q
            w = tab.E[:, q, l]
            C += tau * np.outer(w, w) * dS
            wo = tab.E[:, qo, l]
            F[:, nt * l:nt * (l + 1)] += tau * np.outer(wo, tab.T[:, q]) * dS
            for c in range(2):
                E[c * n:(c + 1) * n, nt * l:nt * (l + 1)] += np.outer(wo * g.normals[l, c], tab.T[:, q]) * dS
            H[nt * l:nt * (l + 1), nt * l:nt * (l + 1)] += np.outer(tab.T[:, q], tab.T[:, q]) * dS
    return dict(A=A, B=B, C=C, E=E, F=F, H=H, be=be, geom=g)


def condense(blk):
    """examples/poisson2D_HDG.jl:155-174.  numpy.linalg.solve == LAPACK getrf/getrs,
    the same partial-pivot LU Julia's factorize(Array(Me)) calls."""
    A, B, C, E, F, H, be = (blk[k] for k in "A B C E F H be".split())
    nv, n = B.shape
    Me = np.block([[A, -B], [B.T, C]])
    EF = np.vstack([-E, F])
    G = np.vstack([E, F]).T
    K_e = np.linalg.solve(Me, EF)
    At = G @ K_e - H
    b_e = np.linalg.solve(Me, np.concatenate([np.zeros(nv), be]))
    bt = -G @ b_e
    return K_e, b_e, At, bt


@dataclass
class Assembled:
    K: sp.csc_matrix
    rhs: np.ndarray
    K_e: np.ndarray   # (ncell, m, t)
    b_e: np.ndarray   # (ncell, m)
    At: np.ndarray    # (ncell, t, t)
    bt: np.ndarray    # (ncell, t)
    gdof: np.ndarray  # (ncell, t) 1-based


def orientations(mesh):
    c = mesh.cells
    return np.stack([c[:, k2] > c[:, k1] for (k1, k2) in REF_EDGE_NODES], axis=1)


def doassemble(mesh, tab, f=source_poisson, tau=1.0):
    """doassemble, examples/poisson2D_HDG.jl:58-186; COO->CSC like sparse(I,J,V)
    (src/assembler.jl:31-49: duplicates summed, explicit zeros kept, rows ascending)."""
    n, nt = tab.n, tab.nt
    m, t = 3 * n, 3 * nt
    nc = mesh.ncells
    ndof = mesh.nfaces * nt
    ori = orientations(mesh)
    K_e = np.empty((nc, m, t)); b_e = np.empty((nc, m))
    At = np.empty((nc, t, t)); bt = np.empty((nc, t))
    gdof = np.empty((nc, t), dtype=np.int64)
    rhs = np.zeros(ndof)
    I = np.empty((nc, t * t), dtype=np.int64); Jc = np.empty_like(I)
    for c in range(nc):
        x = mesh.nodes[mesh.cells[c] - 1]
        blk = local_blocks(tab, x, ori[c], f, tau)
        K_e[c], b_e[c], At[c], bt[c] = condense(blk)
        gd = np.array([fi * nt - (nt - j) for fi in mesh.cell_faces[c] for j in range(1, nt + 1)])
        gdof[c] = gd
        # assemble!: V = Ke column-major, I = edof per column, J = edof[j] repeated
        I[c] = np.tile(gd, t)
        Jc[c] = np.repeat(gd, t)
        for i in range(t):
            rhs[gd[i] - 1] += bt[c, i]
    V = np.transpose(At, (0, 2, 1)).reshape(nc, t * t)   # column-major flattening
    K = sp.coo_matrix((V.ravel(), (I.ravel() - 1, Jc.ravel() - 1)), shape=(ndof, ndof)).tocsc()
    K.sum_duplicates()
    K.sort_indices()
    return Assembled(K, rhs, K_e, b_e, At, bt, gdof)


# --------------------------------------------------------------------------
# Dirichlet + apply!  (src/boundary.jl:7-42, 121-175)
# --------------------------------------------------------------------------
def dirichlet(mesh, tab, faceset="boundary", g=lambda x: 0.0):
    """Trace-space Dirichlet, src/boundary.jl:11-42 - restated bug for bug: the accumulator N
    is not reset per dof (:27) and the *cell* quadrature weights are used (:33).  Exactly 0
    for g == 0, which is the only case the hot-path configs use."""
    fset = mesh.facesets[faceset] if isinstance(faceset, str) else faceset
    nt, nfq = tab.nt, tab.nfq
    dofs, vals = [], []
    for face in range(1, mesh.nfaces + 1):
        if face not in fset:
            continue
        assert mesh.faces[face - 1, 3] == 0, f"Face {face} is not in boundary"
        cell = mesh.faces[face - 1, 2] - 1
        lidx = list(mesh.cell_faces[cell]).index(face)
        ori = face_orientation(mesh, cell, lidx)
        coords = mesh.nodes[mesh.cells[cell] - 1]
        acc = 0.0
        for i in range(nt):
            for q in range(nfq):
                qo = q if ori else nfq - 1 - q
                xq = tab.L[:, qo, lidx] @ coords     # src/TraceFunctionSpaces.jl:47-57
                wq = tab.qw[q] if q < len(tab.qw) else 0.0
                acc += wq * g(xq) * tab.T[i, q]
            vals.append(acc)
            dofs.append(face * nt - nt + i + 1)
    return np.array(dofs, dtype=np.int64), np.array(vals, float)


def meandiag(K):
    """src/boundary.jl:169-175."""
    return np.abs(K.diagonal()).sum() / K.shape[0]


def apply_dirichlet(K, rhs, dofs, vals):
    """apply!(K,f,dbc), src/boundary.jl:121-158 (APPLY_TRANSPOSE).  Pattern is unchanged:
    zeroed entries stay stored.  Returns (K, rhs, m)."""
    K = K.copy().tocsc()
    rhs = rhs.copy()
    m = meandiag(K)
    for d, v in zip(dofs - 1, vals):
        if v != 0:
            for p in range(K.indptr[d], K.indptr[d + 1]):
                rhs[K.indices[p]] -= v * K.data[p]
    dset = np.zeros(K.shape[0], bool)
    dset[dofs - 1] = True
    col_of = np.repeat(np.arange(K.shape[1]), np.diff(K.indptr))
    K.data[dset[col_of]] = 0.0          # zero_out_columns!
    K.data[dset[K.indices]] = 0.0       # rows (transpose trick)
    diag_pos = np.flatnonzero(K.indices == col_of)
    dp = {int(col_of[p]): int(p) for p in diag_pos}
    for d, v in zip(dofs - 1, vals):
        K.data[dp[int(d)]] = m
        rhs[d] = v * m
    return K, rhs, m


def solve_direct(K, rhs):
    """u_hat = K \\ b, examples/poisson2D_HDG.jl:195 (UMFPACK LU in the reference - third party,
    absent; SuperLU through scipy is the same mathematical operation)."""
    return spla.spsolve(K.tocsc(), rhs)


# --------------------------------------------------------------------------
# Recovery + error norm (examples/poisson2D_HDG.jl:197-212, src/DiscreteFunctions.jl:97-120)
# --------------------------------------------------------------------------
def recover(mesh, tab, uhat, asm):
    """get_u_sigma! with the hard-coded nt=2 slice (:205-206) generalised to nt.
    Returns m_values arrays: sigma (ncell,2n), u (ncell,n), uhat_h (ncell,nt,3)."""
    n, nt = tab.n, tab.nt
    nc = mesh.ncells
    sig = np.empty((nc, 2 * n)); u = np.empty((nc, n)); uh = np.empty((nc, nt, 3))
    for c in range(nc):
        ue = np.concatenate([uhat[nt * (fi - 1):nt * fi] for fi in mesh.cell_faces[c]])
        for k, fi in enumerate(mesh.cell_faces[c]):
            uh[c, :, k] = uhat[nt * (fi - 1):nt * fi]
        dof = asm.K_e[c] @ ue + asm.b_e[c]
        sig[c] = dof[:2 * n]
        u[c] = dof[2 * n:]
    return sig, u, uh


def errornorm(mesh, tab, u_vals, u_ex=exact_poisson):
    """Squared L2 error, src/DiscreteFunctions.jl:97-120."""
    tot = 0.0
    for c in range(mesh.ncells):
        x = mesh.nodes[mesh.cells[c] - 1]
        g = reinit(tab, x)
        el = 0.0
        for q in range(tab.nq):
            dO = g.detJ * tab.qw[q]
            uq = 0.0
            for i in range(tab.n):
                uq += u_vals[c, i] * tab.N[i, q]
            el += (uq - u_ex(tab.M[:, q] @ x)) ** 2 * dO
        tot += el
    return tot


def nodal_avg(mesh, tab, u_vals):
    """nodal_avg(u_h) with value(u_h,node,cell), src/DiscreteFunctions.jl:70-95."""
    acc = np.zeros(mesh.nnodes)
    cnt = np.zeros(mesh.nnodes, dtype=np.int64)
    for c in range(mesh.ncells):
        x = mesh.nodes[mesh.cells[c] - 1]
        g = reinit(tab, x)
        for node in mesh.cells[c]:
            xi = g.Jinv @ (mesh.nodes[node - 1] - x[0])          # reference_coordinate, src/DiscreteFunctions.jl:63-64
            xi = np.clip(xi, 0.0, 1.0)                            # rounding may leave the reference triangle by 1 ulp
            u = sum(u_vals[c, i] * dubiner_value(i + 1, xi[0], xi[1]) for i in range(tab.n))
            acc[node - 1] += u
            cnt[node - 1] += 1
    return acc / cnt


def run_poisson(mesh, order=1, quad_degree=None, tau=1.0):
    """The whole driver examples/poisson2D_HDG.jl:37-218 on `mesh`."""
    tab = build_tables(order, quad_degree)
    asm = doassemble(mesh, tab, source_poisson, tau)
    dofs, vals = dirichlet(mesh, tab)
    Kb, rb, m = apply_dirichlet(asm.K, asm.rhs, dofs, vals)
    uhat = solve_direct(Kb, rb)
    sig, u, uh = recover(mesh, tab, uhat, asm)
    err2 = errornorm(mesh, tab, u)
    return dict(tab=tab, asm=asm, dofs=dofs, vals=vals, K_bc=Kb, rhs_bc=rb, meandiag=m,
                uhat=uhat, sigma=sig, u=u, uhat_h=uh, err2=err2)


# --------------------------------------------------------------------------
# CG side of the exported API: DofHandler, sparsity pattern, Dirichlet on a DofHandler.
# Integer logic only; loop-faithful to src/dofhandler.jl:80-220 and src/boundary.jl:44-96.
# Pinned on test/test_handlers.jl:13-19 (tests/test_oracle_goldens.py).
# --------------------------------------------------------------------------
def lagrange_topology(order):
    """gettopology(ContinuousLagrange{2,RefTetrahedron,order}) via get_nodal_points, src/shapes.jl:46-57:
    dofs on the 3 vertices, on the 3 edges together, in the interior."""
    return {0: 3, 1: 3 * (order - 1), 2: (order - 1) * (order - 2) // 2}


def distribute_dofs(mesh, order=1, ncomponents=1):
    """_distribute_dofs for ONE field, src/dofhandler.jl:84-152.  Returns (cell_dofs, cell_dofs_offset), 1-based.
    Restated as written: when an edge with several dofs is met again only `ncomponents` dofs starting at the stored
    first dof are pushed (:121-124), which is what the reference does for order >= 3."""
    topo = lagrange_topology(order)
    geo = {0: 3, 1: 3, 2: 1}                      # gettopology(RefTetrahedron, Val{2}), src/shapes.jl:26-28
    dicts = {0: {}, 1: {}}
    cell_dofs, offsets = [], [1]
    nextdof = 1
    for c in range(mesh.ncells):
        for n_el in (0, 1):
            assert topo[n_el] % geo[n_el] == 0
            nelementdofs = topo[n_el] // geo[n_el]
            if nelementdofs == 0:
                continue
            elements = mesh.cells[c] if n_el == 0 else mesh.cell_faces[c]     # topology_elements, src/mesh.jl:33-41
            for el in elements:
                el = int(el)
                if el in dicts[n_el]:
                    reuse = dicts[n_el][el]
                    for d in range(ncomponents):
                        cell_dofs.append(reuse + d)
                else:
                    for _ in range(nelementdofs):
                        dicts[n_el][el] = nextdof          # _setindex! overwrites: the LAST first-dof is stored (:127)
                        for d in range(ncomponents):
                            cell_dofs.append(nextdof)
                            nextdof += 1
        for _ in range(topo[2]):
            for d in range(ncomponents):
                cell_dofs.append(nextdof)
                nextdof += 1
        offsets.append(len(cell_dofs) + 1)
    return np.array(cell_dofs, dtype=np.int64), np.array(offsets, dtype=np.int64)


def create_sparsity_pattern(cell_dofs, offsets):
    """_create_sparsity_pattern(dh, false), src/dofhandler.jl:180-216: CSC pattern (colptr, rowval; 1-based) of
    sparse(I, J, zeros) with all element couplings plus the diagonal."""
    ncell = offsets.size - 1
    n = int(offsets[1] - offsets[0])
    ndofs = int(cell_dofs.max())
    I, J = [], []
    for e in range(ncell):
        g = cell_dofs[offsets[e] - 1: offsets[e] - 1 + n]
        for j in range(n):
            for i in range(n):
                I.append(g[i]); J.append(g[j])
    for d in range(1, ndofs + 1):
        I.append(d); J.append(d)
    K = sp.coo_matrix((np.zeros(len(I)), (np.array(I) - 1, np.array(J) - 1)), shape=(ndofs, ndofs)).tocsc()
    K.sum_duplicates()
    K.sort_indices()
    return K.indptr.astype(np.int64) + 1, K.indices.astype(np.int64) + 1


def dirichlet_dofhandler(mesh, cell_dofs, offsets, order=1, faceset="boundary"):
    """prescribed_dofs of Dirichlet(u, dh, faceset, f), src/boundary.jl:48-96 (sorted; one scalar field)."""
    fset = mesh.facesets[faceset] if isinstance(faceset, str) else faceset
    topo = lagrange_topology(order)
    nel = [topo[0] // 3, topo[1] // 3]
    edge_nodes = ((2, 3), (3, 1), (1, 2))          # reference_edge_nodes, src/shapes.jl (local faces = node pairs)
    prescribed = []
    for face in range(1, mesh.nfaces + 1):
        if face not in fset:
            continue
        assert mesh.faces[face - 1, 3] == 0, f"Face {face} is not in boundary"
        cell = int(mesh.faces[face - 1, 2]) - 1
        lidx = list(mesh.cell_faces[cell]).index(face) + 1
        off = int(offsets[cell]) - 1
        for j in range(2):
            lo = edge_nodes[lidx - 1][j]
            d = int(cell_dofs[off + lo - 1])
            if d not in prescribed:
                prescribed.append(d)
        for j in range(1, nel[1] + 1):
            lo = topo[0] + nel[1] * (lidx - 1) + j
            d = int(cell_dofs[off + lo - 1])
            if d not in prescribed:
                prescribed.append(d)
    return np.array(sorted(prescribed), dtype=np.int64)


# --------------------------------------------------------------------------
# CG side: examples/poisson2D_CG.jl (ContinuousLagrange{2,RefTetrahedron,order}, one scalar field)
# --------------------------------------------------------------------------
def lagrange_nodal_points(order):
    """get_nodal_points(RefTetrahedron, Val{2}, order), src/shapes.jl:46-57 for order 1, 2: the vertices, then the interior
    points of the reference edges ((1,0)-(0,1), (0,1)-(0,0), (0,0)-(1,0), src/shapes.jl:19-23)."""
    v = [np.array([0.0, 0.0]), np.array([1.0, 0.0]), np.array([0.0, 1.0])]
    pts = list(v)
    edges = [(v[1], v[2]), (v[2], v[0]), (v[0], v[1])]
    for a, b in edges:
        for i in range(1, order):
            pts.append(a + i * (b - a) / order)
    assert order <= 2
    return np.array(pts)


def lagrange_tables(order, quad_degree=None):
    """Lagrange{2,RefTetrahedron,order} as the reference builds it (src/basis.jl:264-293): nodal_base_coefs = inv(V),
    V[i,j] = Dubiner_j(nodal point i); value(ip,k,xi) = nodal_base_coefs[:,k] . Dubiner(xi).  Tables of
    ScalarFunctionSpace(mesh, ContinuousLagrange; quad_degree = order + 1) (src/ScalarFunctionSpaces.jl:24-99)."""
    qd = quad_degree or order + 1
    pts, w = default_quad_2d(qd)
    n = (order + 1) * (order + 2) // 2
    nodal = lagrange_nodal_points(order)
    V = np.array([[dubiner_value(j + 1, p[0], p[1]) for j in range(n)] for p in nodal])
    coefs = np.linalg.inv(V)
    nq = len(w)
    N = np.empty((n, nq)); dN = np.empty((n, nq, 2))
    for q in range(nq):
        d = np.array([dubiner_value(j + 1, pts[q][0], pts[q][1]) for j in range(n)])
        g = np.array([dubiner_grad(j + 1, pts[q][0], pts[q][1]) for j in range(n)])      # (n, 2)
        for k in range(n):
            N[k, q] = coefs[:, k] @ d
            dN[k, q] = coefs[:, k] @ g
    M = np.array([[1 - p[0] - p[1], p[0], p[1]] for p in pts]).T      # geometry, (3, nq)
    return dict(order=order, n=n, nq=nq, N=N, dN=dN, qw=np.asarray(w, float), M=M, qp=np.asarray(pts, float))


def cg_doassemble(mesh, order=1, f=source_poisson):
    """doassemble(Wh, K, dh) of examples/poisson2D_CG.jl:72-126 with assemble!(assembler, dofs, fe, Ke)
    (src/assembler.jl:62-137): returns (K csc with the pattern of create_sparsity_pattern, b, cell_dofs, offsets, tables)."""
    tab = lagrange_tables(order)
    n, nq = tab["n"], tab["nq"]
    cell_dofs, offsets = distribute_dofs(mesh, order)
    colptr, rowval = create_sparsity_pattern(cell_dofs, offsets)
    ndofs = colptr.size - 1
    K = sp.csc_matrix((np.zeros(rowval.size), rowval - 1, colptr - 1), shape=(ndofs, ndofs))
    b = np.zeros(ndofs)
    for c in range(mesh.ncells):
        x = mesh.nodes[mesh.cells[c] - 1]
        J = np.array([x[1] - x[0], x[2] - x[0]])            # rows: d x / d xi_r   (reinit!, src/ScalarFunctionSpaces.jl:101-132)
        detJ = J[0, 0] * J[1, 1] - J[0, 1] * J[1, 0]
        Jinv = np.linalg.inv(J)
        Ke = np.zeros((n, n)); fe = np.zeros(n)
        for q in range(nq):
            dO = detJ * tab["qw"][q]
            xq = tab["M"][:, q] @ x
            fh = f(xq)
            grads = tab["dN"][:, q, :] @ Jinv.T                 # dNdx = dNdxi . Jinv
            for i in range(n):
                fe[i] += fh * tab["N"][i, q] * dO
                for j in range(n):
                    Ke[i, j] += (grads[i] @ grads[j]) * dO
        g = cell_dofs[offsets[c] - 1: offsets[c] - 1 + n] - 1
        for j in range(n):
            for i in range(n):
                # _assemble!: the stored entry (row g[i], column g[j])
                lo, hi = K.indptr[g[j]], K.indptr[g[j] + 1]
                p = lo + np.searchsorted(K.indices[lo:hi], g[i])
                K.data[p] += Ke[i, j]
            b[g[j]] += fe[j]
    return K, b, cell_dofs, offsets, tab


def cg_errornorm(mesh, tab, cell_dofs, offsets, u, u_ex=exact_poisson):
    """reconstruct!(u_h, u, dh) (src/dofhandler.jl:218-228) + errornorm(u_h, u_ex) (src/DiscreteFunctions.jl:97-120)."""
    n, nq = tab["n"], tab["nq"]
    tot = 0.0
    for c in range(mesh.ncells):
        x = mesh.nodes[mesh.cells[c] - 1]
        detJ = (x[1, 0] - x[0, 0]) * (x[2, 1] - x[0, 1]) - (x[2, 0] - x[0, 0]) * (x[1, 1] - x[0, 1])
        uc = u[cell_dofs[offsets[c] - 1: offsets[c] - 1 + n] - 1]
        for q in range(nq):
            uq = uc @ tab["N"][:, q]
            tot += (uq - u_ex(tab["M"][:, q] @ x)) ** 2 * detJ * tab["qw"][q]
    return tot


def run_poisson_cg(mesh, order=1):
    """examples/poisson2D_CG.jl end to end: assemble, Dirichlet(u_h, dh, "boundary", [0.0]), apply!, K \\ b, errornorm."""
    K, b, cell_dofs, offsets, tab = cg_doassemble(mesh, order)
    dofs = dirichlet_dofhandler(mesh, cell_dofs, offsets, order, "boundary")
    K2, b2, m = apply_dirichlet(K, b, dofs, np.zeros(dofs.size))
    u = spla.spsolve(K2.tocsc(), b2)
    return dict(K=K, b=b, cell_dofs=cell_dofs, offsets=offsets, dofs=dofs, K_bc=K2, b_bc=b2, meandiag=m, u=u,
                err2=cg_errornorm(mesh, tab, cell_dofs, offsets, u), tab=tab)
