"""Pins the CPU oracle (oracle/hdg_oracle.py) on every golden vector / known-answer test the
reference's own test-suite holds for the HDG path.  CPU only."""
import math
import os

import numpy as np
import pytest

import hdg_oracle as orc
from fixtures_util import triangle_root

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sq2, sq3, sq6 = math.sqrt(2), math.sqrt(3), math.sqrt(6)


@pytest.fixture(scope="module")
def fig21():
    return orc.parse_mesh_triangle(triangle_root("figure2.1"))


# ---- test/test_mesh.jl -----------------------------------------------------------------------
def test_triangle_mesh_goldens(fig21):
    m = fig21
    assert m.ncells == 4 and m.nnodes == 5                                   # :11-12
    assert m.cells.tolist() == [[2, 3, 5], [4, 1, 5], [5, 3, 4], [1, 2, 5]]   # :17
    assert m.nodes.tolist() == [[0, 0], [1, 0], [1, 1], [0, 1], [.5, .5]]     # :18
    assert tuple(m.cell_faces[0]) == (1, 2, 3)                                # :21
    assert [orc.face_orientation(m, 0, i) for i in range(3)] == [True, False, True]   # :22
    assert m.facesets["boundary"] == {3, 6, 7, 8}                             # :24
    # get_coordinates(1, mesh) == [(1,1),(0.5,0.5)]  :25
    assert m.nodes[m.faces[0, :2] - 1].tolist() == [[1.0, 1.0], [0.5, 0.5]]
    for c in range(4):                                                        # volume == 1/4  :14-16
        x = m.nodes[m.cells[c] - 1]
        assert abs(orc.reinit(orc.build_tables(1), x).detJ / 2 - 0.25) < 1e-15


def test_rectangle_mesh_goldens():
    m = orc.rectangle_mesh(2, 2)                                              # :31-44
    assert m.ncells == 8 and m.nnodes == 9
    assert m.cells.tolist() == [[1, 2, 4], [2, 5, 4], [2, 3, 5], [3, 6, 5], [4, 5, 7], [5, 8, 7], [5, 6, 8], [6, 9, 8]]
    assert m.nodes.tolist() == [[0, 0], [.5, 0], [1, 0], [0, .5], [.5, .5], [1, .5], [0, 1], [.5, 1], [1, 1]]
    assert tuple(m.cells[0]) == (1, 2, 4) and tuple(m.cell_faces[0]) == (1, 2, 3)
    assert [orc.face_orientation(m, 0, i) for i in range(3)] == [True, False, True]
    assert m.facesets["boundary"] == {3, 7, 9, 16, 2, 11, 12, 15}
    assert m.nfaces == 2 + 2 + 3 * 2 * 2                                      # src/generate_mesh.jl:121


@pytest.mark.parametrize("nx,ny", [(1, 1), (3, 2), (2, 5), (6, 6)])
def test_rectangle_closed_form_numbering(nx, ny):
    """SURVEY Appendix B closed forms == the sequential first-encounter algorithm."""
    m = orc.rectangle_mesh(nx, ny)
    na = lambda i, j: i + (j - 1) * (nx + 1)
    for j in range(1, ny + 1):
        for i in range(1, nx + 1):
            q = (j - 1) * nx + i
            base = 4 * (i - 1) + (i > 1) if j == 1 else 4 * nx + 1 + (j - 2) * (3 * nx + 1) + 3 * (i - 1) + (i > 1)
            diag = base + 1
            top = base + 2 + (i == 1) + (j == 1)
            right = base + 3 + (i == 1) + (j == 1)
            assert m.cell_faces[2 * q - 2, 0] == diag and m.cell_faces[2 * q - 1, 1] == diag
            assert m.cell_faces[2 * q - 1, 0] == top and m.cell_faces[2 * q - 1, 2] == right
            assert tuple(m.faces[diag - 1]) == (na(i + 1, j), na(i, j + 1), 2 * q - 1, 2 * q)


# ---- test/test_quadrature.jl -------------------------------------------------------------------
def test_quadrature_goldens():
    p, w = orc.grundmann_moeller(0)
    assert np.allclose(w, [0.5]) and np.allclose(p[0], [1 / 3, 1 / 3])         # :6-8
    p, w = orc.grundmann_moeller(1)
    assert np.allclose(w, 0.5 * np.array([0.520833333333333, 0.520833333333333, 0.520833333333333, -0.5625]))   # :10
    assert np.allclose(p, [[1 / 5, 1 / 5], [3 / 5, 1 / 5], [1 / 5, 3 / 5], [1 / 3, 1 / 3]])                      # :11-14
    for s in range(6):
        assert abs(orc.grundmann_moeller(s)[1].sum() - 0.5) < 1e-13           # :16-19
    for d in range(1, 7):
        assert abs(orc.strang(d)[1].sum() - 0.5) < 1e-14                      # :22-25
    assert [len(orc.default_quad_2d(d)[1]) for d in (2, 3, 4, 5, 6, 9)] == [3, 6, 6, 7, 12, 35]
    with pytest.raises(ValueError):
        orc.default_quad_2d(8)                                                # src/quadrature.jl:24


# ---- test/test_basis.jl ------------------------------------------------------------------------
def _dubiner_closed(j, r, s):
    """Closed forms of src/basis.jl:65-86 typed independently (sympy-checked against the recursion)."""
    a = 2 * r + s - 1
    q2 = 6 * r * r + 6 * r * (s - 1) + s * s - 2 * s + 1
    q3 = 10 * r * r + 10 * r * (s - 1) + s * s - 2 * s + 1
    f = {1: math.sqrt(2), 2: 2 * sq3 * a, 3: 2 * (3 * s - 1), 4: math.sqrt(30) * q2, 5: 3 * sq2 * (5 * s - 1) * a,
         6: sq6 * (10 * s * s - 8 * s + 1), 7: 2 * math.sqrt(14) * a * q3, 8: 2 * math.sqrt(10) * (7 * s - 1) * q2,
         9: 2 * sq6 * (21 * s * s - 12 * s + 1) * a, 10: 2 * sq2 * (35 * s ** 3 - 45 * s * s + 15 * s - 1),
         12: math.sqrt(70) * (9 * s - 1) * a * q3, 13: 5 * sq2 * (36 * s * s - 16 * s + 1) * q2,
         14: math.sqrt(30) * (84 * s ** 3 - 84 * s * s + 21 * s - 1) * a,
         15: math.sqrt(10) * (126 * s ** 4 - 224 * s ** 3 + 126 * s * s - 24 * s + 1)}
    return f[j]


def test_dubiner_closed_forms_equal_recursion():
    pts, _ = orc.strang(5)                                                    # test/test_basis.jl:6-12
    for j in [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 13, 14, 15]:
        for r, s in pts:
            assert abs(orc.dubiner_value(j, r, s) - _dubiner_closed(j, r, s)) < 2e-13
    # gradients: central differences of the closed form
    h = 1e-6
    for j in [2, 3, 4, 5, 6, 7, 8, 9, 10]:
        for r, s in pts:
            gx = (_dubiner_closed(j, r + h, s) - _dubiner_closed(j, r - h, s)) / (2 * h)
            gy = (_dubiner_closed(j, r, s + h) - _dubiner_closed(j, r, s - h)) / (2 * h)
            g = orc.dubiner_grad(j, r, s)
            assert abs(g[0] - gx) < 1e-6 * max(1, abs(gx)) and abs(g[1] - gy) < 1e-6 * max(1, abs(gy))


def test_dubiner_orthonormal():
    pts, w = orc.default_quad_2d(9)
    V = np.array([[orc.dubiner_value(j, r, s) for (r, s) in pts] for j in range(1, 16)])
    assert np.abs((V * w) @ V.T - np.eye(15)).max() < 1e-12


def test_legendre_goldens():
    ref = [lambda x: 1.0, lambda x: sq3 * (2 * x - 1), lambda x: math.sqrt(5) * (6 * x * x - 6 * x + 1),
           lambda x: math.sqrt(7) * (2 * x - 1) * (10 * x * x - 10 * x + 1)]   # test/test_basis.jl:62-93
    xs, _ = orc.gauss_legendre_01(3)
    for i in range(4):
        for x in xs:
            assert abs(orc.legendre_value(i + 1, x) - ref[i](x)) < 1e-14


# ---- test/test_ScalarFuncSp.jl -------------------------------------------------------------------
def test_reinit_goldens(fig21):
    tab = orc.build_tables(1)
    invs = [[[1, 1], [-2, 0]], [[-1, -1], [2, 0]], [[1, 1], [-1, 1]], [[1, -1], [0, 2]]]      # :11-14
    dets = [[sq2 / 2, sq2 / 2, 1], [sq2 / 2, sq2 / 2, 1], [1, sq2 / 2, sq2 / 2], [sq2 / 2, sq2 / 2, 1]]   # :15-18
    for c in range(4):
        g = orc.reinit(tab, fig21.nodes[fig21.cells[c] - 1])
        assert abs(g.detJ - 0.5) < 1e-15                                      # :26
        assert np.allclose(g.Jinv, invs[c], atol=1e-14)                       # :28
        assert np.allclose(g.detJf, dets[c], atol=1e-14)                      # :30
        if c == 0:
            assert np.allclose(g.normals, [[-sq2 / 2, sq2 / 2], [-sq2 / 2, -sq2 / 2], [1.0, 0.0]], atol=1e-14)   # :32


# ---- test/test_FunctionSpace.jl --------------------------------------------------------------------
Be_ex = [
    [[0, 0, 0], [0, 0, 0], [-3 * sq2, 0, 0], [0, 0, 0], [sq6, 0, 0], [0, 0, 0]],
    [[0, 0, 0], [0, 0, 0], [3 * sq2, 0, 0], [0, 0, 0], [-sq6, 0, 0], [0, 0, 0]],
    [[0, 0, 0], [sq6 / 2, 0, 0], [-1.5 * sq2, 0, 0], [0, 0, 0], [1.5 * sq6, 0, 0], [1.5 * sq2, 0, 0]],
    [[0, 0, 0], [sq6, 0, 0], [0, 0, 0], [0, 0, 0], [0, 0, 0], [3 * sq2, 0, 0]],
]
_C0 = [[2 * sq2 + 2, 0, 2 - 2 * sq2], [0, 4 * sq2 + 4, 0], [2 - 2 * sq2, 0, 4 * sq2 + 4]]
Ce_ex = [_C0, _C0, [[2 * sq2 + 2, sq6 - sq3, sq2 - 1], [sq6 - sq3, 4 * sq2 + 4, 0], [sq2 - 1, 0, 4 * sq2 + 4]], _C0]
h = 0.5
Ee_ex = [
    -np.array([[sq2 / 2, 0, sq2 / 2, 0, -sq2, 0], [sq3 / 2, -h, -sq3 / 2, h, 0, -2], [h, sq3 / 2, h, sq3 / 2, 2, 0],
               [-sq2 / 2, 0, sq2 / 2, 0, 0, 0], [-sq3 / 2, h, -sq3 / 2, h, 0, 0], [-h, -sq3 / 2, h, sq3 / 2, 0, 0]]),
    -np.array([[-sq2 / 2, 0, -sq2 / 2, 0, sq2, 0], [-sq3 / 2, h, sq3 / 2, -h, 0, -2], [-h, -sq3 / 2, -h, -sq3 / 2, -2, 0],
               [sq2 / 2, 0, -sq2 / 2, 0, 0, 0], [sq3 / 2, -h, sq3 / 2, -h, 0, 0], [h, sq3 / 2, -h, -sq3 / 2, 0, 0]]),
    -np.array([[0, 0, sq2 / 2, 0, -sq2 / 2, 0], [0, 0, -sq3 / 2, -h, 0, 1], [0, 0, h, -sq3 / 2, 1, 0],
               [-sq2, 0, sq2 / 2, 0, sq2 / 2, 0], [-sq3, 1, -sq3 / 2, -h, 0, -1], [-1, -sq3, h, -sq3 / 2, -1, 0]]),
    -np.array([[-sq2 / 2, 0, sq2 / 2, 0, 0, 0], [-sq3 / 2, h, -sq3 / 2, h, 0, 0], [-h, -sq3 / 2, h, sq3 / 2, 0, 0],
               [-sq2 / 2, 0, -sq2 / 2, 0, sq2, 0], [-sq3 / 2, h, sq3 / 2, -h, 0, 2], [-h, -sq3 / 2, -h, -sq3 / 2, -2, 0]]),
]
_H0 = np.diag([sq2 / 2, sq2 / 2, sq2 / 2, sq2 / 2, 1, 1])
He_ex = [_H0, _H0, np.diag([1, 1, sq2 / 2, sq2 / 2, sq2 / 2, sq2 / 2]), _H0]


def test_local_block_goldens(fig21):
    """Ae ~ 0.5 I, Be, Ce, Ee, He for the 4 fixture cells: test/test_FunctionSpace.jl:49-72,125-126,176-178."""
    tab = orc.build_tables(1)
    assert (2 * tab.n, tab.n, 3 * tab.nt) == (6, 3, 6)                         # :30-32
    ori = orc.orientations(fig21)
    for c in range(4):
        blk = orc.local_blocks(tab, fig21.nodes[fig21.cells[c] - 1], ori[c])
        assert np.allclose(blk["A"], 0.5 * np.eye(6), atol=1e-14)
        assert np.allclose(blk["B"], Be_ex[c], atol=1e-13)
        assert np.allclose(blk["C"], Ce_ex[c], atol=1e-13)
        assert np.allclose(blk["E"], Ee_ex[c], atol=1e-13)
        assert np.allclose(blk["H"], He_ex[c], atol=1e-13)
    g = orc.reinit(tab, fig21.nodes[fig21.cells[3] - 1])                       # state after the loop :33-41
    assert np.allclose(g.detJf * tab.fw[0], [sq2 / 4, sq2 / 4, 0.5])


def test_end_to_end_error_bounds(fig21):
    r = orc.run_poisson(fig21, 1)
    assert r["err2"] <= 0.12                                                  # test/test_FunctionSpace.jl:243
    assert r["dofs"].tolist() == [5, 6, 11, 12, 13, 14, 15, 16]
    r = orc.run_poisson(orc.rectangle_mesh(10, 10), 1)
    assert r["err2"] <= 0.00006                                               # examples/poisson2D_HDG.jl:218
    assert r["asm"].K.shape == (640, 640) and r["asm"].K.nnz == 6080
    # regression values of the restatement recorded in SURVEY.md section 6 / BASELINE.md section 1
    assert abs(r["meandiag"] - 9.485018480631883) < 1e-12
    assert abs(np.abs(r["asm"].K).sum() - 15917.863160055884) < 1e-8
    assert abs(r["asm"].rhs.sum() - (-7.999958742811353)) < 1e-11
    assert abs(np.linalg.norm(r["uhat"]) - 8.658950189552586) < 1e-11
    assert abs(r["uhat"][0] - 1.615527554721488e-02) < 1e-13
    assert abs(r["err2"] - 5.364546646411725e-05) < 1e-15
    # the condensed matrix is symmetric negative semi-definite; apply! makes it indefinite
    K = r["asm"].K.toarray()
    assert np.abs(K - K.T).max() < 1e-13
    ev = np.linalg.eigvalsh((K + K.T) / 2)
    assert ev.max() < 1e-10 and ev.min() < -20


def test_convergence_rates():
    """L2 error converges at the optimal rate h^(k+1) (err^2 ratio 2^(2k+2) per halving)."""
    for k, qd, lo in ((1, 2, 12.0), (2, 4, 50.0)):
        e = [orc.run_poisson(orc.rectangle_mesh(n, n), k, qd)["err2"] for n in (4, 8)]
        assert e[0] / e[1] > lo


def test_dofhandler_goldens():
    """test/test_handlers.jl:13-19: P1 DofHandler on rectangle_mesh(TriangleCell,(2,2)) and its Dirichlet dofs."""
    mo = orc.rectangle_mesh(2, 2)
    cell_dofs, offsets = orc.distribute_dofs(mo, 1)
    assert cell_dofs.tolist() == [1, 2, 3, 2, 4, 3, 2, 5, 4, 5, 6, 4, 3, 4, 7, 4, 8, 7, 4, 6, 8, 6, 9, 8]
    assert offsets.tolist() == [1, 4, 7, 10, 13, 16, 19, 22, 25]
    assert orc.dirichlet_dofhandler(mo, cell_dofs, offsets).tolist() == [1, 2, 3, 5, 6, 7, 8, 9]
    colptr, rowval = orc.create_sparsity_pattern(cell_dofs, offsets)
    assert colptr[-1] - 1 == rowval.size == 9 + 2 * 16      # 9 vertices + 16 edges, both directions


def test_condensation_against_extended_precision_on_the_golden_blocks(fig21):
    """K_e, b_e, Ate, bte have no golden in the reference.  What can be pinned: the condensation
    (examples/poisson2D_HDG.jl:155-174) applied to the reference's GOLDEN blocks A = I/2, Be, Ce, Ee, He
    (test/test_FunctionSpace.jl:49-72) in 40-digit arithmetic must give the oracle's K_e and Ate; only Fe and be (no golden)
    are taken from the restatement itself."""
    import mpmath as mp
    mp.mp.dps = 40
    tab = orc.build_tables(1)
    ori = orc.orientations(fig21)
    for c in range(4):
        blk = orc.local_blocks(tab, fig21.nodes[fig21.cells[c] - 1], ori[c])
        K_e, b_e, At, bt = orc.condense(blk)
        A = 0.5 * np.eye(6)
        B, Cm, E, H = (np.asarray(M, dtype=float) for M in (Be_ex[c], Ce_ex[c], Ee_ex[c], He_ex[c]))
        F, be = blk["F"], blk["be"]
        Me = mp.matrix(np.block([[A, -B], [B.T, Cm]]).tolist())                # doubles in, 40-digit arithmetic from here
        EF = mp.matrix(np.vstack([-E, F]).tolist())
        G = mp.matrix(np.vstack([E, F]).tolist())
        cols = [mp.lu_solve(Me, EF[:, j]) for j in range(6)]                   # lu_solve takes one right-hand side
        Kx = mp.matrix(9, 6)
        for j in range(6):
            for i in range(9):
                Kx[i, j] = cols[j][i]
        Atx = G.T * Kx - mp.matrix(H.tolist())
        bx = mp.lu_solve(Me, mp.matrix(np.concatenate([np.zeros(6), be]).reshape(-1, 1).tolist()))
        btx = -(G.T * bx)
        tonp = lambda M: np.array([[float(M[i, j]) for j in range(M.cols)] for i in range(M.rows)])
        assert np.abs(tonp(Kx) - K_e).max() < 5e-13 * np.abs(K_e).max()
        assert np.abs(tonp(Atx) - At).max() < 5e-13 * np.abs(At).max()
        assert np.abs(tonp(bx)[:, 0] - b_e).max() < 5e-13 * np.abs(b_e).max()
        assert np.abs(tonp(btx)[:, 0] - bt).max() < 5e-13 * np.abs(bt).max()
        assert np.abs(At - At.T).max() < 1e-13                                 # symmetric to rounding
