"""Host-side logic of the Python mirror and the C port of the CPU path, checked against the numpy
oracle.  CPU only."""
import os

import numpy as np
import pytest

import hdg_b200 as hdg
import hdg_oracle as orc
from fixtures_util import triangle_root
import hdg_oracle_c as occ

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("root", ["figure2.1", "figure.1"])
def test_parse_mesh_triangle_matches_oracle(root):
    m = hdg.parse_mesh_triangle(triangle_root(root))
    mo = orc.parse_mesh_triangle(triangle_root(root))
    assert np.array_equal(m.cells[:, :3], mo.cells) and np.array_equal(m.cells[:, 3:], mo.cell_faces)
    assert np.array_equal(m.faces, mo.faces) and np.array_equal(m.nodes, mo.nodes)
    assert m.facesets["boundary"] == mo.facesets["boundary"]
    assert m.faces.flags["F_CONTIGUOUS"] and m.cells.flags["C_CONTIGUOUS"]   # Julia layouts for the C ABI


@pytest.mark.parametrize("nx,ny", [(1, 1), (4, 3), (9, 2), (16, 16)])
def test_sort_based_face_numbering_equals_sequential(nx, ny):
    mo = orc.rectangle_mesh(nx, ny)
    cf, fa = hdg.number_faces(mo.cells)
    assert np.array_equal(cf, mo.cell_faces) and np.array_equal(fa, mo.faces)


def test_face_numbering_random_permutation():
    rng = np.random.default_rng(3)
    mo = orc.rectangle_mesh(7, 6)
    perm = rng.permutation(mo.ncells)
    tri = mo.cells[perm]
    cells, cell_faces, faces = orc._build_cells_sequential([tuple(t) for t in tri], mo.nodes, mo.nfaces)
    cf, fa = hdg.number_faces(tri)
    assert np.array_equal(cf, cell_faces) and np.array_equal(fa, faces)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_face_numbering_random_delaunay(seed):
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    pts = rng.random((60, 2))
    tri = Delaunay(pts).simplices.astype(np.int64) + 1
    ccw = np.array([orc._check_node_data(pts, *t) for t in tri], dtype=np.int64)
    cells, cell_faces, faces = orc._build_cells_sequential([tuple(t) for t in ccw], pts, 3 * len(ccw))
    cf, fa = hdg.number_faces(ccw)
    assert np.array_equal(cf, cell_faces) and np.array_equal(fa, faces)
    # Euler: V - E + F = 1 for a triangulated disc
    assert pts.shape[0] - fa.shape[0] + ccw.shape[0] == 1


def test_non_manifold_rejected():
    with pytest.raises(ValueError):
        hdg.number_faces(np.array([[1, 2, 3], [2, 1, 4], [1, 2, 5]]))


def test_dirichlet_dofs_and_values():
    mo = orc.parse_mesh_triangle(triangle_root("figure2.1"))
    m = hdg.parse_mesh_triangle(triangle_root("figure2.1"))
    for order in (1, 2, 3):
        fe = hdg.GenericFiniteElement(hdg.Dubiner(2, hdg.RefTetrahedron, order))
        Wh = hdg.ScalarFunctionSpace(m, fe, 2 * order)
        Mh = hdg.ScalarTraceFunctionSpace(Wh, hdg.GenericFiniteElement(hdg.Legendre(1, hdg.RefTetrahedron, order)))
        d = hdg.Dirichlet(hdg.TrialFunction(Mh), m, "boundary", lambda x: 0)
        dofs, vals = orc.dirichlet(mo, orc.build_tables(order, 2 * order))
        assert np.array_equal(d.prescribed_dofs, dofs) and np.array_equal(d.values, vals)
    assert hdg.getnlocaldofs(Mh) == 3 * (order + 1)
    d1 = hdg.Dirichlet(hdg.TrialFunction(Mh), m, "boundary", lambda x: 1.0)      # g != 0: the reference's projection, restated
    dofs1, vals1 = orc.dirichlet(mo, orc.build_tables(order, 2 * order), "boundary", lambda x: 1.0)
    assert np.array_equal(d1.prescribed_dofs, dofs1) and np.abs(d1.values - vals1).max() < 1e-13


def test_trial_function_storage():
    m = hdg.parse_mesh_triangle(triangle_root("figure2.1"))
    fe = hdg.GenericFiniteElement(hdg.Dubiner(2, hdg.RefTetrahedron, 1))
    Wh, Vh = hdg.ScalarFunctionSpace(m, fe), hdg.VectorFunctionSpace(m, fe)
    Mh = hdg.ScalarTraceFunctionSpace(Wh, hdg.GenericFiniteElement(hdg.Legendre(1, hdg.RefTetrahedron, 1)))
    assert (hdg.getnlocaldofs(Vh), hdg.getnlocaldofs(Wh), hdg.getnlocaldofs(Mh)) == (6, 3, 6)   # test_FunctionSpace.jl:30-32
    assert hdg.TrialFunction(Wh).m_values.shape == (4, 3) and hdg.TrialFunction(Vh).m_values.shape == (4, 6)
    u = hdg.TrialFunction(Mh)
    assert u.m_values.shape == (4, 2, 3) and np.all(np.isnan(u.m_values)) and u.m_values.flags["F_CONTIGUOUS"]


@pytest.mark.parametrize("order,qd,nx", [(1, 2, 8), (2, 3, 5), (2, 4, 5), (3, 6, 4), (4, 9, 2)])
def test_c_port_matches_numpy_oracle(order, qd, nx):
    mesh = orc.rectangle_mesh(nx, nx + 1, (0, 0), (2.0, 1.0))
    tab = orc.build_tables(order, qd)
    asm = orc.doassemble(mesh, tab)
    for nth in (1, 3):
        K, rhs, Ke, be = occ.doassemble(mesh, tab, nthreads=nth)
        assert np.array_equal(K.indptr, asm.K.indptr) and np.array_equal(K.indices, asm.K.indices)
        assert np.allclose(K.data, asm.K.data, rtol=1e-12, atol=1e-13)
        assert np.allclose(rhs, asm.rhs, rtol=1e-12, atol=1e-14)
        assert np.allclose(Ke, asm.K_e, rtol=1e-11, atol=1e-13) and np.allclose(be, asm.b_e, rtol=1e-11, atol=1e-13)


def test_c_pcg_matches_direct_solve():
    mesh = orc.rectangle_mesh(12, 12)
    tab = orc.build_tables(2, 4)
    asm = orc.doassemble(mesh, tab)
    dofs, vals = orc.dirichlet(mesh, tab)
    Kb, rb, m = orc.apply_dirichlet(asm.K, asm.rhs, dofs, vals)
    isbc = np.zeros(Kb.shape[0], np.uint8)
    isbc[dofs - 1] = 1
    x, it, rel = occ.pcg(Kb, rb, isbc, 1e-13, 5000, 2)
    xd = orc.solve_direct(Kb, rb)
    assert rel <= 1e-13 and np.abs(x - xd).max() < 1e-11 * np.abs(xd).max()


@pytest.mark.parametrize("order,qd", [(1, 2), (2, 4), (2, 3), (3, 6), (4, 9)])
def test_nonhomogeneous_dirichlet_values_match_the_restated_reference(order, qd):
    """Dirichlet(u_hat, mesh, "boundary", g) of the host mirror == src/boundary.jl:11-42 as restated in the oracle
    (accumulator not reset per dof, cell weights) - host-only code, no device."""
    g = lambda x: 1.0 + x[0] + 2.0 * x[1] * x[1]
    mo = orc.rectangle_mesh(3, 2, (0.0, 0.0), (2.0, 1.0))
    tab = orc.build_tables(order, qd)
    dofs, vals = orc.dirichlet(mo, tab, "boundary", g)
    mesh = hdg.PolygonalMesh(np.hstack([mo.cells, mo.cell_faces]), mo.nodes, mo.faces, {k: set(v) for k, v in mo.facesets.items()})
    Wh = hdg.ScalarFunctionSpace(mesh, hdg.GenericFiniteElement(hdg.Dubiner(2, hdg.RefTetrahedron, order)), qd)
    Mh = hdg.ScalarTraceFunctionSpace(Wh, hdg.GenericFiniteElement(hdg.Legendre(1, hdg.RefTetrahedron, order)))
    d = hdg.Dirichlet(hdg.TrialFunction(Mh), mesh, "boundary", g)
    assert np.array_equal(d.prescribed_dofs, dofs)
    assert np.abs(d.values - vals).max() <= 1e-13 * np.abs(vals).max()
    # corrected variant: L2 projection on the orthonormal Legendre basis -> mode 0 = face mean of g
    dc = hdg.Dirichlet(hdg.TrialFunction(Mh), mesh, "boundary", lambda x: 3.0, corrected=True)
    assert np.allclose(dc.values.reshape(-1, order + 1)[:, 0], 3.0, atol=1e-13)
    assert np.allclose(dc.values.reshape(-1, order + 1)[:, 1:], 0.0, atol=1e-13)


@pytest.mark.parametrize("order", [1, 2])
def test_cg_dofhandler_host_mirror(order):
    """DofHandler / create_sparsity_pattern / Dirichlet(u, dh, ...) of the host mirror (sort-based first-encounter ranking)
    == the reference's sequential dictionary walk as restated in the oracle; goldens of test/test_handlers.jl:13-19."""
    from fixtures_util import triangle_root
    for mo in (orc.rectangle_mesh(2, 2), orc.rectangle_mesh(7, 5), orc.parse_mesh_triangle(triangle_root("figure.1"))):
        mesh = hdg.PolygonalMesh(np.hstack([mo.cells, mo.cell_faces]), mo.nodes, mo.faces, {k: set(v) for k, v in mo.facesets.items()})
        u = hdg.LagrangeField(hdg.ContinuousLagrange(2, hdg.RefTetrahedron, order), mesh)
        dh = hdg.DofHandler([u], mesh)
        cd, off = orc.distribute_dofs(mo, order)
        assert np.array_equal(dh.cell_dofs, cd) and np.array_equal(dh.cell_dofs_offset, off)
        assert hdg.ndofs(dh) == cd.max() and hdg.ndofs_per_cell(dh) == (3 if order == 1 else 6)
        assert hdg.dof_range(dh, u) == range(1, (3 if order == 1 else 6) + 1)
        cp, rv = orc.create_sparsity_pattern(cd, off)
        cp2, rv2 = hdg.create_sparsity_pattern(dh)
        assert np.array_equal(cp, cp2) and np.array_equal(rv, rv2)
        dbc = hdg.DirichletCG(u, dh, "boundary", lambda x: x[0] + 2.0 * x[1])
        assert np.array_equal(dbc.prescribed_dofs, orc.dirichlet_dofhandler(mo, cd, off, order))
        # values of a linear function at the dof nodes; reconstruct! puts them back per cell
        uvec = np.zeros(hdg.ndofs(dh))
        uvec[dbc.prescribed_dofs - 1] = dbc.values
        hdg.reconstruct_(u, uvec, dh)
        xv = mesh.nodes[mesh.cells[:, :3] - 1]
        onb = np.isin(dh.cell_dofs.reshape(mo.ncells, -1)[:, :3], dbc.prescribed_dofs)
        assert np.allclose(u.m_values[:, :3][onb], (xv[..., 0] + 2.0 * xv[..., 1])[onb], atol=1e-14)
    mo = orc.rectangle_mesh(2, 2)
    mesh = hdg.PolygonalMesh(np.hstack([mo.cells, mo.cell_faces]), mo.nodes, mo.faces, {k: set(v) for k, v in mo.facesets.items()})
    if order == 1:
        dh = hdg.DofHandler([hdg.LagrangeField(hdg.ContinuousLagrange(2, hdg.RefTetrahedron, 1), mesh)], mesh)
        assert dh.cell_dofs.tolist() == [1, 2, 3, 2, 4, 3, 2, 5, 4, 5, 6, 4, 3, 4, 7, 4, 8, 7, 4, 6, 8, 6, 9, 8]
    assert hdg.getcells_matrix(mesh).shape == (8, 3) and hdg.get_vertices_matrix(mesh).shape == (9, 2)


def test_mesh_queries_match_reference_goldens():
    """test/test_mesh.jl:9-24 (figure2.1) and :27-43 (rectangle_mesh 2x2) for the host-side mesh helpers."""
    from fixtures_util import triangle_root
    mesh = hdg.parse_mesh_triangle(triangle_root("figure2.1"))
    assert hdg.getncells(mesh) == 4 and hdg.getnnodes(mesh) == 5 and hdg.n_faces_per_cell(mesh) == 3
    for c in range(1, 5):
        assert abs(hdg.cell_volume(mesh, c) - 0.25) < 1e-15
        assert abs(hdg.volume(hdg.get_coordinates(mesh.cells[c - 1], mesh)) - 0.25) < 1e-15
    assert hdg.getcells_matrix(mesh).tolist() == [[2, 3, 5], [4, 1, 5], [5, 3, 4], [1, 2, 5]]
    assert hdg.get_vertices_matrix(mesh).tolist() == [[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0], [0.5, 0.5]]
    assert mesh.cells[0].tolist() == [2, 3, 5, 1, 2, 3]
    assert [hdg.face_orientation(mesh, 1, i) for i in (1, 2, 3)] == [True, False, True]
    assert hdg.cell_diameter(mesh, 1) == 1.0
    assert hdg.getfaceset(mesh, "boundary") == {3, 6, 7, 8}
    assert hdg.get_coordinates(1, mesh).tolist() == [[1.0, 1.0], [0.5, 0.5]]
    assert np.allclose(hdg.cell_centroid(mesh, 1), mesh.nodes[[1, 2, 4]].mean(axis=0))
    mo = orc.rectangle_mesh(2, 2)
    mesh = hdg.PolygonalMesh(np.hstack([mo.cells, mo.cell_faces]), mo.nodes, mo.faces, {k: set(v) for k, v in mo.facesets.items()})
    for c in range(1, 9):
        assert abs(hdg.cell_volume(mesh, c) - 0.125) < 1e-15
    assert hdg.getcells_matrix(mesh).tolist() == [[1, 2, 4], [2, 5, 4], [2, 3, 5], [3, 6, 5], [4, 5, 7], [5, 8, 7], [5, 6, 8], [6, 9, 8]]
    assert hdg.cell_diameter(mesh, 1) == np.sqrt(2) / 2
    assert hdg.getfaceset(mesh, "boundary") == {3, 7, 9, 16, 2, 11, 12, 15}
    assert hdg.getnodeset(mesh, "boundary") == {1, 2, 3, 4, 6, 7, 8, 9}
    assert hdg.n_nodes_per_cell(mesh) == 3 and hdg.reference_edge_nodes() == ((2, 3), (3, 1), (1, 2))


def test_function_space_accessors_reproduce_the_reference_test_loop():
    """The loop of test/test_FunctionSpace.jl:97-178 and the checks of test/test_ScalarFuncSp.jl:15-32, transcribed onto
    the host mirror's accessors (reinit_, getdetJdV, shape_value, shape_divergence, getfacedetJdS, get_normal).  The
    accessors read the reference tables the device kernels consume (hdg_ref_table), so this pins those tables directly
    on the reference's golden blocks Ae, Be, Ce, Ee, He."""
    from fixtures_util import triangle_root
    from test_oracle_goldens import Be_ex, Ce_ex, Ee_ex, He_ex
    sq2 = np.sqrt(2.0)
    mesh = hdg.parse_mesh_triangle(triangle_root("figure2.1"))
    fe = hdg.GenericFiniteElement(hdg.Dubiner(2, hdg.RefTetrahedron, 1))
    Wh = hdg.ScalarFunctionSpace(mesh, fe)
    Vh = hdg.VectorFunctionSpace(mesh, fe)
    Mh = hdg.ScalarTraceFunctionSpace(Wh, hdg.GenericFiniteElement(hdg.Legendre(1, hdg.RefTetrahedron, 1)))
    assert (hdg.getnlocaldofs(Vh), hdg.getnlocaldofs(Wh), hdg.getnlocaldofs(Mh)) == (6, 3, 6)          # :30-32
    invs = ([[1.0, 1.0], [-2.0, 0.0]], [[-1.0, -1.0], [2.0, 0.0]], [[1.0, 1.0], [-1.0, 1.0]], [[1.0, -1.0], [0.0, 2.0]])
    detsJf = ([sq2 / 2, sq2 / 2, 1], [sq2 / 2, sq2 / 2, 1], [1, sq2 / 2, sq2 / 2], [sq2 / 2, sq2 / 2, 1])
    nv, ns, nt, tau = 6, 3, 2, 1.0
    for c in range(4):
        cell = mesh.cells[c]
        Wh.reinit_(cell)
        Vh.reinit_(cell)
        assert abs(Wh.detJ - 0.5) < 1e-15 and abs(Wh.getdetJdV(1) / Wh._arrays()["qw"][0] - 0.5) < 1e-15   # test_ScalarFuncSp.jl:25-26
        assert np.allclose(Wh.Jinv, invs[c], atol=1e-14) and np.allclose(Wh.detJf, detsJf[c], atol=1e-14)
        if c == 0:
            assert np.allclose(Wh.normals, [[-sq2 / 2, sq2 / 2], [-sq2 / 2, -sq2 / 2], [1.0, 0.0]], atol=1e-14)
        Ae, Be, Ce = np.zeros((nv, nv)), np.zeros((nv, ns)), np.zeros((ns, ns))
        Ee, He = np.zeros((nv, 3 * nt)), np.zeros((3 * nt, 3 * nt))
        for q in range(1, Vh.getnquadpoints() + 1):
            dO = Vh.getdetJdV(q)
            for i in range(1, nv + 1):
                vh, div_vh = Vh.shape_value(q, i), Vh.shape_divergence(q, i)
                for j in range(1, nv + 1):
                    Ae[i - 1, j - 1] += (Vh.shape_value(q, j) @ vh) * dO
                for j in range(1, ns + 1):
                    Be[i - 1, j - 1] += Wh.shape_value(q, j) * div_vh * dO
        for face in (1, 2, 3):
            ori = hdg.face_orientation(mesh, c + 1, face)
            for q in range(1, Wh.getnfacequadpoints() + 1):
                dS = Wh.getfacedetJdS(face, q)
                for i in range(1, ns + 1):
                    w = Wh.shape_value(face, q, i)
                    for j in range(1, ns + 1):
                        Ce[i - 1, j - 1] += tau * Wh.shape_value(face, q, j) * w * dS
                dS = Mh.getfacedetJdS(face, q)
                n = Vh.get_normal(face)
                for i in range(1, nv + 1):
                    v = Vh.shape_value(face, q, i, ori)
                    for j in range(1, nt + 1):
                        Ee[i - 1, nt * (face - 1) + j - 1] += Mh.shape_value(q, j) * (v @ n) * dS
                for i in range(1, nt + 1):
                    for j in range(1, nt + 1):
                        He[nt * (face - 1) + i - 1, nt * (face - 1) + j - 1] += Mh.shape_value(q, j) * Mh.shape_value(q, i) * dS
        assert np.allclose(Ae, 0.5 * np.eye(6), atol=1e-14)                   # :125
        assert np.allclose(Be, Be_ex[c], atol=1e-13)                          # :126
        assert np.allclose(Ce, Ce_ex[c], atol=1e-13) and np.allclose(Ee, Ee_ex[c], atol=1e-13) and np.allclose(He, He_ex[c], atol=1e-13)   # :176-178
    assert np.allclose([Mh.getfacedetJdS(f, 1) for f in (1, 2, 3)], [sq2 / 4, sq2 / 4, 0.5])                # :39-41 (state after the loop)
    with pytest.raises(hdg.BadGeometryError):
        Wh.reinit_(mesh.nodes[mesh.cells[0, [0, 2, 1]] - 1])                  # clockwise cell: det(J) is not positive


# ---- the C oracle's full-size helpers (used by tests/test_full_size_parity.py and bench.py) against the numpy oracle ----
@pytest.mark.parametrize("nx,ny,LL,UR", [(1, 1, (0, 0), (1, 1)), (7, 5, (0.0, 0.0), (2.0, 1.0)), (10, 10, (0, 0), (1, 1)),
                                         (3, 17, (-1.0, 0.5), (2.0, 3.0)), (40, 1, (0.0, 0.0), (1.0, 1.0))])
def test_c_rectangle_mesh_is_the_numpy_oracles_bit_for_bit(nx, ny, LL, UR):
    a, b = orc.rectangle_mesh(nx, ny, LL, UR), occ.rectangle_mesh(nx, ny, LL, UR)
    assert np.array_equal(a.cells, b.cells) and np.array_equal(a.cell_faces, b.cell_faces)
    assert np.array_equal(a.faces, b.faces) and np.array_equal(a.nodes, b.nodes)       # coordinates bit for bit
    assert a.facesets["boundary"] == b.facesets["boundary"]


@pytest.mark.parametrize("order,qd", [(1, 2), (3, 6)])
def test_c_recover_errornorm_apply_match_numpy_oracle(order, qd):
    mesh = orc.rectangle_mesh(6, 5, (0.0, 0.0), (2.0, 1.0))
    r = orc.run_poisson(mesh, order, qd)
    tab, asm = r["tab"], r["asm"]
    sig, u = occ.recover(mesh, tab, r["uhat"], asm.K_e, asm.b_e, nthreads=2)
    assert np.abs(sig - r["sigma"]).max() < 1e-13 * np.abs(r["sigma"]).max()
    assert np.abs(u - r["u"]).max() < 1e-13 * np.abs(r["u"]).max()
    assert abs(occ.errornorm(mesh, tab, r["u"], nthreads=2) - r["err2"]) < 1e-12 * r["err2"]
    K2, f2, m2, dset = occ.apply_dirichlet_homogeneous(asm.K, asm.rhs, r["dofs"])
    assert m2 == pytest.approx(r["meandiag"], rel=1e-15)
    assert np.array_equal(K2.indptr, r["K_bc"].indptr) and np.array_equal(K2.indices, r["K_bc"].indices)
    assert np.array_equal(K2.data, r["K_bc"].data) and np.array_equal(f2, r["rhs_bc"])
    assert dset.sum() == len(r["dofs"])


def test_renumbered_mesh_maps_back():
    """renumber_mesh (host mirror of the partitioner's pre-processing): with a given permutation and the numpy first-encounter
    numbering, faces keep their vertex pairs, face sets follow, and the maps bring trace / cell data back to the caller's ids."""
    rng = np.random.default_rng(3)
    mo = orc.rectangle_mesh(7, 5)
    bnd = set((np.flatnonzero(mo.faces[:, 3] == 0) + 1).tolist())
    mesh = hdg.PolygonalMesh(np.hstack([mo.cells, mo.cell_faces]), mo.nodes, mo.faces, {"boundary": bnd, "some": {3, 17, 40}})

    def number(tri, nodes):
        cf, faces = hdg.number_faces(tri)
        return np.hstack([tri, cf]), faces
    perm = rng.permutation(mesh.cells.shape[0])
    rm = hdg.renumber_mesh(mesh, perm, number=number)
    new = rm.mesh
    assert np.array_equal(new.cells[:, :3], mesh.cells[perm, :3])
    old_pairs = np.sort(np.asarray(mesh.faces)[:, :2], axis=1)
    new_pairs = np.sort(np.asarray(new.faces)[:, :2], axis=1)
    assert np.array_equal(new_pairs[rm.face_new], old_pairs)                      # a face keeps its two vertices
    assert new.facesets["boundary"] == set((np.flatnonzero(np.asarray(new.faces)[:, 3] == 0) + 1).tolist())
    assert new.facesets["some"] == {int(rm.face_new[f - 1]) + 1 for f in (3, 17, 40)}
    nt = 3
    x_new = rng.random(new.faces.shape[0] * nt)
    x_old = rm.trace_to_original(x_new, nt).reshape(-1, nt)
    for f in (0, 5, old_pairs.shape[0] - 1):
        assert np.array_equal(x_old[f], x_new.reshape(-1, nt)[rm.face_new[f]])
    u_new = rng.random((new.cells.shape[0], 6))
    u_old = rm.cells_to_original(u_new)
    assert np.array_equal(u_old[perm], u_new)
