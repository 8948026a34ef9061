"""CPU-side checks of the C-ABI library: it loads, exports every symbol the header declares, the
host-side table builder matches the oracle tables, and compute entry points fail loudly without
a GPU (no CPU fallback).  No compute call is made here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import hdg_b200 as hdg
import hdg_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "hdg_b200.h")).read()
    declared = set(re.findall(r"\b(hdg_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = C.CDLL(hdg.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/hdg_b200.h but not exported"
    assert declared == set(hdg.SIGNATURES), declared ^ set(hdg.SIGNATURES)


def test_struct_layouts_match_header():
    assert C.sizeof(hdg.api.Params) == 32
    assert C.sizeof(hdg.api.Sizes) == 4 * 8 + 6 * 4 + 2 * 8
    assert C.sizeof(hdg.api.SolveInfo) == 32


def test_sm100a_cubin_present():
    out = os.popen(f"cuobjdump -lelf {hdg.LIB_PATH} 2>/dev/null").read()
    if not out:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out


@pytest.mark.parametrize("order,qd", [(1, 2), (1, 3), (2, 3), (2, 4), (3, 6), (3, 7), (4, 9)])
def test_reference_tables_match_oracle(order, qd):
    tab = orc.build_tables(order, qd)
    g = lambda nm: hdg.ref_table(order, qd, nm)
    assert np.array_equal(g("qpoints").reshape(-1, 2), tab.qpts) or np.allclose(g("qpoints").reshape(-1, 2), tab.qpts, atol=1e-15)
    assert np.allclose(g("qweights"), tab.qw, atol=3e-15)
    assert np.allclose(g("fpoints"), tab.fs, atol=3e-16)
    assert np.allclose(g("fweights"), tab.fw, atol=3e-16)
    assert np.allclose(g("N").reshape(tab.nq, tab.n).T, tab.N, atol=2e-14)
    assert np.allclose(g("dNdxi").reshape(tab.nq, tab.n, 2).transpose(1, 0, 2), tab.dN, rtol=1e-13, atol=1e-13)
    assert np.allclose(g("E").reshape(3, tab.nfq, tab.n).transpose(2, 1, 0), tab.E, atol=5e-14)
    assert np.allclose(g("T").reshape(tab.nfq, tab.nt).T, tab.T, atol=1e-14)


def test_default_quad_degree_is_order_plus_one():
    assert hdg.ref_table(2, 0, "qweights").size == 6      # Strang(3), src/ScalarFunctionSpaces.jl:24-25
    assert hdg.ref_table(1, 0, "fweights").size == 2


def test_unsupported_rule_and_order_errors():
    with pytest.raises(hdg.UnsupportedRuleError):          # src/quadrature.jl:24
        hdg.ref_table(2, 8, "qweights")
    with pytest.raises(hdg.HDGError):
        hdg.ref_table(5, 0, "qweights")
    fe = hdg.GenericFiniteElement(hdg.Dubiner(2, hdg.RefTetrahedron, 2))
    mesh = hdg.PolygonalMesh(np.array([[1, 2, 3, 1, 2, 3]]), np.array([[0., 0.], [1., 0.], [0., 1.]]),
                             np.array([[2, 3, 1, 0], [3, 1, 1, 0], [1, 2, 1, 0]]), {"boundary": {1, 2, 3}})
    with pytest.raises(hdg.UnsupportedRuleError):
        hdg.ScalarFunctionSpace(mesh, fe, quad_degree=8).getnquadpoints()


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    lib = hdg.load()
    prm = hdg.api.Params(1, 2, 1.0, 1, -1, 0, 0)
    h = C.c_void_p()
    st = lib.hdg_create(C.byref(prm), C.byref(h))
    assert st == 5 and not h.value                          # HDG_ERR_CUDA
    assert b"no CPU fallback" in lib.hdg_last_error(None)
    with pytest.raises(hdg.HDGError):
        hdg.rectangle_mesh(hdg.TriangleCell, (2, 2), (0, 0), (1, 1))


def test_null_context_is_rejected():
    lib = hdg.load()
    assert lib.hdg_assemble(None) == 1
    assert lib.hdg_solve(None, 1e-8, 10, None) == 1
    assert lib.hdg_launch_count(None) == 0
    assert lib.hdg_set_dirichlet_faces(None, None, 0) == 1
    assert lib.hdg_errornorm_values(None, None, None) == 1
    assert lib.hdg_measure_fp64_peak(None, None) == 1


def test_library_basis_values_match_the_reference_closed_forms():
    """test/test_basis.jl:6-12 and :62-93 against the LIBRARY's own basis evaluation (hdg_basis_value, the functions its
    table builder uses): Dubiner values equal the reference's closed forms dubiner_basis at the Strang(5) points, gradients
    match their central differences, the Legendre functions equal the closed forms at the 3-point Gauss rule."""
    import math
    from test_oracle_goldens import _dubiner_closed
    dub, leg = hdg.Dubiner(2, hdg.RefTetrahedron, 4), hdg.Legendre(1, hdg.RefTetrahedron, 3)
    pts = hdg.ref_table(4, 5, "qpoints").reshape(-1, 2)          # Strang(5), src/StrangQuad.jl
    assert pts.shape == (7, 2)
    h = 1e-6
    for j in [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 13, 14, 15]:
        for r, s in pts:
            assert abs(hdg.value(dub, j, (r, s)) - _dubiner_closed(j, r, s)) < 2e-13
            g = hdg.gradient_value(dub, j, (r, s))
            gx = (_dubiner_closed(j, r + h, s) - _dubiner_closed(j, r - h, s)) / (2 * h)
            gy = (_dubiner_closed(j, r, s + h) - _dubiner_closed(j, r, s - h)) / (2 * h)
            assert abs(g[0] - gx) < 2e-6 * max(1, abs(gx)) and abs(g[1] - gy) < 2e-6 * max(1, abs(gy))
    ref = [lambda x: 1.0, lambda x: math.sqrt(3) * (2 * x - 1), lambda x: math.sqrt(5) * (6 * x * x - 6 * x + 1),
           lambda x: math.sqrt(7) * (2 * x - 1) * (10 * x * x - 10 * x + 1)]
    dref = [lambda x: 0.0, lambda x: math.sqrt(3) * 2, lambda x: math.sqrt(5) * (12 * x - 6), lambda x: math.sqrt(7) * 12 * (5 * x * x - 5 * x + 1)]
    for x in hdg.ref_table(2, 3, "fpoints"):                     # 3-point Gauss-Legendre on (0,1)
        for i in range(4):
            assert abs(hdg.value(leg, i + 1, x) - ref[i](x)) < 1e-14
            assert abs(hdg.gradient_value(leg, i + 1, x)[0] - dref[i](x)) < 1e-7
    with pytest.raises(hdg.HDGError):
        hdg.value(dub, 16, (0.2, 0.2))


def test_library_quadrature_rules():
    """test/test_quadrature.jl:16-25 on the library's rules: cell weights sum to the area of the reference triangle, the
    1-D Gauss rule to 1; point counts of the default rules (src/quadrature.jl:17-39)."""
    counts = {2: 3, 3: 6, 4: 6, 5: 7, 6: 12, 9: 35}
    for qd, npts in counts.items():
        w = hdg.ref_table(1, qd, "qweights")
        assert w.size == npts and abs(w.sum() - 0.5) < 1e-14
        fw, fp = hdg.ref_table(1, qd, "fweights"), hdg.ref_table(1, qd, "fpoints")
        assert fw.size == qd and abs(fw.sum() - 1.0) < 1e-14 and np.all(np.diff(fp) > 0) and 0.0 < fp[0] and fp[-1] < 1.0
        p = hdg.ref_table(1, qd, "qpoints").reshape(-1, 2)
        assert np.all(p > 0) and np.all(p.sum(axis=1) < 1)               # interior points
