"""examples/poisson2D_HDG.jl of the reference, line for line, on the B200 library."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hdg_b200 as hdg

mesh = hdg.rectangle_mesh(hdg.TriangleCell, (10, 10), (0.0, 0.0), (1.0, 1.0))
dim = 2
finiteElement = hdg.GenericFiniteElement(hdg.Dubiner(dim, hdg.RefTetrahedron, 1))
Wh = hdg.ScalarFunctionSpace(mesh, finiteElement)
Vh = hdg.VectorFunctionSpace(mesh, finiteElement)
Mh = hdg.ScalarTraceFunctionSpace(Wh, hdg.GenericFiniteElement(hdg.Legendre(dim - 1, hdg.RefTetrahedron, 1)))
uhat_h, sigma_h, u_h = hdg.TrialFunction(Mh), hdg.TrialFunction(Vh), hdg.TrialFunction(Wh)
dbc = hdg.Dirichlet(uhat_h, mesh, "boundary", lambda x: 0)
K, b, K_e, b_e = hdg.doassemble(Vh, Wh, Mh, 1.0, hdg.poisson_source)
hdg.apply_(K, b, dbc)
uhat, info = hdg.solve(K, b)
hdg.get_usigma_(sigma_h, u_h, uhat_h, uhat, K_e, b_e, mesh)
Etu_h = hdg.errornorm(u_h, hdg.poisson_exact)
print(f"PCG iterations {info['iterations']}, squared L2 error {Etu_h:.6e} (reference bound 6e-5)")
assert Etu_h <= 0.00006
