"""examples/poisson2D_CG.jl of the reference on the B200 library: DofHandler numbering and sparsity pattern (bit-identical to
src/dofhandler.jl), doassemble + assemble!, Dirichlet + apply!, K \\ b, reconstruct! + errornorm - all on the device (hdg_cg_*)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hdg_b200 as hdg

mesh = hdg.rectangle_mesh(hdg.TriangleCell, (10, 10), (0.0, 0.0), (1.0, 1.0))
res = hdg.poisson2D_CG(mesh, order=1)
dh = res["dofhandler"]
print(f"{res['ndofs']} dofs, PCG iterations {res['info']['iterations']}, squared L2 error {res['err2']:.6e} (reference bound 2e-4, test/test_CGExample.jl:93)")
assert res["err2"] <= 0.0002
