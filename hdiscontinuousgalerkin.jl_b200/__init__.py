"""hdiscontinuousgalerkin.jl_b200 - B200-native HDG Poisson hot path behind the reference's API.

The directory name carries a dot, so it cannot be imported with a plain `import`; use the
loader module at the repository root:

    import hdg_b200 as hdg      # loads this package under the name `hdg_b200`
"""
from ._lib import (HDGError, BadGeometryError, UnsupportedRuleError, SingularLocalError, NotBoundaryError,
                   NotConvergedError, LIB_PATH, SIGNATURES, load)
from .api import *  # noqa: F401,F403
from .api import _Context  # noqa: F401
from . import cg  # noqa: F401  (CG side of the exported API: DofHandler, create_sparsity_pattern, ...)
from .cg import (ContinuousLagrange, LagrangeField, DofHandler, DirichletCG, create_sparsity_pattern, ndofs,  # noqa: F401
                 ndofs_per_cell, dof_range, reconstruct_, get_vertices_matrix, getcells_matrix, CGDevice, poisson2D_CG)
