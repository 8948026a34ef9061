"""hdiscontinuousgalerkin.jl_b200 - B200-native HDG Poisson hot path behind the reference's API.

The directory name carries a dot, so it cannot be imported with a plain `import`; use the
loader module at the repository root:

    import hdg_b200 as hdg      # loads this package under the name `hdg_b200`
"""
from ._lib import (HDGError, BadGeometryError, UnsupportedRuleError, SingularLocalError, NotBoundaryError,
                   NotConvergedError, LIB_PATH, SIGNATURES, load)
from .api import *  # noqa: F401,F403
from .api import _Context  # noqa: F401
