"""levelsetpy_b200 -- B200-native (sm_100a) drop-in for LevelSetPy's explicit Hamilton-Jacobi time-stepping path:
upwindFirstWENO5/5a + addGhostExtrapolate/Periodic + termLaxFriedrichs(artificialDissipationGLF) + odeCFL3, as
driven by HJIPDE_solve.  Same names and call surface as the reference; the numerics run in hand-written CUDA
behind the C-ABI in include/hjb200.h.  No CPU fallback: importing works anywhere, computing needs the built
library and a GPU.
"""
from .utilities import *  # noqa: F401,F403
from .boundary import addGhostExtrapolate, addGhostPeriodic  # noqa: F401
from .grids import createGrid, processGrid, flockGrid  # noqa: F401
from .initial_conditions import *  # noqa: F401,F403
from .spatial import upwindFirstWENO5, upwindFirstWENO5a, upwindFirstENO3a, upwindFirstENO3, upwindFirstENO2  # noqa: F401
from .dissipation import artificialDissipationGLF, artificialDissipationLLF  # noqa: F401
from .term import termLaxFriedrichs, termRestrictUpdate  # noqa: F401
from .integration import odeCFL3, odeCFL2, odeCFLset  # noqa: F401
from .systems import DubinsVehicleRel, DoubleIntegrator, Bird, Flock, ProductSystem  # noqa: F401
from .generic import genericHam, genericPartial, DubinsCar  # noqa: F401
from .solver import HJIPDE_solve  # noqa: F401
from .engine import Engine, engine_for_grid, clear_engine_cache  # noqa: F401

__version__ = "0.1.0"
from .batch import BatchSolver, batch_step_plan  # noqa: F401
