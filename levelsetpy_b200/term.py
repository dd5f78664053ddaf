"""ExplicitIntegration/Term call surface: ``termLaxFriedrichs``."""
import copy

from . import _lib as L
from .engine import engine_for_grid, weno_mode_of
from .functors import resolve
from .utilities import iscell, isfield

__all__ = ["termLaxFriedrichs", "prepare_scheme"]

_COSTATE = ("upwindFirstWENO5", "upwindFirstWENO5a")


def prepare_scheme(schemeData):
    """Validate schemeData the way term_lax_friedrich.py:85-89 does, check that every callable in it is one this
    library has a device implementation for, and return (engine, adapter)."""
    sd = schemeData[0] if iscell(schemeData) else schemeData
    for f in ("grid", "CoStateCalc", "dissFunc", "hamFunc", "partialFunc"):
        assert isfield(sd, f), "%s not in bundle thisschemeData" % f      # same messages as the reference
    name = getattr(sd.CoStateCalc, "__name__", None)
    if name not in _COSTATE:
        raise NotImplementedError("CoStateCalc=%r: only upwindFirstWENO5 / upwindFirstWENO5a run on the device" % (sd.CoStateCalc,))
    if getattr(sd.dissFunc, "__name__", None) != "artificialDissipationGLF":
        raise NotImplementedError("dissFunc=%r: only artificialDissipationGLF is fused into the stage kernel" % (sd.dissFunc,))
    adapter = sd.__dict__.get("_hjb200_adapter")
    if adapter is None or adapter[0] is not sd.hamFunc or adapter[1] is not sd.partialFunc:
        ad = resolve(sd.hamFunc, sd.partialFunc, sd.grid)
        try:
            sd._hjb200_adapter = (sd.hamFunc, sd.partialFunc, ad)
        except Exception:
            pass
    else:
        ad = adapter[2]
    eng = engine_for_grid(sd.grid, weno_mode_of(sd))
    return eng, ad


def termLaxFriedrichs(t, y, schemeData):
    """[ydot, stepBound, schemeData] = termLaxFriedrichs(t, y, schemeData)
    -- ExplicitIntegration/Term/term_lax_friedrich.py:8-130, one fused kernel launch (hj_rhs).

    ``y``: numpy (n,1)/(n,) array or torch CUDA tensor; ``ydot`` comes back as the same kind, shape (n,1).
    schemeData fields: grid, CoStateCalc (upwindFirstWENO5/5a), dissFunc (artificialDissipationGLF),
    hamFunc / partialFunc (bound methods of a registered DynamicalSystem)."""
    eng, ad = prepare_scheme(schemeData)
    if iscell(y):
        y = y[0]
    block = ad.block()
    eng.set_system(ad.system_id, block, list(enumerate(ad.tables(eng_grid(schemeData)))))
    ydot, step_bound, red = eng.rhs(t, y)
    if iscell(schemeData):
        schemeData[0] = copy.copy(schemeData[0])
    return ydot, step_bound, schemeData


def eng_grid(schemeData):
    sd = schemeData[0] if iscell(schemeData) else schemeData
    return sd.grid
