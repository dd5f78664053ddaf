"""ExplicitIntegration/Term call surface: ``termLaxFriedrichs``."""
import copy

from .engine import engine_for_grid, weno_mode_of
from .functors import resolve
from .utilities import iscell, isfield

__all__ = ["termLaxFriedrichs", "termRestrictUpdate", "prepare_scheme", "unwrap_scheme"]

_COSTATE = ("upwindFirstWENO5", "upwindFirstWENO5a", "upwindFirstENO3a", "upwindFirstENO3", "upwindFirstENO2")


def prepare_scheme(schemeData):
    """Validate schemeData the way term_lax_friedrich.py:85-89 does, check that every callable in it is one this
    library has a device implementation for, and return (engine, adapter)."""
    sd = schemeData[0] if iscell(schemeData) else schemeData
    for f in ("grid", "CoStateCalc", "dissFunc", "hamFunc", "partialFunc"):
        assert isfield(sd, f), "%s not in bundle thisschemeData" % f      # same messages as the reference
    name = getattr(sd.CoStateCalc, "__name__", None)
    if name not in _COSTATE:
        raise NotImplementedError("CoStateCalc=%r: only upwindFirstWENO5(a) / upwindFirstENO3(a) / upwindFirstENO2 run on the device" % (sd.CoStateCalc,))
    diss_name = getattr(sd.dissFunc, "__name__", None)
    if diss_name not in ("artificialDissipationGLF", "artificialDissipationLLF"):
        raise NotImplementedError("dissFunc=%r: only artificialDissipationGLF / artificialDissipationLLF are fused into "
                                  "the stage kernel" % (sd.dissFunc,))
    adapter = sd.__dict__.get("_hjb200_adapter")
    if (adapter is None or adapter[0] is not sd.hamFunc or adapter[1] is not sd.partialFunc
            or (adapter[2].dynamic and (adapter[2].dyn is not sd.__dict__.get("dynSys") or adapter[2].sd is not sd))):
        ad = resolve(sd.hamFunc, sd.partialFunc, sd.grid, sd)
        try:
            sd._hjb200_adapter = (sd.hamFunc, sd.partialFunc, ad)
        except Exception:
            pass
    else:
        ad = adapter[2]
    if diss_name == "artificialDissipationLLF" and not ad.host_alpha:
        # as shipped LLF only runs for systems whose alphas are all scalars (Bird, Flock), where it equals GLF; with an
        # array alpha its `(1 / stepBoundInv).get().item()` raises (diss_local_laxfried.py:126-134)
        raise ValueError("can only convert an array of size 1 to a Python scalar")
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
    tables = list(enumerate(ad.tables(eng_grid(schemeData))))
    if ad.dynamic:
        # genericPartial: alpha needs the derivative range of this very field (generic_partial.py:28-40) -- one
        # reduce-only pass over it first, the dynSys's get_opt_u / get_opt_v on the range, then the fused RHS
        lo, hi = eng.deriv_range(y)
        block = ad.block_for_range(lo, hi, t)
    else:
        block = ad.block()
    eng.set_system(ad.system_id, block, tables)
    ydot, step_bound, red = eng.rhs(t, y)
    if iscell(schemeData):
        schemeData[0] = copy.copy(schemeData[0])
    return ydot, step_bound, schemeData


def unwrap_scheme(schemeFunc, schemeData):
    """(Lax-Friedrichs schemeData, restrict sign) for the two term functions the device path runs:
    ``termLaxFriedrichs`` (sign 0) and ``termRestrictUpdate`` wrapped around it (term_restrict_update.py:68-94:
    .innerFunc / .innerData / optional .positive, default True -> +1, False -> -1)."""
    name = getattr(schemeFunc, "__name__", None)
    if name == "termLaxFriedrichs":
        return schemeData, 0
    if name == "termRestrictUpdate":
        sd = schemeData[0] if iscell(schemeData) else schemeData
        assert isfield(sd, "innerFunc"), "innerFunc not in schemeData"     # term_restrict_update.py:65-66
        assert isfield(sd, "innerData"), "innerData not in schemeData"
        if getattr(sd.innerFunc, "__name__", None) != "termLaxFriedrichs":
            raise NotImplementedError("termRestrictUpdate.innerFunc=%r: only termLaxFriedrichs is compiled for the device"
                                      % (sd.innerFunc,))
        positive = sd.positive if isfield(sd, "positive") else True         # :85-88
        return sd.innerData, (1 if positive else -1)
    raise NotImplementedError("schemeFunc=%r: only termLaxFriedrichs / termRestrictUpdate(termLaxFriedrichs) are "
                              "compiled for the device" % (schemeFunc,))


def termRestrictUpdate(t, y, schemeData):
    """[ydot, stepBound, schemeData] = termRestrictUpdate(t, y, schemeData)
    -- ExplicitIntegration/Term/term_restrict_update.py:9-96: the inner term's update with its sign restricted,
    ``max(ydot, 0)`` (schemeData.positive, default) or ``min(ydot, 0)``; the restriction is fused into the same
    kernel launch as the inner termLaxFriedrichs.  ``ydot`` is returned squeezed to (n,) like the reference (:92,:94)."""
    inner, sign = unwrap_scheme(termRestrictUpdate, schemeData)
    eng, ad = prepare_scheme(inner)
    if iscell(y):
        y = y[0]
    if ad.dynamic:
        lo, hi = eng.deriv_range(y)
        block = ad.block_for_range(lo, hi, t)
    else:
        block = ad.block()
    eng.set_system(ad.system_id, block, list(enumerate(ad.tables(eng_grid(inner)))))
    eng.set_restrict(sign)
    try:
        ydot, step_bound, _ = eng.rhs(t, y)
    finally:
        eng.set_restrict(0)
    return ydot.reshape(-1), step_bound, schemeData


def eng_grid(schemeData):
    sd = schemeData[0] if iscell(schemeData) else schemeData
    return sd.grid
