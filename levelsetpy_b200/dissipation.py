"""ExplicitIntegration/Dissipation call surface: ``artificialDissipationGLF`` and ``artificialDissipationLLF``.

Inside ``termLaxFriedrichs`` / ``odeCFL3`` the dissipation is not a separate pass -- the stage kernel forms
0.5*(R-L)*alpha_d, the derivative min / max and max alpha_d in registers while the derivatives are still there -- and
``schemeData.dissFunc`` only has to NAME the scheme.  Called on their own (the reference's termLaxFriedrichs does,
term_lax_friedrich.py:123) these functions evaluate the same thing on dense arrays on the device (C-ABI hj_diss_glf).
"""
import numpy as np

from .utilities import isfield

__all__ = ["artificialDissipationGLF", "artificialDissipationLLF"]


def _amin(a):
    return a.min().item() if hasattr(a, "is_cuda") else np.min(a)


def _amax(a):
    return a.max().item() if hasattr(a, "is_cuda") else np.max(a)


def _adapter_of(schemeData, t=0.0, derivL=(), derivR=()):
    assert isfield(schemeData, "grid"), "grid not in schemeData"                    # artificial_diss_glf.py:64-65
    assert isfield(schemeData, "partialFunc"), "partialFunc not in schemeData"
    from .engine import engine_for_grid
    from .functors import _adapter_for_owner
    pf = schemeData.partialFunc
    owner = getattr(pf, "__self__", None)
    if getattr(pf, "__name__", None) == "genericPartial":
        # alpha needs the derivative range (generic_partial.py:28-40): the min / max of the arrays the caller handed in
        from .functors import generic_adapter
        ad = generic_adapter(schemeData)
        eng = engine_for_grid(schemeData.grid)
        lo = [float(min(_amin(a), _amin(b))) for a, b in zip(derivL, derivR)]       # artificial_diss_glf.py:82-88
        hi = [float(max(_amax(a), _amax(b))) for a, b in zip(derivL, derivR)]
        eng.set_system(ad.system_id, ad.block_for_range(lo, hi, t), list(enumerate(ad.tables(schemeData.grid))))
        return eng, ad
    if owner is None:
        raise NotImplementedError("partialFunc=%r is not a bound method of a registered DynamicalSystem "
                                  "(no CPU fallback)" % (pf,))
    ham_name = "hamiltonian_abs" if pf.__name__ == "dissipation_abs" else "hamiltonian"
    ad = _adapter_for_owner(owner, ham_name, pf.__name__)
    eng = engine_for_grid(schemeData.grid)
    block = ad.block(False) if ad.time_varying else ad.block()      # partialFunc alone does not mutate a Flock
    eng.set_system(ad.system_id, block, list(enumerate(ad.tables(schemeData.grid))))
    return eng, ad


def artificialDissipationGLF(t, data, derivL, derivR, schemeData):
    """diss, stepBound = artificialDissipationGLF(t, data, derivL, derivR, schemeData) -- global Lax-Friedrichs
    dissipation, ExplicitIntegration/Dissipation/artificial_diss_glf.py:7-111: diss = sum_d 0.5 (R_d - L_d) alpha_d
    (:100, dims in order), stepBound = 1 / sum_d max_x alpha_d / dx_d (:104-109).  ``derivL`` / ``derivR``: lists of
    grid.dim numpy arrays or torch CUDA tensors of grid.shape; ``diss`` comes back as the same kind."""
    eng, _ = _adapter_of(schemeData, t, derivL, derivR)
    diss, step_bound, _ = eng.diss_glf(t, list(derivL), list(derivR))
    return diss, step_bound


def artificialDissipationLLF(t, data, derivL, derivR, schemeData):
    """diss, stepBound = artificialDissipationLLF(...) -- local Lax-Friedrichs,
    ExplicitIntegration/Dissipation/diss_local_laxfried.py:14-136.  For dim i LLF hands partialFunc the LOCAL costate
    range [min(L_i, R_i), max(L_i, R_i)] (:117-120); every registered system's alpha ignores the costate range, so
    ``diss`` equals GLF's.  As shipped the step bound is ``(1 / stepBoundInv).get().item()`` with stepBoundInv =
    sum_i alpha_i / dx_i UN-maximised (:126-134): that is an array -- and ``.item()`` raises -- as soon as one alpha is an
    array (DubinsVehicleRel, DoubleIntegrator); with all-scalar alphas (Bird, Flock) it is GLF's bound.  Same here."""
    eng, ad = _adapter_of(schemeData, t, derivL, derivR)
    if not ad.host_alpha:
        raise ValueError("can only convert an array of size 1 to a Python scalar")   # diss_local_laxfried.py:134 as shipped
    diss, step_bound, _ = eng.diss_glf(t, list(derivL), list(derivR))
    return diss, step_bound
