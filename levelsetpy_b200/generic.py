"""Hamiltonians call surface: ``genericHam`` / ``genericPartial`` (Hamiltonians/generic_ham.py:5-57,
generic_partial.py:6-58) over ``schemeData.dynSys``, and the registered dynSys class ``DubinsCar``.

The reference evaluates both through three Python methods of the dynSys -- ``get_opt_u(t, deriv, uMode, x)``,
``get_opt_v(t, deriv, dMode, x)`` and ``dynamics(t, x, u, d)`` -- on full-grid arrays.  Here a dynSys is a compiled
device functor (csrc/hj_systems.cuh: ``GenericF<Dyn>``: optimal inputs, dynamics and the Hamiltonian per node in the
stage kernel); what stays on the host is exactly what generic_partial.py:28-40 does with grid-wide SCALARS: before every
RHS evaluation the derivative range of the field is reduced on the device (C-ABI ``hj_deriv_range``), the dynSys's own
``get_opt_u`` / ``get_opt_v`` are called on ``derivMax`` / ``derivMin`` and the four resulting input sets go into the
functor's parameter block, from which the kernel forms alpha = max |dynamics(x, uU|uL, dU|dL)[dim]| (:44-56).

Inside ``termLaxFriedrichs`` / ``odeCFL3`` / ``HJIPDE_solve`` the two functions are tokens naming that path; a dynSys
class without a compiled functor raises NotImplementedError (no CPU fallback).
"""
import numpy as np

__all__ = ["genericHam", "genericPartial", "DubinsCar"]


def genericHam(t, data, deriv, schemeData):
    """hamValue = genericHam(t, data, deriv, schemeData) -- Hamiltonians/generic_ham.py:5-57 for a registered dynSys:
    the functor's Hamiltonian on dense arrays (C-ABI hj_ham).  uMode / dMode / tMode default to 'min' / 'max' /
    'backward' and are written back into schemeData like the reference does (:11-18)."""
    from .functors import generic_adapter
    from .engine import engine_for_grid
    ad = generic_adapter(schemeData, set_defaults="ham")
    eng = engine_for_grid(schemeData.grid)
    eng.set_system(ad.system_id, ad.block_for_range(None, None, t), list(enumerate(ad.tables(schemeData.grid))))
    return eng.ham(t, list(deriv))


def genericPartial(t, data, derivMin, derivMax, schemeData, dim):
    """alpha = genericPartial(t, data, derivMin, derivMax, schemeData, dim) -- Hamiltonians/generic_partial.py:6-58:
    the dynSys's optimal inputs at derivMax / derivMin (host scalars, through its own get_opt_u / get_opt_v) and
    alpha = max over the four input combinations of |dynamics[dim]|, evaluated on the device (C-ABI hj_alpha).
    dMode defaults to 'min' HERE (:19-20) -- as shipped; inside termLaxFriedrichs genericHam has already set 'max'."""
    from .functors import generic_adapter
    from .engine import engine_for_grid
    ad = generic_adapter(schemeData, set_defaults="partial")
    eng = engine_for_grid(schemeData.grid)
    eng.set_system(ad.system_id, ad.block_for_range(derivMin, derivMax, t), list(enumerate(ad.tables(schemeData.grid))))
    return eng.alpha(t, dim, data)


class DubinsCar:
    """The 3-D Dubins car with disturbances that genericHam / genericPartial were written for (helperOC's DubinsCar):

        dx0 = speed cos x2 + d0,   dx1 = speed sin x2 + d1,   dx2 = u + d2,   |u| <= wMax,  |d_i| <= dMax[i]

    with the dynSys methods the reference's generic functions call.  They are plain numpy (the reference calls them on
    full-grid arrays; here the host only ever calls them on the scalar derivative range), and the class has a compiled
    device counterpart (csrc/hj_systems.cuh: DubinsCarDyn), which is what makes it usable as ``schemeData.dynSys``."""

    nx = 3

    def __init__(self, speed=1.0, wMax=1.0, dMax=(0.0, 0.0, 0.0)):
        self.speed = float(speed)
        self.wMax = float(wMax)
        self.dMax = [float(v) for v in dMax]
        if len(self.dMax) != 3:
            raise ValueError("dMax needs one bound per state")

    def get_opt_u(self, t, deriv, uMode="min", y=None):
        if uMode == "max":
            return (deriv[2] >= 0) * self.wMax + (deriv[2] < 0) * (-self.wMax)
        if uMode == "min":
            return (deriv[2] >= 0) * (-self.wMax) + (deriv[2] < 0) * self.wMax
        raise ValueError("Unknown uMode!")

    def get_opt_v(self, t, deriv, dMode="max", y=None):
        if dMode not in ("max", "min"):
            raise ValueError("Unknown dMode!")
        s = 1.0 if dMode == "max" else -1.0
        return [(deriv[i] >= 0) * (s * self.dMax[i]) + (deriv[i] < 0) * (-(s * self.dMax[i])) for i in range(3)]

    def dynamics(self, t, x, u, d):
        return [self.speed * np.cos(x[2]) + d[0], self.speed * np.sin(x[2]) + d[1], u + d[2]]
