"""SpatialDerivative call surface: ``schemeData.CoStateCalc`` tokens + standalone GPU operators."""
from .engine import engine_for_grid

__all__ = ["upwindFirstWENO5a", "upwindFirstWENO5", "upwindFirstENO3a", "upwindFirstENO3", "upwindFirstENO2"]


def upwindFirstWENO5a(grid, data, dim, generateAll=False, wenoMode="as_shipped"):
    """[derivL, derivR] = upwindFirstWENO5a(grid, data, dim) -- SpatialDerivative/upwind_first_weno5a.py:13.

    ``data``: numpy array or torch CUDA tensor of shape grid.shape; result has the same kind and shape.
    ``wenoMode``: 'as_shipped' (the reference's behaviour) or 'intended' (true WENO5 weights)."""
    if dim < 0 or dim > grid.dim:
        raise ValueError("Illegal dim parameter")           # upwind_first_weno5a.py:59-60
    if generateAll:                                         # :73-75: the three ENO approximations of the helper
        return engine_for_grid(grid, wenoMode).deriv_candidates(data, dim)
    return engine_for_grid(grid, wenoMode).deriv(data, dim)


def upwindFirstWENO5(grid, data, dim, generateAll=False, wenoMode="as_shipped"):
    """Alias of upwindFirstWENO5a -- SpatialDerivative/upwind_first_weno5.py:11-48."""
    return upwindFirstWENO5a(grid, data, dim, generateAll, wenoMode)


def _eno(grid, data, dim, generateAll, scheme):
    if dim < 0 or dim > grid.dim:
        raise ValueError("Illegal dim parameter")           # upwind_first_eno2.py:52-53, upwind_first_eno3a.py:77-78
    if generateAll:
        if scheme != "eno3a":
            raise NotImplementedError("generateAll=True of upwindFirstENO2 (its two second-order candidates) is outside the hot path")
        return engine_for_grid(grid, scheme).deriv_candidates(data, dim)    # upwind_first_eno3a.py:82-84
    return engine_for_grid(grid, scheme).deriv(data, dim)


def upwindFirstENO3a(grid, data, dim, generateAll=False):
    """[derivL, derivR] = upwindFirstENO3a(grid, data, dim) -- SpatialDerivative/upwind_first_eno3a.py:13: third-order
    ENO, the candidate on the minimum-modulus D2 / D3 neighbours."""
    return _eno(grid, data, dim, generateAll, "eno3a")


def upwindFirstENO3(grid, data, dim, generateAll=False):
    """Alias of upwindFirstENO3a -- SpatialDerivative/upwind_first_eno3.py."""
    return _eno(grid, data, dim, generateAll, "eno3a")


def upwindFirstENO2(grid, data, dim, generateAll=False):
    """[derivL, derivR] = upwindFirstENO2(grid, data, dim) -- SpatialDerivative/upwind_first_eno2.py:13: second-order
    ENO (two ghost cells), minimum-modulus second-order term."""
    return _eno(grid, data, dim, generateAll, "eno2")
