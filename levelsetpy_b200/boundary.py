"""Ghost-cell operators of the call surface: ``grid.bdry[d]`` holds one of these two callables.

Inside the fused stage kernels ghost cells are generated on the fly from tile indices (no padded copy is ever
made); the callables themselves are the reference's standalone operators
(BoundaryCondition/add_ghost_extrapolate.py:16, add_ghost_periodic.py:12) evaluated on the GPU (hj_add_ghost),
and they are the *tokens* by which the engine recognises a dim's boundary condition.
"""
import numpy as np


__all__ = ["addGhostExtrapolate", "addGhostPeriodic"]

_ENGINES = {}


class _PseudoGrid:
    pass


def _ghost(data, dim, width, kind_fn, toward_zero):
    from .engine import Engine
    shape = tuple(int(s) for s in data.shape)
    if width is None or not width:
        width = 1                                            # add_ghost_extrapolate.py:55-56
    if width < 0 or width > shape[dim]:
        raise ValueError("Illegal width parameter")          # :58-59
    key = (shape, dim, kind_fn.__name__, bool(toward_zero))
    eng = _ENGINES.get(key)
    if eng is None:
        g = _PseudoGrid()
        g.dim = len(shape)
        g.N = np.array(shape).reshape(-1, 1)
        g.dx = np.ones((g.dim, 1))
        g.vs = [np.arange(n, dtype=np.float64) for n in shape]
        g.bdry = [kind_fn] * g.dim
        td = _PseudoGrid()
        td.towardZero = bool(toward_zero)
        g.bdryData = [td] * g.dim
        if len(_ENGINES) > 8:
            _ENGINES.pop(next(iter(_ENGINES))).close()
        eng = _ENGINES[key] = Engine(g)
    return eng.add_ghost(data, dim, int(width))


def addGhostExtrapolate(dataIn, dim, width=None, ghostData=None):
    """dataOut = addGhostExtrapolate(dataIn, dim, width, ghostData): ``width`` ghost cells on each side of axis
    ``dim``, linearly extrapolated with slope ``m*|edge-next|*sign(edge)`` (m = -1 if ghostData.towardZero)."""
    tz = bool(getattr(ghostData, "towardZero", False)) if ghostData is not None else False
    return _ghost(dataIn, dim, width, addGhostExtrapolate, tz)


def addGhostPeriodic(dataIn, dim, width=None, ghostData=None):
    """dataOut = addGhostPeriodic(dataIn, dim, width, ghostData): wrap ``width`` cells from the opposite side."""
    return _ghost(dataIn, dim, width, addGhostPeriodic, False)
