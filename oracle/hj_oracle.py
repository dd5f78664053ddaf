"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

A numpy fp64 restatement of the reference's explicit Hamilton-Jacobi time-stepping
hot path (robotsorcerer/LevelSetPy).  Nothing under ``levelsetpy_b200/`` imports
this module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and only as the checker / CPU timing arm.

PARITY PIN.  The reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so the pin is made here: ``tests/golden/make_golden.py`` runs
the *literal* reference files (through ``oracle/ref_shim.py``) in the build container
and stores their outputs under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
checks this oracle against them bit-for-bit (``weno='as_shipped'``).

Two selectable semantics for the WENO5 weights:

``as_shipped``  bug-compatible with upwind_first_weno5a.py as shipped: the list
                aliasing at :97 makes every shifted view of D1 the same view, so the
                smoothness indicators are (numerically) zero and the scheme is the
                fixed-weight fifth-order upwind combination.
``intended``    the scheme the reference's own comments and ENO3bHelper.py:136-160
                write down (Osher & Fedkiw (3.32)-(3.41), Mitchell's toolbox): real
                smoothness indicators taken from the unstripped D1 table.

All arrays are C-order float64.  ``grid`` objects are duck-typed (``dim, N, dx, vs,
xs, bdry, bdryData, shape``) so the reference's own grids work as well.
"""
from __future__ import annotations

import sys

import numpy as np

EPS = sys.float_info.epsilon
REALMAX = sys.float_info.max

__all__ = [
    "add_ghost_extrapolate", "add_ghost_periodic", "add_ghost", "eno3a_helper", "upwind_first_weno5a",
    "upwind_first_eno3a", "upwind_first_eno2", "upwind_first",
    "artificial_dissipation_glf", "term_lax_friedrichs", "ode_cfl3_step", "ode_cfl3", "hji_solve",
    "bc_kind_of", "OracleSchemeData",
]


def _ax(a, dim, sl):
    """a[..., sl, ...] with the slice applied on axis ``dim``."""
    idx = [slice(None)] * a.ndim
    idx[dim] = sl
    return a[tuple(idx)]


# --------------------------------------------------------------------------------------
# BoundaryCondition
# --------------------------------------------------------------------------------------
def add_ghost_extrapolate(data, dim, width=3, toward_zero=False):
    """Follows BoundaryCondition/add_ghost_extrapolate.py:55-113.

    slope = m * |edge - next| * sign(edge) (:88-100), m = -1 if towardZero else +1 (:61-64);
    the ghost k cells outside the edge is ``edge + k*slope`` (:103-110).
    """
    m = -1 if toward_zero else +1
    n = data.shape[dim]
    bot0, bot1 = _ax(data, dim, slice(0, 1)), _ax(data, dim, slice(1, 2))
    top0, top1 = _ax(data, dim, slice(n - 1, n)), _ax(data, dim, slice(n - 2, n - 1))
    slope_bot = m * np.abs(bot0 - bot1) * np.sign(bot0)
    slope_top = m * np.abs(top0 - top1) * np.sign(top0)
    lo = [bot0 + (width - i) * slope_bot for i in range(width)]              # out[i], i = 0..w-1
    hi = [top0 + (width - i) * slope_top for i in range(width - 1, -1, -1)]  # out[end-i]
    return np.concatenate(lo + [data] + hi, axis=dim)


def add_ghost_periodic(data, dim, width=3):
    """Follows BoundaryCondition/add_ghost_periodic.py:50-89: wrap ``width`` cells from the far side."""
    n = data.shape[dim]
    return np.concatenate([_ax(data, dim, slice(n - width, n)), data, _ax(data, dim, slice(0, width))], axis=dim)


def bc_kind_of(grid, dim):
    """('periodic'|'extrapolate', toward_zero) for ``grid.bdry[dim]`` -- recognised by function name so the
    reference's own ``addGhostPeriodic`` / ``addGhostExtrapolate`` objects work (create_grid.py:61-65)."""
    fn = grid.bdry[dim]
    name = getattr(fn, "__name__", str(fn))
    if name == "addGhostPeriodic":
        return "periodic", False
    if name == "addGhostExtrapolate":
        gd = grid.bdryData[dim] if getattr(grid, "bdryData", None) is not None else None
        tz = bool(getattr(gd, "towardZero", False)) if gd is not None else False
        return "extrapolate", tz
    raise ValueError("oracle: unsupported boundary condition %r" % (name,))


def add_ghost(grid, data, dim, width=3):
    kind, tz = bc_kind_of(grid, dim)
    if kind == "periodic":
        return add_ghost_periodic(data, dim, width)
    return add_ghost_extrapolate(data, dim, width, tz)


# --------------------------------------------------------------------------------------
# SpatialDerivative
# --------------------------------------------------------------------------------------
def eno3a_helper(grid, data, dim):
    """Divided-difference ENO3 candidates.  Follows SpatialDerivative/ENO3aHelper.py:54-191.

    Returns (dL[3], dR[3], D1_stripped, D1_unstripped).  Operation order is the reference's:
    scalar products first (:78,:83,:88), candidates built by successive in-place ``+=`` (:132-189).
    """
    dx = float(np.asarray(grid.dx).reshape(-1)[dim])
    dx_inv = 1 / dx
    n = data.shape[dim]
    g = add_ghost(grid, data, dim, 3)                                      # :64  node i <-> g[i+3]
    d1u = dx_inv * (_ax(g, dim, slice(1, None)) - _ax(g, dim, slice(0, -1)))               # :78  N+5
    d2u = 0.5 * dx_inv * (_ax(d1u, dim, slice(1, None)) - _ax(d1u, dim, slice(0, -1)))     # :83  N+4
    d3 = (1 / 3) * dx_inv * (_ax(d2u, dim, slice(1, None)) - _ax(d2u, dim, slice(0, -1)))  # :88  N+3
    d1 = _ax(d1u, dim, slice(2, d1u.shape[dim] - 2))                        # :99-100  N+1
    d2 = _ax(d2u, dim, slice(1, d2u.shape[dim] - 1))                        # :105-106 N+2

    dL = [_ax(d1, dim, slice(0, n)).copy() for _ in range(3)]               # :118-119
    dR = [_ax(d1, dim, slice(1, n + 1)).copy() for _ in range(3)]           # :121-122
    cL, cR = +1 * dx, -1 * dx                                               # :132-133
    dL[0] += cL * _ax(d2, dim, slice(0, n))                                 # :139-141
    dL[1] += cL * _ax(d2, dim, slice(0, n))
    dL[2] += cL * _ax(d2, dim, slice(1, n + 1))
    dR[0] += cR * _ax(d2, dim, slice(1, n + 1))                             # :147-149
    dR[1] += cR * _ax(d2, dim, slice(1, n + 1))
    dR[2] += cR * _ax(d2, dim, slice(2, n + 2))
    cLL, cLR = +2 * dx ** 2, -1 * dx ** 2                                   # :165-168
    cRL, cRR = -1 * dx ** 2, +2 * dx ** 2
    dL[0] += cLL * _ax(d3, dim, slice(0, n))                                # :170-179
    dL[1] += cLL * _ax(d3, dim, slice(1, n + 1))
    dL[2] += cLR * _ax(d3, dim, slice(2, n + 2))
    dR[0] += cRL * _ax(d3, dim, slice(1, n + 1))                            # :181-189
    dR[1] += cRL * _ax(d3, dim, slice(2, n + 2))
    dR[2] += cRR * _ax(d3, dim, slice(3, n + 3))
    return dL, dR, d1, d1u


def _weight_weno(d, s, w, eps):
    """weightWENO, upwind_first_weno5a.py:177-196."""
    a1 = w[0] / (s[0] + eps) ** 2
    a2 = w[1] / (s[1] + eps) ** 2
    a3 = w[2] / (s[2] + eps) ** 2
    return (a1 * d[0] + a2 * d[1] + a3 * d[2]) / (a1 + a2 + a3)


def _smooth3(v0, v1, v2, v3, v4):
    """The three smoothness estimates as written at upwind_first_weno5a.py:107-123."""
    s0 = (13 / 12) * (v0 - 2 * v1 + v2) ** 2 + (1 / 4) * (v0 - 4 * v1 + 3 * v2) ** 2
    s1 = (13 / 12) * (v1 - 2 * v2 + v3) ** 2 + (1 / 4) * (v1 - v3) ** 2
    s2 = (13 / 12) * (v2 - 2 * v3 + v4) ** 2 + (1 / 4) * (3 * v2 - 4 * v3 + v4) ** 2
    return [s0, s1, s2]


def upwind_first_weno5a(grid, data, dim, weno="as_shipped"):
    """(derivL, derivR) along ``dim``.  Follows SpatialDerivative/upwind_first_weno5a.py:56-174.

    upwindFirstWENO5 (upwind_first_weno5.py:11-48) is an alias of this function.
    """
    n = data.shape[dim]
    dL, dR, d1, d1u = eno3a_helper(grid, data, dim)
    if weno == "as_shipped":
        # :97 aliases one index list five times; after the loop at :102-103 every entry holds
        # arange(size(D1,dim)-1), so every "shifted" view is D1[0:N].
        a = _ax(d1, dim, slice(0, n))
        smooth = _smooth3(a, a, a, a, a)
        sL = smooth                                                        # :130-131  indices 0..N-1
        sR = [np.roll(s, -1, axis=dim) for s in smooth]                    # :143-145  indices 1..N wrap on CuPy
        eps = 1e-6 * np.max(d1 ** 2) + 1e-99                               # :154-156  stripped table
    elif weno == "intended":
        # Mitchell's upwindFirstWENO5a / O&F (3.32)-(3.41): the unstripped table, five genuinely shifted views.
        v = [_ax(d1u, dim, slice(k, k + n + 1)) for k in range(5)]        # positions j = 0..N
        smooth = _smooth3(*v)
        sL = [_ax(s, dim, slice(0, n)) for s in smooth]
        sR = [_ax(s, dim, slice(1, n + 1)) for s in smooth]
        eps = 1e-6 * np.max(d1u ** 2) + 1e-99                              # 'maxOverGrid' over the unstripped table
    else:
        raise ValueError("weno must be 'as_shipped' or 'intended'")
    derivL = _weight_weno(dL, sL, [0.1, 0.6, 0.3], eps)                    # :134,:171
    derivR = _weight_weno(dR, sR, [0.3, 0.6, 0.1], eps)                    # :147,:172
    return derivL, derivR


def upwind_first_eno3a(grid, data, dim):
    """(derivL, derivR) of the third-order ENO scheme.  Follows SpatialDerivative/upwind_first_eno3a.py:87-142
    (upwindFirstENO3, upwind_first_eno3.py, is an alias).  The helper's DD tables are always the stripped ones
    (ENO3aHelper.py:109 overwrites :92): D2 has N+2 entries, D3 N+3.  The choice is applied the way the reference
    applies it, as a sum of candidate * boolean mask (:134-141)."""
    n = data.shape[dim]
    dx = float(np.asarray(grid.dx).reshape(-1)[dim])
    dx_inv = 1 / dx
    dL, dR, _, d1u = eno3a_helper(grid, data, dim)
    d2u = 0.5 * dx_inv * (_ax(d1u, dim, slice(1, None)) - _ax(d1u, dim, slice(0, -1)))
    d3 = (1 / 3) * dx_inv * (_ax(d2u, dim, slice(1, None)) - _ax(d2u, dim, slice(0, -1)))
    d2 = _ax(d2u, dim, slice(1, d2u.shape[dim] - 1))
    d2abs = np.abs(d2)                                                                     # :104
    smallerL = _ax(d2abs, dim, slice(0, n + 1)) < _ax(d2abs, dim, slice(1, n + 2))          # :108  N+1
    smallerR = np.logical_not(smallerL)
    d3abs = np.abs(d3)                                                                     # :114
    temp = _ax(d3abs, dim, slice(0, n + 2)) < _ax(d3abs, dim, slice(1, n + 3))              # :117  N+2
    lo, hi = slice(0, n + 1), slice(1, n + 2)
    smallerLL = np.logical_and(_ax(temp, dim, lo), smallerL)                               # :121-122
    smallerRL = np.logical_and(_ax(temp, dim, hi), smallerR)
    ntemp = np.logical_not(temp)
    smallerLR = np.logical_and(_ax(ntemp, dim, lo), smallerL)                              # :124-125
    smallerRR = np.logical_and(_ax(ntemp, dim, hi), smallerR)
    smallerM = np.logical_or(smallerRL, smallerLR)                                         # :127
    a, b = slice(0, n), slice(1, n + 1)
    derivL = dL[0] * _ax(smallerLL, dim, a) + dL[1] * _ax(smallerM, dim, a) + dL[2] * _ax(smallerRR, dim, a)   # :132-135
    derivR = dR[0] * _ax(smallerLL, dim, b) + dR[1] * _ax(smallerM, dim, b) + dR[2] * _ax(smallerRR, dim, b)   # :137-140
    return derivL, derivR


def upwind_first_eno2(grid, data, dim):
    """(derivL, derivR) of the second-order ENO scheme.  Follows SpatialDerivative/upwind_first_eno2.py:50-150:
    two ghost cells, D1 (N+3) / D2 (N+2) tables, minimum-modulus choice of the second-order term, applied as a sum
    of candidate * boolean mask (:145-148)."""
    n = data.shape[dim]
    dx = float(np.asarray(grid.dx).reshape(-1)[dim])
    dx_inv = 1 / dx
    g = add_ghost(grid, data, dim, 2)                                                       # :66
    d1 = dx_inv * (_ax(g, dim, slice(1, None)) - _ax(g, dim, slice(0, -1)))                 # :82  N+3
    d2 = 0.5 * dx_inv * (_ax(d1, dim, slice(1, None)) - _ax(d1, dim, slice(0, -1)))         # :86  N+2
    d1 = _ax(d1, dim, slice(1, d1.shape[dim] - 1))                                          # :92-93  N+1
    dL = [_ax(d1, dim, slice(0, n)).copy() for _ in range(2)]                               # :99-100
    dR = [_ax(d1, dim, slice(1, n + 1)).copy() for _ in range(2)]                           # :103-104
    dL[0] += dx * _ax(d2, dim, slice(0, n))                                                 # :111-112
    dL[1] += dx * _ax(d2, dim, slice(1, n + 1))
    dR[0] -= dx * _ax(d2, dim, slice(1, n + 1))                                             # :117-118
    dR[1] -= dx * _ax(d2, dim, slice(2, n + 2))
    d2abs = np.abs(d2)                                                                      # :137
    smallerL = _ax(d2abs, dim, slice(0, n + 1)) < _ax(d2abs, dim, slice(1, n + 2))          # :140
    smallerR = np.logical_not(smallerL)
    a, b = slice(0, n), slice(1, n + 1)
    derivL = dL[0] * _ax(smallerL, dim, a) + dL[1] * _ax(smallerR, dim, a)                  # :145
    derivR = dR[0] * _ax(smallerL, dim, b) + dR[1] * _ax(smallerR, dim, b)                  # :148
    return derivL, derivR


def upwind_first(grid, data, dim, scheme="as_shipped"):
    """schemeData.CoStateCalc by name: 'as_shipped' / 'intended' (upwindFirstWENO5a), 'eno3a' (upwindFirstENO3a /
    upwindFirstENO3), 'eno2' (upwindFirstENO2)."""
    if scheme == "eno3a":
        return upwind_first_eno3a(grid, data, dim)
    if scheme == "eno2":
        return upwind_first_eno2(grid, data, dim)
    return upwind_first_weno5a(grid, data, dim, scheme)


# --------------------------------------------------------------------------------------
# ExplicitIntegration / Dissipation, Term, Integration
# --------------------------------------------------------------------------------------
class OracleSchemeData:
    """Attribute bag standing in for the reference's Bundle (Utilities/matlab_utils.py:41-57)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def artificial_dissipation_glf(t, data, derivL, derivR, sd):
    """Follows ExplicitIntegration/Dissipation/artificial_diss_glf.py:64-111.

    Returns (diss, stepBound, derivMin, derivMax, alphaMax)."""
    grid = sd.grid
    D = grid.dim
    dmin, dmax, ddiff = [], [], []
    for i in range(D):
        dmin.append(min(np.min(derivL[i]), np.min(derivR[i])))            # :82-84
        dmax.append(max(np.max(derivL[i]), np.max(derivR[i])))            # :86-88
        ddiff.append(derivR[i] - derivL[i])                                # :90
    diss = 0
    sb_inv = 0
    amax = []
    for i in range(D):
        alpha = sd.partialFunc(t, data, dmin, dmax, sd, i)                 # :98
        diss = diss + (0.5 * ddiff[i] * alpha)                             # :100
        if isinstance(alpha, np.ndarray):
            alpha = np.max(alpha)                                          # :104
        amax.append(float(alpha))
        sb_inv = sb_inv + (alpha / float(np.asarray(grid.dx).reshape(-1)[i]))   # :107
    step_bound = float(1 / sb_inv)                                         # :109
    return diss, step_bound, [float(x) for x in dmin], [float(x) for x in dmax], amax


def term_lax_friedrichs(t, y, sd, weno="as_shipped", full=False):
    """ydot = -(H(x, derivC) - diss).  Follows ExplicitIntegration/Term/term_lax_friedrich.py:78-130."""
    grid = sd.grid
    data = np.asarray(y).reshape(grid.shape)                               # :97
    derivL, derivR, derivC = [], [], []
    for i in range(grid.dim):                                              # :106-108
        L, R = upwind_first(grid, data, i, weno)
        derivL.append(L)
        derivR.append(R)
        derivC.append(0.5 * (L + R))
    ham = sd.hamFunc(t, data, derivC, sd)                                  # :111
    if isinstance(ham, tuple):
        ham = ham[0]
    diss, step_bound, dmin, dmax, amax = artificial_dissipation_glf(t, data, derivL, derivR, sd)   # :123
    ydot = np.expand_dims(-(np.asarray(ham) - diss).flatten(), 1)          # :124-128
    if full:
        return ydot, step_bound, dict(derivL=derivL, derivR=derivR, derivMin=dmin, derivMax=dmax, alphaMax=amax)
    return ydot, step_bound


def ode_cfl3_step(t, t_end, y, sd, factor_cfl=0.5, max_step=REALMAX, weno="as_shipped", warn=None):
    """One TVD-RK3 step; the body of the while loop at ode_cfl_3.py:125-251.  Returns (t_new, y_new, dt)."""
    safety = min(1.0, 1.2 * factor_cfl)                                    # :95
    ydot, sb = term_lax_friedrichs(t, y, sd, weno)                         # :129
    dt = float(np.min(np.hstack((factor_cfl * sb, t_end - t, max_step))))  # :142-143
    t1 = t + dt
    y1 = y + dt * ydot                                                     # :151
    ydot, sb = term_lax_friedrichs(t1, y1, sd, weno)                       # :159
    if dt > safety * sb and warn is not None:                              # :173-175
        warn("second", dt / sb)
    t2 = t1 + dt
    y2 = y1 + dt * ydot                                                    # :184
    t_half = 0.25 * (3 * t + t2)                                           # :187
    y_half = 0.25 * (3 * y + y2)                                           # :193
    ydot, sb = term_lax_friedrichs(t_half, y_half, sd, weno)               # :199
    if dt > safety * sb and warn is not None:                              # :215-217
        warn("third", dt / sb)
    t_three_half = t_half + dt
    y_three_half = y_half + dt * ydot                                      # :226
    t_new = (1 / 3) * (t + 2 * t_three_half)                               # :236
    y_new = (1 / 3) * (y + 2 * y_three_half)                               # :241
    return t_new, y_new, dt


def ode_cfl3(tspan, y0, sd, factor_cfl=0.5, max_step=REALMAX, single_step=False, weno="as_shipped"):
    """Follows ExplicitIntegration/Integration/ode_cfl_3.py:79-277 (two-entry tspan, no hooks).

    Returns (t, y, dts)."""
    small = 100 * EPS                                                      # :81
    t = tspan[0]
    y = np.array(y0, dtype=np.float64, copy=True)
    dts = []
    while tspan[1] - t >= small * np.abs(tspan[1]):                        # :125
        t, y, dt = ode_cfl3_step(t, tspan[1], y, sd, factor_cfl, max_step, weno)
        dts.append(dt)
        if single_step:                                                    # :250-251
            break
    return t, y, dts


def term_restrict_update(t, y, sd, positive=True, weno="as_shipped"):
    """termRestrictUpdate around termLaxFriedrichs: ExplicitIntegration/Term/term_restrict_update.py:79-96.
    ydot = max(inner, 0) if positive else min(inner, 0), squeezed to (n,) like the reference (:92,:94)."""
    unrestricted, step_bound = term_lax_friedrichs(t, y, sd, weno)         # :79
    if positive:
        ydot = np.maximum(unrestricted, 0).squeeze()                       # :92
    else:
        ydot = np.minimum(unrestricted, 0).squeeze()                       # :94
    return ydot, step_bound


def ode_cfl2(tspan, y0, sd, factor_cfl=0.5, max_step=REALMAX, single_step=False, weno="as_shipped", restrict=None):
    """ExplicitIntegration/Integration/ode_cfl_2.py (two-entry tspan, no hooks): forward Euler to t+dt, forward Euler
    to t+2dt with the same dt, then the average.  ``restrict``: None -> schemeFunc = termLaxFriedrichs (y of shape
    (n,1)); True/False -> schemeFunc = termRestrictUpdate(positive=restrict) (y of shape (n,), because the restricted
    ydot is squeezed).  Returns (t, y, dts)."""
    small = 100 * EPS
    t = tspan[0]
    y = np.array(y0, dtype=np.float64, copy=True)

    def f(tt, yy):
        if restrict is None:
            return term_lax_friedrichs(tt, yy, sd, weno)
        return term_restrict_update(tt, yy, sd, restrict, weno)
    dts = []
    while tspan[1] - t >= small * np.abs(tspan[1]):
        ydot, sb = f(t, y)
        dt = float(np.min(np.hstack((factor_cfl * sb, tspan[1] - t, max_step))))
        t1 = t + dt
        y1 = y + dt * ydot
        ydot, sb = f(t1, y1)
        t2 = t1 + dt
        y2 = y1 + dt * ydot
        t = 0.5 * (t + t2)
        y = 0.5 * (y + y2)
        dts.append(dt)
        if single_step:
            break
    return t, y, dts


def ode_cfl3_restricted(tspan, y0, sd, positive, factor_cfl=0.5, max_step=REALMAX, single_step=False, weno="as_shipped"):
    """odeCFL3 (ode_cfl_3.py:125-251) with schemeFunc = termRestrictUpdate(positive); y of shape (n,)."""
    small = 100 * EPS
    t = tspan[0]
    y = np.array(y0, dtype=np.float64, copy=True)
    dts = []
    while tspan[1] - t >= small * np.abs(tspan[1]):
        ydot, sb = term_restrict_update(t, y, sd, positive, weno)
        dt = float(np.min(np.hstack((factor_cfl * sb, tspan[1] - t, max_step))))
        t1 = t + dt
        y1 = y + dt * ydot
        ydot, sb = term_restrict_update(t1, y1, sd, positive, weno)
        t2 = t1 + dt
        y2 = y1 + dt * ydot
        t_half = 0.25 * (3 * t + t2)
        y_half = 0.25 * (3 * y + y2)
        ydot, sb = term_restrict_update(t_half, y_half, sd, positive, weno)
        t_three_half = t_half + dt
        y_three_half = y_half + dt * ydot
        t = (1 / 3) * (t + 2 * t_three_half)
        y = (1 / 3) * (y + 2 * y_three_half)
        dts.append(dt)
        if single_step:
            break
    return t, y, dts


def hji_solve(data0, tau, sd, comp_method="minVOverTime", obstacle=None, target=None, weno="as_shipped",
              factor_cfl=0.8, discount=None, discount_mode=None, stop_converge=False, converge_threshold=1e-5):
    """The driver loop of ValueFuncs/hji_solver.py:509-728 in ``keepLast`` mode (the only storage mode
    that works in general): for each tau[i] single-step odeCFL3 until tau[i]-1e-4
    (:536-542, small=1e-4 at :127... used at :536), then the compMethod epilogue (:566-599), discounting (:603-637,
    intended: the shipped branch trips over ``eisfield`` / ``extraArgs.targets``) and the obstacle mask (:641-644,
    intended pointwise max; the shipped ``omax`` returns a scalar, matlab_utils.py:102-112).  ``obstacle`` / ``target``
    with one more dim than the grid are time-varying: slice i serves the interval ending at tau[i] (:642-650; slice 0
    masks data0, :213-222).  ``stop_converge``: stop after the interval whose max |y - y(tau[i-1])| is below the
    threshold (:661-672, :700-726).

    Returns (data, [all dt], [t after every step])."""
    small = 1e-4
    grid = sd.grid
    gdim = len(grid.shape)
    data = np.asarray(data0, dtype=np.float64)
    col = lambda a: np.expand_dims(np.asarray(a, dtype=np.float64).flatten(), 1)
    obs_tv = obstacle is not None and np.ndim(obstacle) == gdim + 1
    tgt_tv = target is not None and np.ndim(target) == gdim + 1
    obs = None if obstacle is None else col(obstacle[0] if obs_tv else obstacle)
    if obstacle is not None:
        data = np.maximum(data, -obs.reshape(grid.shape))                   # :222: data0 is masked before the march
    d0 = col(data)                                                          # 'minVWithV0' sees the masked data0
    tgt = None if target is None else col(target[0] if tgt_tv else target)
    dts, ts = [], []
    for i in range(1, len(tau)):
        y = col(data)                                                       # :532
        y_start = y
        t_now = tau[i - 1]
        if obs_tv:
            obs = col(obstacle[i])                                          # :642-643
        if tgt_tv:
            tgt = col(target[i])                                            # :596, :648-650
        while t_now < tau[i] - small:                                      # :536
            y_last = y
            t_now, y, dt = ode_cfl3(                                       # :542 (singleStep='on', factorCFL=0.8 :445)
                [t_now, tau[i]], y, sd, factor_cfl=factor_cfl, single_step=True, weno=weno)
            dts += dt
            ts.append(t_now)
            if np.any(np.isnan(y)):                                        # :544
                raise ValueError("Nans encountered in the integrated result of HJI PDE data")
            is_min = comp_method in ("minVWithL", "minVwithL", "minVWithTarget")
            is_max = comp_method in ("maxVWithL", "maxVwithL", "maxVWithTarget")
            if discount and discount_mode == "Kene":                        # :615-637
                if tgt is None:
                    raise ValueError("Need to define target function l(x)!")
                max_val = np.max(np.abs(tgt))
                yt = (y - max_val) * discount
                tt = tgt - max_val
                if is_min:
                    yt = np.minimum(yt, tt)
                elif is_max:
                    yt = np.maximum(yt, tt)
                else:
                    raise ValueError("check your compMethod!")
                y = yt + max_val
            else:
                if comp_method in (None, "none", "set", "zero"):
                    pass
                elif comp_method == "minVOverTime":
                    y = np.minimum(y, y_last)                              # :571-573
                elif comp_method == "maxVOverTime":
                    y = np.maximum(y, y_last)
                elif comp_method == "minVWithV0":
                    y = np.minimum(y, d0)
                elif comp_method == "maxVWithV0":
                    y = np.maximum(y, d0)
                elif is_min:                                               # :592
                    y = np.minimum(y, tgt)
                elif is_max:                                               # :583
                    y = np.maximum(y, tgt)
                else:
                    raise ValueError("Check which compMethod you are using")
                if discount:                                               # :603-611
                    y = y * discount
                    y = y + (1 - discount) * (tgt if tgt is not None else d0)
            if obs is not None:
                y = np.maximum(y, -obs)                                    # :641-644 (intended)
        data = y.reshape(grid.shape)                                       # :652
        if stop_converge and np.max(np.abs(y - y_start)) < converge_threshold:   # :672, :700
            hji_solve.last_index = i
            break
    else:
        hji_solve.last_index = len(tau) - 1
    return data, dts, ts
