"""ExplicitIntegration/Integration call surface: ``odeCFLset``, ``odeCFL3`` and ``odeCFL2``."""
import numpy as np

from . import _lib as L
from .term import eng_grid, prepare_scheme, unwrap_scheme
from .utilities import Bundle, cputime, eps, info, isbundle, iscell, isfield, realmax, strcmp, warn

__all__ = ["odeCFLset", "odeCFL3", "odeCFL2", "rk3_times", "rk2_times"]


def odeCFLset(kwargs=None):
    """options = odeCFLset(Bundle(...)) -- ExplicitIntegration/Integration/ode_cfl_set.py:5-133 (same defaults,
    same error behaviour)."""
    if not kwargs:
        raise ValueError("kwargs cannot be None")                      # ode_cfl_set.py:81-89
    assert isbundle(kwargs), "kwargs must be a bundle type."
    d = kwargs.__dict__
    options = Bundle({})
    options.factorCFL = d.get("factorCFL", 0.5)
    options.maxStep = d.get("realmax", realmax)                        # sic: key is 'realmax', :96
    options.postTimeStep = d.get("postTimeStep", None)
    options.singleStep = d.get("singleStep", "off")
    options.stats = d.get("stats", "off")
    options.terminalEvent = d.get("terminalEvent", None)
    if options.factorCFL < 0.0:
        raise ValueError("FactorCFL must be a positive scalar double value")
    if options.maxStep < 0.0:
        raise ValueError("MaxStep must be a positive scalar double value")
    if options.postTimeStep is not None:
        if isinstance(options.postTimeStep, list):
            for f in options.postTimeStep:
                if not callable(f):
                    raise ValueError("Each element in a postTimeStep cell vector must be a function handle.")
        else:
            raise ValueError("PostTimeStep parameter must be a function handle or a list of function handles.")
    if options.terminalEvent is not None and not callable(options.terminalEvent):
        raise ValueError("PostTimeStep parameter must be a function handle.")
    return options


def rk3_times(t, dt):
    """Stage times and the new time, arithmetic verbatim from ode_cfl_3.py:145,178,187,220,236."""
    t1 = t + dt
    t2 = t1 + dt
    tHalf = 0.25 * (3 * t + t2)
    tThreeHalf = tHalf + dt
    tNew = (1 / 3) * (t + 2 * tThreeHalf)
    return t1, tHalf, tNew


def _step_bound(eng, ad, block):
    """stepBound for one RHS: host scalars when the functor carries its alphas (Flock), else the cached device
    maximum of the state-only alpha (artificial_diss_glf.py:104-109)."""
    al = ad.alphas(block)
    if al is None:
        return eng.alpha_max()[1]
    inv = 0
    for d in range(eng.D):
        inv = inv + (al[d] / eng.dx[d])
    return 1 / inv


def _rk_step_dynamic(eng, ad, tables, t, t_end, factorCFL, maxStep, safety, comp, use_obstacle, order):
    """One TVD-RK3 (odeCFL3) or TVD-RK2 (odeCFL2) step for a functor whose alpha depends on the derivative range of the
    field (genericPartial, generic_partial.py:28-56): before each RHS the range of that stage's input is reduced on the
    device (hj_deriv_range), the dynSys's get_opt_u / get_opt_v turn it into the four input sets, and the stage kernel
    runs with that block.  deltaT comes from the first RHS's bound (ode_cfl_3.py:142-143 / ode_cfl_2.py); later bounds
    only warn (ode_cfl_3.py:173-175, :215-217)."""
    if not hasattr(eng, "deriv_range"):
        raise NotImplementedError("genericHam / genericPartial run on single-GPU contexts only")
    stages = (1, 2, 3) if order == 3 else (1, 4)        # 4: the final stage of the RK2 scheme (reads y1)
    deltaT, times, tNew = None, [t] * order, t
    for k, stage in enumerate(stages):
        lo, hi = eng.deriv_range(None, stage)
        eng.set_system(ad.system_id, ad.block_for_range(lo, hi, times[k]), tables)
        bound = eng.alpha_max()[1]                                      # artificial_diss_glf.py:104-109
        if k == 0:
            deltaT = float(np.min(np.hstack((factorCFL * bound, t_end - t, maxStep))))   # ode_cfl_3.py:142-143
            if order == 3:
                t1, tHalf, tNew = rk3_times(t, deltaT)
                times = [t, t1, tHalf]
            else:
                times, tNew = [t, t + deltaT], rk2_times(t, deltaT)
        elif deltaT > safety * bound:                                   # :173-175, :215-217
            warn("%s substep violated CFL effective number %s" % ("Second" if k == 1 else "Third", deltaT / bound))
        last = k == order - 1
        eng.stage(stage, times[k], deltaT, None, comp if last else L.COMP_NONE, use_obstacle and last)
    return tNew, deltaT


def rk2_times(t, dt):
    """New time of one odeCFL2 step, arithmetic verbatim from ode_cfl_2.py (t1 = t+dt; t2 = t1+dt; t = 0.5 (t+t2))."""
    t1 = t + dt
    t2 = t1 + dt
    return 0.5 * (t + t2)


def rk3_step_resident(eng, ad, grid, t, t_end, factorCFL, maxStep, comp=L.COMP_NONE, use_obstacle=False, order=3):
    """One CFL-limited TVD-RK3 (``order=3``, odeCFL3) or TVD-RK2 (``order=2``, odeCFL2) step on the engine's
    resident state.  Returns (t_new, dt)."""
    safetyFactorCFL = min(1.0, 1.2 * factorCFL)                         # ode_cfl_3.py:95
    tables = list(enumerate(ad.tables(grid)))
    if ad.dynamic:
        return _rk_step_dynamic(eng, ad, tables, t, t_end, factorCFL, maxStep, safetyFactorCFL, comp, use_obstacle, order)
    blocks = [ad.block()]
    eng.set_system(ad.system_id, blocks[0], tables)
    stepBound = _step_bound(eng, ad, blocks[0])
    deltaT = float(np.min(np.hstack((factorCFL * stepBound, t_end - t, maxStep))))   # ode_cfl_3.py:142-143
    if ad.time_varying:
        # the reference's hamFunc mutates the system on each of the three RHS evaluations (flock.py:213)
        bounds = [stepBound]
        for _ in range(order - 1):
            b = ad.block()
            blocks.append(b)
            bounds.append(_step_bound(eng, ad, b))
        for k, name in ((1, "Second"), (2, "Third"))[:order - 1]:
            if deltaT > safetyFactorCFL * bounds[k]:                    # ode_cfl_3.py:173-175, :215-217
                warn("%s substep violated CFL effective number %s" % (name, deltaT / bounds[k]))
        (eng.step if order == 3 else eng.step_rk2)(t, deltaT, np.concatenate(blocks), comp, use_obstacle)
    else:
        # state-only alpha: the later stages' bounds equal the stage-1 bound, the CFL check cannot fire
        (eng.step if order == 3 else eng.step_rk2)(t, deltaT, None, comp, use_obstacle)
    return (rk3_times(t, deltaT)[2] if order == 3 else rk2_times(t, deltaT)), deltaT


def _ode_cfl(order, schemeFunc, tspan, y0, options, schemeData):
    small = 100 * eps                                                   # ode_cfl_3.py:81 / ode_cfl_2.py
    if not options:
        options = odeCFLset()                                           # raises, like the reference (:85-86)
    inner, sign = unwrap_scheme(schemeFunc, schemeData)                 # termLaxFriedrichs | termRestrictUpdate(LF)
    if iscell(y0):
        raise NotImplementedError("vector level sets (cell y0) are outside the hot path")
    if isfield(options, "postTimeStep") and options.postTimeStep:
        raise NotImplementedError("postTimeStep hooks would need the field on the host every step")
    if isfield(options, "terminalEvent") and options.terminalEvent:
        raise NotImplementedError("terminalEvent hooks would need the field on the host every step")
    numT = len(tspan)
    if numT != 2:
        raise NotImplementedError("tspan must have exactly two entries (odeCFLmultipleSteps is outside the hot path)")
    eng, ad = prepare_scheme(inner)
    grid = eng_grid(inner)
    t = tspan[0]
    steps = 0
    startTime = cputime()
    single = isfield(options, "singleStep") and strcmp(options.singleStep, "on")
    if (single and order == 3 and isinstance(y0, np.ndarray) and not ad.time_varying
            and tspan[1] - t >= small * np.abs(tspan[1])):
        # the driver's call (hji_solver.py:542: one CFL step per odeCFL3 call on a host array): one C-ABI call that
        # pipelines upload, the three stage kernels and download (hj_ode_cfl3_step); y comes back in one of the
        # engine's pinned arrays (Engine.pinned_out: valid until three further calls)
        eng.set_system(ad.system_id, ad.block(), list(enumerate(ad.tables(grid))))
        y = eng.pinned_out()
        eng.set_restrict(sign)
        try:
            t, _ = eng.ode_cfl3_step(t, tspan[1], options.factorCFL, options.maxStep,
                                     np.ascontiguousarray(y0, dtype=np.float64).reshape(-1), y)
        finally:
            eng.set_restrict(0)
        if isfield(options, "stats") and strcmp(options.stats, "on"):
            info("1 steps in %.2g seconds from  %.2f to %.2f." % (cputime() - startTime, tspan[0], t))
        return t, y.reshape(tuple(y0.shape)), schemeData
    eng.upload(y0)
    eng.set_restrict(sign)
    try:
        while tspan[1] - t >= small * np.abs(tspan[1]):                 # ode_cfl_3.py:125
            t, _ = rk3_step_resident(eng, ad, grid, t, tspan[1], options.factorCFL, options.maxStep, order=order)
            steps += 1
            if isfield(options, "singleStep") and strcmp(options.singleStep, "on"):
                break                                                   # :250-251
    finally:
        eng.set_restrict(0)
    shape = tuple(y0.shape)
    y = eng.download(like=y0, shape=shape)
    endTime = cputime()
    if isfield(options, "stats") and strcmp(options.stats, "on"):
        info("%d steps in %.2g seconds from  %.2f to %.2f." % (steps, endTime - startTime, tspan[0], t))
    return t, y, schemeData


def odeCFL3(schemeFunc, tspan, y0, options=None, schemeData=None):
    """[t, y, schemeData] = odeCFL3(schemeFunc, tspan, y0, options, schemeData)
    -- ExplicitIntegration/Integration/ode_cfl_3.py:11-277: third-order TVD Runge-Kutta with a CFL-limited step.

    Each step is three fused sm_100a stage kernels on a field that stays resident in HBM; ``y0`` is uploaded once
    and ``y`` downloaded once per call (numpy in -> numpy out; torch CUDA tensor in -> tensor out, no host copy).
    ``schemeFunc`` must be ``termLaxFriedrichs`` or ``termRestrictUpdate`` around it (this package's or the
    reference's own function objects)."""
    return _ode_cfl(3, schemeFunc, tspan, y0, options, schemeData)


def odeCFL2(schemeFunc, tspan, y0, options=None, schemeData=None):
    """[t, y, schemeData] = odeCFL2(schemeFunc, tspan, y0, options, schemeData)
    -- ExplicitIntegration/Integration/ode_cfl_2.py: second-order TVD Runge-Kutta (two forward-Euler substeps with the
    CFL step of the first, then the average), two fused stage kernels per step.  Same argument rules as odeCFL3."""
    return _ode_cfl(2, schemeFunc, tspan, y0, options, schemeData)
