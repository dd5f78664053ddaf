"""HJIPDE_solve (the driver of ValueFuncs/hji_solver.py:24-868) on the device against the numpy oracle's restatement of
the same loop: the FULL horizon of the configs[0] case (north_star: identical dt sequence, <= 1e-9 of the value range
after the full horizon, sign mask >= 99.99 %), every compMethod epilogue, target / obstacle fields, frame stacking,
stopConverge, the as-shipped 'zero' / 'minWithZero' behaviour, lone-Bird systems and the NaN contract."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import hj_oracle as orc
from oracle import systems as osys

pytestmark = pytest.mark.gpu
FIELD_TOL = 1e-9


def rng_of(a):
    return float(np.max(a) - np.min(a)) or 1.0


def air3d(lsp, N, perturb=0.0, seed=3):
    g = lsp.createGrid(np.array([-6.0, -10.0, 0.0]), np.array([20.0, 10.0, 2 * np.pi * (1 - 1 / N[2])]), np.array(N),
                       pdDims=2)
    d0 = lsp.shapeCylinder(g, 2, np.zeros((3, 1)), 5)
    if perturb:
        d0 = d0 + perturb * np.random.default_rng(seed).standard_normal(g.shape)
    return g, np.ascontiguousarray(d0)


def bundles(lsp, g, u=5, w=1):
    s = lsp.DubinsVehicleRel(g, u, w)
    o = osys.DubinsVehicleRel(g, u, w)
    sd = lsp.Bundle(dict(grid=g, hamFunc=s.hamiltonian, partialFunc=s.dissipation))
    osd = orc.OracleSchemeData(grid=g, hamFunc=o.hamiltonian, partialFunc=o.dissipation)
    return sd, osd


def test_full_horizon_air3d_101_vs_oracle(lsp):
    """configs[0] over the whole horizon t in [0, 1] (~60 CFL steps with the minVOverTime clamp active at the zero
    level set): the collapsed 13-flop stencil of the stage kernels differs from the reference's operation order by a
    few ulp per evaluation -- this is what that amounts to after the full march."""
    g, d0 = air3d(lsp, [101, 101, 101])
    sd, osd = bundles(lsp, g)
    tau = np.linspace(0.0, 1.0, 5)
    data, tau_out, extra = lsp.HJIPDE_solve(d0, tau, sd, "minVOverTime", lsp.Bundle(dict(quiet=True, keepLast=True)))
    want, dts, ts = orc.hji_solve(d0, tau, osd, "minVOverTime")
    assert len(dts) >= 40, "the horizon must be tens of steps, got %d" % len(dts)
    assert list(extra.dts) == list(dts), "dt sequence must be identical over the full horizon"
    err = float(np.max(np.abs(data - want)))
    assert err <= FIELD_TOL * rng_of(want), "max abs err %.3e after %d steps (range %.3e)" % (err, len(dts), rng_of(want))
    same = float(np.mean((data < 0) == (want < 0)))
    assert same >= 0.9999, "zero-level-set sign mask agrees on %.6f of the nodes" % same
    assert np.array_equal(tau_out, tau)
    print("full horizon: %d steps, max rel err %.3e, sign mask agreement %.8f" % (len(dts), err / rng_of(want), same))


@pytest.mark.parametrize("comp", ["minVWithTarget", "maxVWithTarget", "minVWithL", "maxVwithL", "minVWithV0",
                                  "maxVWithV0", "maxVOverTime", "set", None])
def test_hjipde_comp_methods_with_target(lsp, comp):
    g, d0 = air3d(lsp, [24, 20, 18], perturb=0.05)
    sd, osd = bundles(lsp, g)
    x = np.meshgrid(*[v.reshape(-1) for v in g.vs], indexing="ij")
    target = np.sqrt((x[0] - 3.0) ** 2 + (x[1] + 1.0) ** 2) - 4.0 + 0.2 * np.cos(x[2])
    tau = np.array([0.0, 0.06, 0.12])
    extra = lsp.Bundle(dict(quiet=True, keepLast=True, targetFunction=target))
    data, _, out = lsp.HJIPDE_solve(d0, tau, sd, comp, extra)
    want, dts, _ = orc.hji_solve(d0, tau, osd, comp, target=target)
    assert list(out.dts) == list(dts)
    assert float(np.max(np.abs(data - want))) <= FIELD_TOL * rng_of(want), comp


def test_hjipde_obstacle_and_target(lsp):
    """Reach-avoid: min with the target, then the obstacle mask max(V, -obstacle) (intended pointwise semantics of
    hji_solver.py:641-644), data0 masked before the march (:222)."""
    g, d0 = air3d(lsp, [24, 20, 18], perturb=0.05)
    sd, osd = bundles(lsp, g)
    x = np.meshgrid(*[v.reshape(-1) for v in g.vs], indexing="ij")
    obstacle = np.sqrt((x[0] - 10.0) ** 2 + (x[1] - 2.0) ** 2) - 3.0 + 0.0 * x[2]
    tau = np.array([0.0, 0.05, 0.1, 0.15])
    for comp, tgt in (("minVWithTarget", d0), ("minVOverTime", None), ("set", None)):
        extra = lsp.Bundle(dict(quiet=True, keepLast=True, obstacleFunction=obstacle))
        if tgt is not None:
            extra.targetFunction = tgt
        data, _, out = lsp.HJIPDE_solve(d0, tau, sd, comp, extra)
        want, dts, _ = orc.hji_solve(d0, tau, osd, comp, obstacle=obstacle, target=tgt)
        assert list(out.dts) == list(dts)
        assert float(np.max(np.abs(data - want))) <= FIELD_TOL * rng_of(want), comp
        assert (data >= -obstacle - 1e-12).all(), "the value never drops below -obstacle"
    with pytest.raises(ValueError):
        lsp.HJIPDE_solve(d0, tau, sd, "minVWithTarget", lsp.Bundle(dict(quiet=True, keepLast=True)))   # no l(x)


def test_hjipde_frames_and_stop_converge(lsp):
    """Non-keepLast storage: frame i is the field at tau[i] (time on axis 0); stopConverge truncates tau."""
    g, d0 = air3d(lsp, [24, 20, 18])
    sd, osd = bundles(lsp, g)
    tau = np.array([0.0, 0.04, 0.08, 0.12])
    data, tau_out, out = lsp.HJIPDE_solve(d0, tau, sd, "minVOverTime", lsp.Bundle(dict(quiet=True)))
    assert data.shape == (len(tau),) + tuple(g.shape)
    assert np.array_equal(data[0], d0)
    for i in range(1, len(tau)):
        want, _, _ = orc.hji_solve(d0, tau[: i + 1], osd, "minVOverTime")
        assert float(np.max(np.abs(data[i] - want))) <= FIELD_TOL * rng_of(want), "frame %d" % i
    # the same march with a huge threshold converges at the first tau
    data2, tau2, _ = lsp.HJIPDE_solve(d0, tau, sd, "minVOverTime",
                                      lsp.Bundle(dict(quiet=True, stopConverge=True, convergeThreshold=1e9)))
    assert len(tau2) == 2 and data2.shape[0] == 2
    assert np.array_equal(data2[1], data[1])
    # ... and with a threshold nothing meets it runs to the end
    data3, tau3, _ = lsp.HJIPDE_solve(d0, tau, sd, "minVOverTime",
                                      lsp.Bundle(dict(quiet=True, stopConverge=True, convergeThreshold=1e-300)))
    assert len(tau3) == len(tau) and np.array_equal(data3, data)


def test_hjipde_zero_is_set_as_shipped(lsp):
    """As shipped the driver's time loop hard-codes termLaxFriedrichs (hji_solver.py:542): 'zero' integrates exactly
    like 'set' and 'minWithZero' ends in error('Check which compMethod you are using') (:599) -- pinned against the
    literal reference in tests/test_reference_shim.py.  extraArgs.restrictUpdate opts into the restricted term."""
    gold = load_golden("hji_air3d_21x17x13")
    g = lsp.createGrid(gold["grid_min"], gold["grid_max"], gold["grid_N"], pdDims=2)
    sd, osd = bundles(lsp, g)
    d0, tau = gold["data0"], gold["tau"]
    q = lsp.Bundle(dict(quiet=True, keepLast=True))
    a, _, ea = lsp.HJIPDE_solve(d0, tau, sd, "zero", q)
    b, _, eb = lsp.HJIPDE_solve(d0, tau, sd, "set", q)
    assert np.array_equal(a, b) and list(ea.dts) == list(eb.dts)
    want, dts, _ = orc.hji_solve(d0, tau, osd, "zero")
    assert list(ea.dts) == list(dts)
    assert float(np.max(np.abs(a - want))) <= FIELD_TOL * rng_of(want)
    with pytest.raises(ValueError):
        lsp.HJIPDE_solve(d0, tau, sd, "minWithZero", q)
    r, _, _ = lsp.HJIPDE_solve(d0, tau, sd, "minWithZero", lsp.Bundle(dict(quiet=True, keepLast=True, restrictUpdate=True)))
    assert (r <= d0 + 1e-12).all() and not np.array_equal(r, a)


def _bird_with_neighbours(mod, g):
    mk = lambda k, w, xyw: mod.Bird(g, 1.0, w, init_xyw=np.array(xyw, dtype=np.float64).reshape(3, 1), label=k, neigh_rad=3)
    b = mk(0, 0.9, [0.15, -0.2, 0.4])
    for n in (mk(1, 1.2, [0.3, 0.1, -0.7]), mk(2, 0.5, [-0.2, 0.25, 1.1])):
        b.update_neighbor(n)
    return b


@pytest.mark.parametrize("pair", [("hamiltonian", "dissipation"), ("hamiltonian_abs", "dissipation_abs")])
def test_lone_bird_vs_oracle(lsp, pair):
    """A Bird on its own grid (bird.py:266-273 / :305-316 with the scalar alphas of :339-344 / :367-372), i.e. the
    'bird' / 'bird_abs' adapter modes of functors.py, against oracle/systems.py."""
    n = 17
    g = lsp.createGrid(np.array([-1.0, -1.0, -np.pi]), np.array([1.0, 1.0, np.pi * (1 - 2 / n)]), np.array([n, n + 2, n]),
                       pdDims=2)
    x = np.meshgrid(*[v.reshape(-1) for v in g.vs], indexing="ij")
    d0 = np.sqrt(x[0] ** 2 + x[1] ** 2) - 0.3 + 0.05 * np.sin(2 * x[2] + x[0])
    b, ob = _bird_with_neighbours(lsp, g), _bird_with_neighbours(osys, g)
    sd = lsp.Bundle(dict(grid=g, hamFunc=getattr(b, pair[0]), partialFunc=getattr(b, pair[1]),
                         dissFunc=lsp.artificialDissipationGLF, CoStateCalc=lsp.upwindFirstWENO5a))
    osd = orc.OracleSchemeData(grid=g, hamFunc=getattr(ob, pair[0]), partialFunc=getattr(ob, pair[1]))
    y0 = d0.reshape(-1, 1)
    ydot, sb, _ = lsp.termLaxFriedrichs(0.0, y0, sd)
    oydot, osb = orc.term_lax_friedrichs(0.0, y0, osd, "as_shipped")
    assert sb == osb
    assert float(np.max(np.abs(ydot - oydot))) <= 1e-12 * rng_of(oydot)
    opts = lsp.odeCFLset(lsp.Bundle(dict(factorCFL=0.8, singleStep="off")))
    t, y, _ = lsp.odeCFL3(lsp.termLaxFriedrichs, [0.0, 0.05], y0, opts, sd)
    to, yo, _ = orc.ode_cfl3([0.0, 0.05], y0, osd, factor_cfl=0.8, single_step=False)
    assert t == to
    assert float(np.max(np.abs(y - yo))) <= FIELD_TOL * rng_of(yo)


def test_nan_propagates_through_the_epilogue(lsp):
    """np.minimum / np.maximum of the driver epilogue propagate NaN (hji_solver.py:571-599) and the driver raises
    'Nans encountered' (:544).  A NaN planted in the field must survive minVOverTime and be reported."""
    g, d0 = air3d(lsp, [24, 20, 18])
    sd, _ = bundles(lsp, g)
    bad = d0.copy()
    bad[12, 10, 9] = np.nan
    with pytest.raises(ValueError, match="Nans"):
        lsp.HJIPDE_solve(bad, np.array([0.0, 0.03]), sd, "minVOverTime", lsp.Bundle(dict(quiet=True, keepLast=True)))


def test_ode_cfl3_single_refuses_flock(lsp):
    """hj_ode_cfl3_single carries no per-stage parameter blocks: it must refuse a Flock (whose hamFunc re-derives the
    headings on each of the three RHS evaluations, flock.py:213) instead of stepping it with a frozen block."""
    from levelsetpy_b200.term import prepare_scheme
    gold = load_golden("flock4_15x15x15")
    from conftest import make_grid
    g = make_grid(lsp, gold)
    birds = [lsp.Bird(g, float(gold["u_bound"]), float(gold["w_bounds"][j]), init_xyw=np.array([gold["init_xyw"][j]]).T.copy(),
                      label=j, neigh_rad=3) for j in range(4)]
    f = lsp.Flock(g, birds)
    sd = lsp.Bundle(dict(grid=g, hamFunc=f.hamiltonian, partialFunc=f.dissipation, dissFunc=lsp.artificialDissipationGLF,
                         CoStateCalc=lsp.upwindFirstWENO5a))
    eng, ad = prepare_scheme(sd)
    eng.set_system(ad.system_id, ad.block(), [])
    y = np.ascontiguousarray(gold["data0"].reshape(-1)).copy()
    with pytest.raises(NotImplementedError):
        eng.ode_cfl3_single(0.0, 1.0, 0.8, np.finfo(np.float64).max, y)


def _fields(g, nt):
    x = np.meshgrid(*[v.reshape(-1) for v in g.vs], indexing="ij")
    s = np.linspace(0.0, 1.0, nt).reshape(-1, 1, 1, 1)
    target = np.sqrt((x[0] - 3.0 - 2.0 * s) ** 2 + (x[1] + 1.0) ** 2) - 4.0 + 0.2 * np.cos(x[2])
    obstacle = np.sqrt((x[0] - 10.0) ** 2 + (x[1] - 2.0 + 3.0 * s) ** 2) - 3.0 + 0.0 * x[2]
    return target, obstacle


def test_hjipde_time_varying_target_and_obstacle(lsp):
    """Moving target / obstacle stacks (len(tau),) + grid.shape: slice i serves the interval ending at tau[i]
    (hji_solver.py:596, :642-650), slice 0 masks data0 (:213-222); static and moving fields mix freely."""
    g, d0 = air3d(lsp, [24, 20, 18], perturb=0.05)
    sd, osd = bundles(lsp, g)
    tau = np.array([0.0, 0.05, 0.1, 0.15])
    tgt, obs = _fields(g, len(tau))
    for comp, t_arg, o_arg in (("minVWithTarget", tgt, obs), ("maxVWithTarget", tgt, obs[1]), ("minVOverTime", None, obs),
                               ("minVWithL", tgt[2], obs)):
        extra = lsp.Bundle(dict(quiet=True, keepLast=True, obstacleFunction=o_arg))
        if t_arg is not None:
            extra.targetFunction = t_arg
        data, _, out = lsp.HJIPDE_solve(d0, tau, sd, comp, extra)
        want, dts, _ = orc.hji_solve(d0, tau, osd, comp, obstacle=o_arg, target=t_arg)
        assert list(out.dts) == list(dts)
        assert float(np.max(np.abs(data - want))) <= FIELD_TOL * rng_of(want), comp
    with pytest.raises(ValueError, match="Inconsistent"):
        lsp.HJIPDE_solve(d0, tau, sd, "minVWithTarget", lsp.Bundle(dict(quiet=True, keepLast=True, targetFunction=tgt[:, :5])))


@pytest.mark.parametrize("mode", [None, "Kene"])
def test_hjipde_discounting(lsp, mode):
    """Discounted value functions (hji_solver.py:603-637): default mode y = gamma y + (1 - gamma) l after the compMethod
    epilogue, 'Kene' mode shifts below zero, discounts, then takes the min / max with the shifted target; the obstacle
    mask follows both (:641-644)."""
    g, d0 = air3d(lsp, [24, 20, 18], perturb=0.05)
    sd, osd = bundles(lsp, g)
    tau = np.array([0.0, 0.05, 0.1])
    tgt, obs = _fields(g, len(tau))
    cases = [("minVWithTarget", tgt[0], None), ("maxVWithL", tgt, obs[0])]
    if mode is None:
        cases += [("minVOverTime", None, None), ("set", None, obs)]          # discount towards data0 (:610)
    for comp, t_arg, o_arg in cases:
        extra = lsp.Bundle(dict(quiet=True, keepLast=True, discountFactor=0.97))
        if mode:
            extra.discountMode = mode
        if t_arg is not None:
            extra.targetFunction = t_arg
        if o_arg is not None:
            extra.obstacleFunction = o_arg
        data, _, out = lsp.HJIPDE_solve(d0, tau, sd, comp, extra)
        want, dts, _ = orc.hji_solve(d0, tau, osd, comp, obstacle=o_arg, target=t_arg, discount=0.97, discount_mode=mode)
        assert list(out.dts) == list(dts)
        assert float(np.max(np.abs(data - want))) <= FIELD_TOL * rng_of(want), (comp, mode)
    if mode == "Kene":
        with pytest.raises(ValueError):
            lsp.HJIPDE_solve(d0, tau, sd, "minVOverTime", lsp.Bundle(dict(quiet=True, keepLast=True, discountFactor=0.9,
                                                                       discountMode="Kene", targetFunction=tgt[0])))


def test_hjipde_stop_conditions(lsp):
    """stopConverge from the device-side change reduction (no frame leaves the GPU for it in keepLast mode), stopInit
    (:676-685) and stopSetIntersect / stopSetInclude (:688-698) against the host evaluation of the same rules on the
    frames of an unstopped run."""
    g, d0 = air3d(lsp, [24, 20, 18])
    sd, osd = bundles(lsp, g)
    tau = np.linspace(0.0, 0.3, 7)
    frames, _, _ = lsp.HJIPDE_solve(d0, tau, sd, "minVOverTime", lsp.Bundle(dict(quiet=True)))
    changes = [float(np.max(np.abs(frames[i] - frames[i - 1]))) for i in range(1, len(tau))]
    thr = 0.5 * (changes[2] + changes[3]) if changes[3] < changes[2] else 1.0001 * max(changes)
    first = next(i for i, c in enumerate(changes, 1) if c < thr)
    data, tau_c, out = lsp.HJIPDE_solve(d0, tau, sd, "minVOverTime",
                                        lsp.Bundle(dict(quiet=True, keepLast=True, stopConverge=True, convergeThreshold=thr)))
    assert len(tau_c) == first + 1 and out.stoptau == tau[first]
    assert np.array_equal(data, frames[first])
    want, _, _ = orc.hji_solve(d0, tau, osd, "minVOverTime", stop_converge=True, converge_threshold=thr)
    assert orc.hji_solve.last_index == first
    assert float(np.max(np.abs(data - want))) <= FIELD_TOL * rng_of(want)
    # stopInit: a state the growing reachable set reaches (value <= 0 at that state, multilinear interpolation)
    from levelsetpy_b200.solver import _interp_at
    cand = np.argwhere((frames[0] > 0) & (frames[-1] <= 0))            # nodes the growing reachable set swallows
    assert len(cand) > 0
    node = cand[len(cand) // 2]
    p = np.array([float(np.asarray(g.vs[d]).reshape(-1)[node[d]]) for d in range(3)])
    vals = [_interp_at(g, frames[i], p) for i in range(len(tau))]
    assert vals[0] > 0 and vals[-1] <= 0, vals
    assert vals[2] == frames[2][tuple(node)]                           # at a node the interpolant is the node value
    hit = next(i for i in range(1, len(tau)) if vals[i] <= 0)
    _, tau_i, out_i = lsp.HJIPDE_solve(d0, tau, sd, "minVOverTime", lsp.Bundle(dict(quiet=True, keepLast=True, stopInit=p)))
    assert len(tau_i) == hit + 1 and out_i.stoptau == tau[hit]
    with pytest.raises(ValueError):
        lsp.HJIPDE_solve(d0, tau, sd, "minVOverTime", lsp.Bundle(dict(quiet=True, stopInit=p[:2])))
    # stop sets: a small ball around that state; Intersect fires when any of its nodes is inside, Include when all are
    x = np.meshgrid(*[v.reshape(-1) for v in g.vs], indexing="ij")
    ball = np.sqrt((x[0] - p[0]) ** 2 + (x[1] - p[1]) ** 2) - 1.2 + 0.0 * x[2]
    inside = [frames[i][ball < 0] <= 0 for i in range(len(tau))]
    any_i = next((i for i in range(1, len(tau)) if inside[i].any()), None)
    all_i = next((i for i in range(1, len(tau)) if inside[i].all()), None)
    assert any_i is not None
    _, tau_s, _ = lsp.HJIPDE_solve(d0, tau, sd, "minVOverTime", lsp.Bundle(dict(quiet=True, keepLast=True, stopSetIntersect=ball)))
    assert len(tau_s) == any_i + 1
    _, tau_a, _ = lsp.HJIPDE_solve(d0, tau, sd, "minVOverTime", lsp.Bundle(dict(quiet=True, keepLast=True, stopSetInclude=ball)))
    assert len(tau_a) == (all_i + 1 if all_i is not None else len(tau))
    assert all_i is None or all_i >= any_i
