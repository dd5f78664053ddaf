"""genericHam / genericPartial (Hamiltonians/generic_ham.py:5-57, generic_partial.py:6-58) over a
``schemeData.dynSys`` -- the hooks HJIPDE_solve installs when a dynSys is present (hji_solver.py:413-415).

CPU part: the numpy restatement (oracle/generic.py) against the fixture written from the LITERAL reference
(tests/golden/make_golden_generic.py) -- bit-exact -- and the host logic of the device path (mode defaults, which dynSys
classes resolve, what raises).  GPU part: the device dynSys functor (csrc/hj_systems.cuh: GenericF<DubinsCarDyn>) against
the same fixture: alphas and stepBound bit for bit, every t identical, ham / ydot within 1e-12 of range, fields within
1e-9 of range with >= 99.99 % sign agreement (north_star)."""
import copy

import numpy as np
import pytest

from conftest import load_golden, make_grid
from oracle import hj_oracle as orc
from oracle import generic as ogen

CASES = (("default", {}), ("umax_dmin", dict(uMode="max", dMode="min")), ("forward", dict(tMode="forward")))


def rng_of(a):
    return float(np.max(a) - np.min(a)) or 1.0


def _dyn_args(gold):
    return dict(speed=float(gold["speed"]), wMax=float(gold["wMax"]), dMax=[float(v) for v in gold["dMax"]])


def _oracle_sd(g, gold, modes):
    return orc.OracleSchemeData(grid=g, dynSys=ogen.DubinsCar(**_dyn_args(gold)), hamFunc=ogen.generic_ham,
                                partialFunc=ogen.generic_partial, **modes)


def _derivs(g, d0):
    dL, dR = [], []
    for d in range(g.dim):
        L, R = orc.upwind_first_weno5a(g, d0, d)
        dL.append(np.ascontiguousarray(L))
        dR.append(np.ascontiguousarray(R))
    return dL, dR, [0.5 * (a + b) for a, b in zip(dL, dR)]


# ------------------------------------------------------------------------------------------------- CPU: oracle pin
@pytest.mark.parametrize("tag,modes", CASES)
def test_oracle_generic_golden_bit_exact(lsp, tag, modes):
    gold = load_golden("generic_dyn")
    g = make_grid(lsp, gold)
    d0 = gold["data0"]
    osd = _oracle_sd(g, gold, modes)
    dL, dR, dC = _derivs(g, d0)
    lo = [min(np.min(dL[d]), np.min(dR[d])) for d in range(3)]
    hi = [max(np.max(dL[d]), np.max(dR[d])) for d in range(3)]
    assert np.array_equal(lo, gold[tag + "_derivMin"]) and np.array_equal(hi, gold[tag + "_derivMax"])
    o1 = copy.copy(osd)
    assert np.array_equal(ogen.generic_ham(0.0, d0, dC, o1), gold[tag + "_ham"])
    for d in range(3):
        a = np.asarray(ogen.generic_partial(0.0, d0, lo, hi, o1, d), dtype=np.float64)
        assert np.array_equal(a, gold[tag + "_alpha%d" % d]), d
    y = np.expand_dims(d0.flatten(), 1)
    ydot, sb = orc.term_lax_friedrichs(0.0, y, osd)
    assert np.array_equal(ydot, gold[tag + "_ydot"])
    assert sb == float(gold[tag + "_stepBound"])
    t = 0.0
    for k in range(3):
        t, y, _ = orc.ode_cfl3([t, 1.0], y, osd, factor_cfl=0.8, single_step=True)
        assert t == gold[tag + "_t"][k]
    assert np.array_equal(y, gold[tag + "_y"])
    t, y = 0.0, np.expand_dims(d0.flatten(), 1)
    for k in range(3):
        t, y, _ = orc.ode_cfl2([t, 1.0], y, osd, factor_cfl=0.8, single_step=True)
        assert t == gold[tag + "_rk2_t"][k]
    assert np.array_equal(y, gold[tag + "_rk2_y"])


def test_oracle_generic_restricted_rk2_golden_bit_exact(lsp):
    gold = load_golden("generic_dyn")
    g = make_grid(lsp, gold)
    osd = _oracle_sd(g, gold, {})
    t, y = 0.0, gold["data0"].flatten()
    for k in range(3):
        t, y, _ = orc.ode_cfl2([t, 1.0], y, osd, factor_cfl=0.8, single_step=True, restrict=False)
        assert t == gold["restrict_neg_rk2_t"][k]
    assert np.array_equal(y, gold["restrict_neg_rk2_y"])


def test_oracle_generic_partial_defaults_to_dmode_min():
    """As shipped the two functions disagree on the dMode default: 'max' in genericHam (generic_ham.py:13-14), 'min' in
    genericPartial (generic_partial.py:19-20); whichever runs first writes it into the bundle."""
    class G:
        xs = [np.zeros(3)] * 3
    sd = orc.OracleSchemeData(grid=G(), dynSys=ogen.DubinsCar(1.0, 1.0, [0.1, 0.2, 0.3]))
    ogen.generic_partial(0.0, None, [-1.0] * 3, [1.0] * 3, sd, 0)
    assert sd.dMode == "min" and sd.uMode == "min"
    sd = orc.OracleSchemeData(grid=G(), dynSys=ogen.DubinsCar(1.0, 1.0, [0.1, 0.2, 0.3]))
    ogen.generic_ham(0.0, None, [np.ones(3)] * 3, sd)
    assert sd.dMode == "max" and sd.uMode == "min" and sd.tMode == "backward"


# ---------------------------------------------------------------------------------------------- CPU: host logic
def test_generic_adapter_host_logic(lsp):
    from levelsetpy_b200 import functors
    from levelsetpy_b200 import _lib as L
    g = lsp.createGrid(np.array([-1.0, -1.0, 0.0]), np.array([1.0, 1.0, 2 * np.pi * (1 - 1 / 9)]), np.array([11, 10, 9]), pdDims=2)
    dyn = lsp.DubinsCar(speed=1.3, wMax=0.9, dMax=[0.15, 0.25, 0.1])
    sd = lsp.Bundle(dict(grid=g, dynSys=dyn, hamFunc=lsp.genericHam, partialFunc=lsp.genericPartial))
    ad = functors.resolve(sd.hamFunc, sd.partialFunc, g, sd)
    assert ad.system_id == L.SYS_GENERIC_DUBINS_CAR and ad.dynamic and ad.time_varying and ad.ndim == 3
    # inside termLaxFriedrichs genericHam runs first: defaults min / max / backward (generic_ham.py:10-17)
    blk = ad.block_for_range([-1.0, -2.0, -3.0], [1.0, 2.0, 3.0], 0.0)
    assert blk.shape == (16,)
    assert list(blk[:3]) == [-1.0, 1.0, -1.0]
    # uU = get_opt_u(derivMax, 'min') = -wMax, uL = +wMax;  dU = +dMax, dL = -dMax for 'max'
    assert list(blk[3:5]) == [-0.9, 0.9]
    assert list(blk[5:8]) == [0.15, 0.25, 0.1] and list(blk[8:11]) == [-0.15, -0.25, -0.1]
    assert list(blk[11:]) == [1.3, 0.9, 0.15, 0.25, 0.1]
    # the Hamiltonian-only block needs no range
    assert list(ad.block_for_range(None, None)[3:11]) == [0.0] * 8
    # explicit modes
    sd2 = lsp.Bundle(dict(grid=g, dynSys=dyn, uMode="max", dMode="min", tMode="forward"))
    blk = functors.generic_adapter(sd2).block_for_range([-1.0] * 3, [1.0] * 3)
    assert list(blk[:5]) == [1.0, -1.0, 1.0, 0.9, -0.9]
    sd3 = lsp.Bundle(dict(grid=g, dynSys=dyn, uMode="sideways"))
    with pytest.raises(ValueError):
        functors.generic_adapter(sd3).block_for_range([-1.0] * 3, [1.0] * 3)
    # tables: numpy trig of the heading axis (what dynSys.dynamics computes with np.cos(x[2]))
    tc, ts = ad.tables(g)
    assert np.array_equal(tc, np.cos(np.asarray(g.vs[2]).reshape(-1))) and np.array_equal(ts, np.sin(np.asarray(g.vs[2]).reshape(-1)))


def test_generic_refusals(lsp):
    from levelsetpy_b200 import functors
    g = lsp.createGrid(np.array([-1.0, -1.0, 0.0]), np.array([1.0, 1.0, 6.0]), np.array([11, 10, 9]), pdDims=2)
    dyn = lsp.DubinsCar()

    class Unknown:
        nx = 3
    # a dynSys with no compiled functor: loud, names what is registered
    with pytest.raises(NotImplementedError, match="DubinsCar"):
        functors.generic_adapter(lsp.Bundle(dict(grid=g, dynSys=Unknown())))
    # no dynSys at all: the reference's own AttributeError (generic_ham.py:8)
    with pytest.raises(AttributeError):
        functors.generic_adapter(lsp.Bundle(dict(grid=g)))
    # mismatched pair
    with pytest.raises(ValueError):
        functors.resolve(lsp.genericHam, lsp.DubinsVehicleRel(g, 1, 1).dissipation, g, lsp.Bundle(dict(grid=g, dynSys=dyn)))
    # reference fields the device path does not cover
    for f in ("uIn", "dIn", "deriv", "side"):
        with pytest.raises(NotImplementedError):
            functors.generic_adapter(lsp.Bundle({"grid": g, "dynSys": lsp.DubinsCar(), f: 1.0}))
    # wrong grid dimension
    g2 = lsp.createGrid(-np.ones(2), np.ones(2), np.array([9, 9]))
    with pytest.raises(ValueError):
        functors.resolve(lsp.genericHam, lsp.genericPartial, g2, lsp.Bundle(dict(grid=g2, dynSys=dyn)))


def test_product_dubins_car_matches_the_test_dynsys():
    """The product's DubinsCar (host methods called on the scalar derivative range only) and the independent test dynSys
    the golden was made with return identical inputs / dynamics."""
    import levelsetpy_b200 as lsp
    a, b = lsp.DubinsCar(1.3, 0.9, [0.15, 0.25, 0.1]), ogen.DubinsCar(1.3, 0.9, [0.15, 0.25, 0.1])
    rng = np.random.default_rng(5)
    p = [rng.standard_normal(7) for _ in range(3)]
    x = [rng.standard_normal(7) for _ in range(3)]
    for m in ("min", "max"):
        assert np.array_equal(a.get_opt_u(0, p, m), b.get_opt_u(0, p, m))
        for va, vb in zip(a.get_opt_v(0, p, m), b.get_opt_v(0, p, m)):
            assert np.array_equal(va, vb)
    u, d = a.get_opt_u(0, p, "min"), a.get_opt_v(0, p, "max")
    for va, vb in zip(a.dynamics(0, x, u, d), b.dynamics(0, x, u, d)):
        assert np.array_equal(va, vb)
    with pytest.raises(ValueError):
        a.get_opt_u(0, p, "sideways")


# ------------------------------------------------------------------------------------------------- GPU: parity
@pytest.mark.gpu
@pytest.mark.parametrize("tag,modes", CASES)
def test_device_generic_hooks_vs_golden(lsp, tag, modes):
    gold = load_golden("generic_dyn")
    g = make_grid(lsp, gold)
    d0 = np.ascontiguousarray(gold["data0"])
    dyn = lsp.DubinsCar(**_dyn_args(gold))
    sd = lsp.Bundle(dict(grid=g, dynSys=dyn, hamFunc=lsp.genericHam, partialFunc=lsp.genericPartial,
                         dissFunc=lsp.artificialDissipationGLF, CoStateCalc=lsp.upwindFirstWENO5a, **modes))
    dL, dR, dC = _derivs(g, d0)
    lo, hi = gold[tag + "_derivMin"], gold[tag + "_derivMax"]
    # the device's own reduction of the derivative range (hj_deriv_range) agrees with the reference's to rounding
    eng = lsp.engine_for_grid(g)
    dlo, dhi = eng.deriv_range(d0)
    assert np.max(np.abs(dlo - lo)) <= 1e-12 * rng_of(hi - lo) and np.max(np.abs(dhi - hi)) <= 1e-12 * rng_of(hi - lo)
    assert np.array_equal(np.sign(dlo), np.sign(lo)) and np.array_equal(np.sign(dhi), np.sign(hi))
    s1 = copy.copy(sd)
    ham = lsp.genericHam(0.0, d0, dC, s1)
    want = gold[tag + "_ham"]
    assert ham.shape == want.shape
    assert np.max(np.abs(ham - want)) <= 1e-12 * rng_of(want)
    assert s1.uMode == modes.get("uMode", "min") and s1.dMode == modes.get("dMode", "max")
    assert s1.tMode == modes.get("tMode", "backward")
    for d in range(3):
        a = lsp.genericPartial(0.0, d0, list(lo), list(hi), s1, d)
        wa = np.broadcast_to(gold[tag + "_alpha%d" % d], g.shape)
        assert np.array_equal(np.asarray(a), wa), d
    # dissFunc on its own with partialFunc = genericPartial: the range comes from the arrays handed in
    diss, sb = lsp.artificialDissipationGLF(0.0, d0, dL, dR, s1)
    osd = _oracle_sd(g, gold, dict(uMode=s1.uMode, dMode=s1.dMode, tMode=s1.tMode))
    wdiss, wsb, _, _, _ = orc.artificial_dissipation_glf(0.0, d0, dL, dR, osd)
    assert sb == wsb == float(gold[tag + "_stepBound"])
    assert np.array_equal(diss, np.broadcast_to(wdiss, g.shape))


@pytest.mark.gpu
@pytest.mark.parametrize("tag,modes", CASES)
def test_device_generic_term_and_ode_vs_golden(lsp, tag, modes):
    gold = load_golden("generic_dyn")
    g = make_grid(lsp, gold)
    d0 = gold["data0"]
    dyn = lsp.DubinsCar(**_dyn_args(gold))
    sd = lsp.Bundle(dict(grid=g, dynSys=dyn, hamFunc=lsp.genericHam, partialFunc=lsp.genericPartial,
                         dissFunc=lsp.artificialDissipationGLF, CoStateCalc=lsp.upwindFirstWENO5a, **modes))
    y = np.expand_dims(d0.flatten(), 1)
    ydot, sb, _ = lsp.termLaxFriedrichs(0.0, y, sd)
    want = gold[tag + "_ydot"]
    assert ydot.shape == want.shape
    assert np.max(np.abs(ydot - want)) <= 1e-12 * rng_of(want)
    assert sb == float(gold[tag + "_stepBound"])
    opts = lsp.odeCFLset(lsp.Bundle(dict(factorCFL=0.8, singleStep="on")))
    t = 0.0
    for k in range(3):
        t, y, _ = lsp.odeCFL3(lsp.termLaxFriedrichs, [t, 1.0], y, opts, sd)
        assert t == gold[tag + "_t"][k]
    want = gold[tag + "_y"]
    assert np.max(np.abs(y - want)) <= 1e-9 * rng_of(want)
    assert np.mean(np.sign(y) == np.sign(want)) >= 0.9999


@pytest.mark.gpu
@pytest.mark.parametrize("tag,modes", CASES)
def test_device_generic_ode_cfl2_vs_golden(lsp, tag, modes):
    """odeCFL2 over the generic hooks: hj_stage(1), hj_stage(4) with the derivative range of y resp. y1 reduced before each."""
    gold = load_golden("generic_dyn")
    g = make_grid(lsp, gold)
    sd = lsp.Bundle(dict(grid=g, dynSys=lsp.DubinsCar(**_dyn_args(gold)), hamFunc=lsp.genericHam,
                         partialFunc=lsp.genericPartial, dissFunc=lsp.artificialDissipationGLF,
                         CoStateCalc=lsp.upwindFirstWENO5a, **modes))
    opts = lsp.odeCFLset(lsp.Bundle(dict(factorCFL=0.8, singleStep="on")))
    t, y = 0.0, np.expand_dims(gold["data0"].flatten(), 1)
    for k in range(3):
        t, y, _ = lsp.odeCFL2(lsp.termLaxFriedrichs, [t, 1.0], y, opts, sd)
        assert t == gold[tag + "_rk2_t"][k]
    want = gold[tag + "_rk2_y"]
    assert y.shape == want.shape
    assert np.max(np.abs(y - want)) <= 1e-9 * rng_of(want)
    assert np.mean(np.sign(y) == np.sign(want)) >= 0.9999


@pytest.mark.gpu
def test_device_generic_restricted_ode_cfl2_vs_golden(lsp):
    """termRestrictUpdate(positive=False) around the generic term under odeCFL2 (what the RCBRT notebooks drive)."""
    gold = load_golden("generic_dyn")
    g = make_grid(lsp, gold)
    inner = lsp.Bundle(dict(grid=g, dynSys=lsp.DubinsCar(**_dyn_args(gold)), hamFunc=lsp.genericHam,
                            partialFunc=lsp.genericPartial, dissFunc=lsp.artificialDissipationGLF,
                            CoStateCalc=lsp.upwindFirstWENO5a))
    sd = lsp.Bundle(dict(innerFunc=lsp.termLaxFriedrichs, innerData=inner, positive=False))
    opts = lsp.odeCFLset(lsp.Bundle(dict(factorCFL=0.8, singleStep="on")))
    t, y = 0.0, gold["data0"].flatten()
    for k in range(3):
        t, y, _ = lsp.odeCFL2(lsp.termRestrictUpdate, [t, 1.0], y, opts, sd)
        assert t == gold["restrict_neg_rk2_t"][k]
    want = gold["restrict_neg_rk2_y"]
    assert y.shape == want.shape
    assert np.max(np.abs(y - want)) <= 1e-9 * rng_of(want)
    # the restriction clamps ydot at 0: nodes that did not move are exactly data0 in both
    assert np.mean(np.sign(y) == np.sign(want)) >= 0.9999


@pytest.mark.gpu
def test_device_generic_on_both_backends(lsp):
    """The generic functor runs in the plane-ring kernel and in the gather kernel: same t, fields equal to rounding."""
    from levelsetpy_b200 import _lib as L
    gold = load_golden("generic_dyn")
    g = make_grid(lsp, gold)
    sd = lsp.Bundle(dict(grid=g, dynSys=lsp.DubinsCar(**_dyn_args(gold)), hamFunc=lsp.genericHam,
                         partialFunc=lsp.genericPartial, dissFunc=lsp.artificialDissipationGLF,
                         CoStateCalc=lsp.upwindFirstWENO5a))
    opts = lsp.odeCFLset(lsp.Bundle(dict(factorCFL=0.8, singleStep="on")))
    eng = lsp.engine_for_grid(g)
    out = {}
    try:
        for be in (L.BACKEND_GATHER, L.BACKEND_TMA):
            eng.set_backend(be)
            t, y = 0.0, np.expand_dims(gold["data0"].flatten(), 1)
            for k in range(3):
                t, y, _ = lsp.odeCFL3(lsp.termLaxFriedrichs, [t, 1.0], y, opts, sd)
                assert t == gold["default_t"][k]
            out[be] = np.array(y, copy=True)
    finally:
        eng.set_backend(L.BACKEND_AUTO)
    want = gold["default_y"]
    for be in out:
        assert np.max(np.abs(out[be] - want)) <= 1e-9 * rng_of(want)
    assert np.max(np.abs(out[L.BACKEND_GATHER] - out[L.BACKEND_TMA])) <= 1e-12 * rng_of(want)


@pytest.mark.gpu
def test_hjipde_solve_installs_generic_hooks_for_a_dynsys(lsp):
    """hji_solver.py:413-415: schemeData.dynSys alone selects genericHam / genericPartial; the driver result matches the
    literal reference's (dt sequence identical)."""
    gold = load_golden("generic_dyn")
    g = make_grid(lsp, gold)
    sd = lsp.Bundle(dict(grid=g, dynSys=lsp.DubinsCar(**_dyn_args(gold)), uMode="min", dMode="max",
                         CoStateCalc=lsp.upwindFirstWENO5a))
    data, tau, extra = lsp.HJIPDE_solve(gold["data0"], gold["hji_tau"], sd, "minVOverTime",
                                        lsp.Bundle(dict(quiet=True, keepLast=True)))
    assert sd.hamFunc is lsp.genericHam and sd.partialFunc is lsp.genericPartial
    want = gold["hji_data"]
    got = np.asarray(data).reshape(want.shape)
    assert np.max(np.abs(got - want)) <= 1e-9 * rng_of(want)
    assert np.mean(np.sign(got) == np.sign(want)) >= 0.9999
    dts = getattr(extra, "dts", None)
    if dts is not None:
        assert list(dts) == list(gold["hji_dts"])


# ------------------------------------------------------------------------- CPU: the two-pass stage protocol on the host
class _RecordingEngine:
    """Stand-in for Engine: records the order of the C-ABI calls one step makes (no device)."""

    def __init__(self, bounds):
        self.calls, self.bounds, self.blocks = [], list(bounds), []

    def deriv_range(self, y=None, stage=1):
        self.calls.append(("deriv_range", stage))
        return np.array([-1.0, -2.0, -3.0]), np.array([1.0, 2.0, 3.0])

    def set_system(self, system_id, block, tables):
        self.calls.append(("set_system", system_id))
        self.blocks.append(np.array(block))

    def alpha_max(self, t=0.0):
        self.calls.append(("alpha_max",))
        return np.zeros(3), self.bounds.pop(0)

    def stage(self, stage, t, dt, params=None, comp=0, use_obstacle=False, want_reduce=False, which_pass=0):
        self.calls.append(("stage", stage, t, dt, comp, bool(use_obstacle)))


@pytest.mark.parametrize("order", [3, 2])
def test_dynamic_step_protocol_on_the_host(lsp, order, caplog):
    """Before EVERY RHS: reduce the range of that stage's input, refresh the block through the dynSys's own methods, take
    the bound; dt from the first bound only (ode_cfl_3.py:142-143), later bounds only warn (:173-175); the driver
    epilogue rides on the last stage; RK2 = stages 1 and 4 with the times of ode_cfl_2.py."""
    import logging
    from levelsetpy_b200 import functors, integration
    from levelsetpy_b200 import _lib as L
    g = lsp.createGrid(np.array([-1.0, -1.0, 0.0]), np.array([1.0, 1.0, 2 * np.pi * (1 - 1 / 9)]), np.array([11, 10, 9]), pdDims=2)
    calls = []

    class Car(lsp.DubinsCar):
        def get_opt_u(self, t, deriv, uMode="min", y=None):
            calls.append(("u", t, tuple(float(v) for v in deriv), uMode))
            return super().get_opt_u(t, deriv, uMode, y)

    Car.__name__ = "DubinsCar"                      # registered by class name
    sd = lsp.Bundle(dict(grid=g, dynSys=Car(1.3, 0.9, [0.15, 0.25, 0.1]), hamFunc=lsp.genericHam, partialFunc=lsp.genericPartial))
    ad = functors.resolve(sd.hamFunc, sd.partialFunc, g, sd)
    eng = _RecordingEngine([0.5] + [0.45] * (order - 2) + [0.1])      # the last bound violates CFL: a warning, not a new dt
    with caplog.at_level(logging.WARNING):
        t_new, dt = integration.rk3_step_resident(eng, ad, g, 0.25, 10.0, 0.8, 1e9, comp=L.COMP_MIN_OVER_TIME, order=order)
    assert dt == 0.8 * 0.5
    stages = (1, 2, 3) if order == 3 else (1, 4)
    want_times = ([0.25] + list(integration.rk3_times(0.25, dt)[:2])) if order == 3 else [0.25, 0.25 + dt]
    assert t_new == (integration.rk3_times(0.25, dt)[2] if order == 3 else integration.rk2_times(0.25, dt))
    per_stage = [eng.calls[4 * k:4 * k + 4] for k in range(order)]
    for k, (st, grp) in enumerate(zip(stages, per_stage)):
        assert [c[0] for c in grp] == ["deriv_range", "set_system", "alpha_max", "stage"], grp
        assert grp[0][1] == st and grp[3][1] == st
        assert grp[3][2] == want_times[k] and grp[3][3] == dt
        assert grp[3][4] == (L.COMP_MIN_OVER_TIME if k == order - 1 else L.COMP_NONE)
    # the dynSys's own get_opt_u saw the reduced range, twice per RHS (derivMax then derivMin), at that RHS's time
    assert len(calls) == 2 * order
    assert calls[0] == ("u", 0.25, (1.0, 2.0, 3.0), "min") and calls[1] == ("u", 0.25, (-1.0, -2.0, -3.0), "min")
    assert [c[1] for c in calls[::2]] == want_times
    msgs = [r.getMessage() for r in caplog.records]
    assert any(("Third" if order == 3 else "Second") + " substep violated CFL" in m for m in msgs), msgs
    assert all(b.shape == (16,) for b in eng.blocks)
