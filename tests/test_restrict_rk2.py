"""SURVEY.md 8(f).1 "next" row: termRestrictUpdate (term_restrict_update.py) + odeCFL2 (ode_cfl_2.py).

CPU part: the numpy oracle against the golden fixture made from the LITERAL reference
(tests/golden/make_golden_restrict.py) -- bit-exact.  GPU part: the fused device path against the same fixture:
identical t sequence, restricted ydot within 1e-12 of range, fields within 1e-9 of range (north_star)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import hj_oracle as orc
from oracle import systems as osys


def _case(lsp, gold, name):
    pd = [int(i) for i in np.nonzero(gold[name + "_periodic"])[0]]
    g = lsp.createGrid(gold[name + "_grid_min"], gold[name + "_grid_max"], gold[name + "_grid_N"], pdDims=pd if pd else None)
    if name == "air3d":
        mk = lambda m: m.DubinsVehicleRel(g, float(gold["air3d_u_bound"]), float(gold["air3d_w_bound"]))
    else:
        mk = lambda m: m.DoubleIntegrator(g, float(gold["dint_u_bound"]))
    return g, mk, gold[name + "_data0"]


@pytest.mark.parametrize("name", ["air3d", "dint"])
def test_oracle_restrict_rk2_golden_bit_exact(lsp, name):
    gold = load_golden("restrict_rk2")
    g, mk, d0 = _case(lsp, gold, name)
    yflat = d0.flatten()
    for positive in (True, False):
        tag = "%s_%s" % (name, "pos" if positive else "neg")
        s = mk(osys)
        osd = orc.OracleSchemeData(grid=g, hamFunc=s.hamiltonian, partialFunc=s.dissipation)
        ydot, _ = orc.term_restrict_update(0.0, yflat, osd, positive)
        assert np.array_equal(ydot, gold[tag + "_ydot"])
        assert (ydot >= 0).all() if positive else (ydot <= 0).all()
        for order in (2, 3):
            t, y = 0.0, yflat
            for k in range(3):
                if order == 2:
                    t, y, _ = orc.ode_cfl2([t, 1.0], y, osd, factor_cfl=0.8, single_step=True, restrict=positive)
                else:
                    t, y, _ = orc.ode_cfl3_restricted([t, 1.0], y, osd, positive, factor_cfl=0.8, single_step=True)
                assert t == gold["%s_rk%d_t" % (tag, order)][k]
            assert np.array_equal(y, gold["%s_rk%d_y" % (tag, order)])
    s = mk(osys)
    osd = orc.OracleSchemeData(grid=g, hamFunc=s.hamiltonian, partialFunc=s.dissipation)
    t, y = 0.0, np.expand_dims(yflat, 1)
    for k in range(3):
        t, y, _ = orc.ode_cfl2([t, 1.0], y, osd, factor_cfl=0.8, single_step=True)
        assert t == gold[name + "_plain_rk2_t"][k]
    assert np.array_equal(y, gold[name + "_plain_rk2_y"])


def test_unwrap_scheme_validation(lsp):
    """Host logic (no device): what the restricted term accepts and how `positive` maps to the kernel's sign."""
    from levelsetpy_b200.term import unwrap_scheme
    inner = lsp.Bundle(dict(grid=None))
    assert unwrap_scheme(lsp.termLaxFriedrichs, inner) == (inner, 0)
    sd = lsp.Bundle(dict(innerFunc=lsp.termLaxFriedrichs, innerData=inner))
    assert unwrap_scheme(lsp.termRestrictUpdate, sd) == (inner, 1)           # default positive (:85-88)
    sd.positive = False
    assert unwrap_scheme(lsp.termRestrictUpdate, sd) == (inner, -1)
    with pytest.raises(AssertionError):
        unwrap_scheme(lsp.termRestrictUpdate, lsp.Bundle(dict(innerData=inner)))
    with pytest.raises(NotImplementedError):
        unwrap_scheme(lsp.termRestrictUpdate, lsp.Bundle(dict(innerFunc=lsp.termRestrictUpdate, innerData=inner)))
    with pytest.raises(NotImplementedError):
        unwrap_scheme(lambda *a: None, inner)
    assert lsp.integration.rk2_times(0.25, 0.5) == 0.5 * (0.25 + ((0.25 + 0.5) + 0.5))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["air3d", "dint"])
@pytest.mark.parametrize("backend", ["gather", "tma"])
def test_device_restrict_rk2_vs_golden(lsp, name, backend):
    from levelsetpy_b200 import _lib as L
    gold = load_golden("restrict_rk2")
    g, mk, d0 = _case(lsp, gold, name)
    if backend == "tma" and g.dim < 3:
        pytest.skip("2-D grids run on the gather backend")
    be = L.BACKEND_GATHER if backend == "gather" else L.BACKEND_TMA
    lsp.engine_for_grid(g, "as_shipped").set_backend(be)
    try:
        yflat = d0.flatten()
        opts = lsp.odeCFLset(lsp.Bundle(dict(factorCFL=0.8, singleStep="on")))
        rng_of = lambda a: float(np.max(a) - np.min(a)) or 1.0
        for positive in (True, False):
            tag = "%s_%s" % (name, "pos" if positive else "neg")
            s = mk(lsp)
            inner = lsp.Bundle(dict(grid=g, hamFunc=s.hamiltonian, partialFunc=s.dissipation,
                                    dissFunc=lsp.artificialDissipationGLF, CoStateCalc=lsp.upwindFirstWENO5a))
            sd = lsp.Bundle(dict(innerFunc=lsp.termLaxFriedrichs, innerData=inner, positive=positive))
            ydot, sb, _ = lsp.termRestrictUpdate(0.0, yflat, sd)
            want = gold[tag + "_ydot"]
            assert ydot.shape == want.shape
            assert np.max(np.abs(ydot - want)) <= 1e-12 * rng_of(want)
            assert (ydot >= 0).all() if positive else (ydot <= 0).all()
            for order, fn in ((2, lsp.odeCFL2), (3, lsp.odeCFL3)):
                t, y = 0.0, yflat
                for k in range(3):
                    t, y, _ = fn(lsp.termRestrictUpdate, [t, 1.0], y, opts, sd)
                    assert t == gold["%s_rk%d_t" % (tag, order)][k]
                want = gold["%s_rk%d_y" % (tag, order)]
                assert y.shape == want.shape
                assert np.max(np.abs(y - want)) <= 1e-9 * rng_of(want)
        s = mk(lsp)
        inner = lsp.Bundle(dict(grid=g, hamFunc=s.hamiltonian, partialFunc=s.dissipation,
                                dissFunc=lsp.artificialDissipationGLF, CoStateCalc=lsp.upwindFirstWENO5a))
        t, y = 0.0, np.expand_dims(yflat, 1)
        for k in range(3):
            t, y, _ = lsp.odeCFL2(lsp.termLaxFriedrichs, [t, 1.0], y, opts, inner)
            assert t == gold[name + "_plain_rk2_t"][k]
        want = gold[name + "_plain_rk2_y"]
        assert np.max(np.abs(y - want)) <= 1e-9 * rng_of(want)
        # the restriction is switched off again after the restricted calls
        yd, _, _ = lsp.termLaxFriedrichs(0.0, np.expand_dims(yflat, 1), inner)
        assert (yd > 0).any() and (yd < 0).any()
    finally:
        lsp.engine_for_grid(g, "as_shipped").set_backend(L.BACKEND_AUTO)


@pytest.mark.gpu
def test_device_restrict_on_split_path(lsp):
    """termRestrictUpdate acts on the TOTAL ydot of a product system (pass 2 of the dimension-split path)."""
    from levelsetpy_b200 import _lib as L
    g = lsp.createGrid(-np.ones(4), np.ones(4), np.array([14, 11, 20, 26]))
    x = np.meshgrid(*[np.asarray(v).reshape(-1) for v in g.vs], indexing="ij")
    d0 = np.sqrt((x[0] - x[2]) ** 2 + (x[1] - x[3]) ** 2) - 0.2 + 0.1 * np.sin(3 * x[0] + 2 * x[3])
    s = lsp.ProductSystem(g, [lsp.DoubleIntegrator(g, 1.0), lsp.DoubleIntegrator(g, 0.6)])
    o = osys.ProductSystem([osys.DoubleIntegrator(g, 1.0, dims=(0, 1)), osys.DoubleIntegrator(g, 0.6, dims=(2, 3))])
    osd = orc.OracleSchemeData(grid=g, hamFunc=o.hamiltonian, partialFunc=o.dissipation)
    inner = lsp.Bundle(dict(grid=g, hamFunc=s.hamiltonian, partialFunc=s.dissipation,
                            dissFunc=lsp.artificialDissipationGLF, CoStateCalc=lsp.upwindFirstWENO5a))
    sd = lsp.Bundle(dict(innerFunc=lsp.termLaxFriedrichs, innerData=inner, positive=False))
    opts = lsp.odeCFLset(lsp.Bundle(dict(factorCFL=0.8, singleStep="on")))
    yflat = d0.flatten()
    for order, fn in ((2, lsp.odeCFL2), (3, lsp.odeCFL3)):
        if order == 2:
            to, yo, _ = orc.ode_cfl2([0.0, 1.0], yflat, osd, factor_cfl=0.8, single_step=True, restrict=False)
        else:
            to, yo, _ = orc.ode_cfl3_restricted([0.0, 1.0], yflat, osd, False, factor_cfl=0.8, single_step=True)
        for be in (L.BACKEND_GATHER, L.BACKEND_TMA):
            lsp.engine_for_grid(g, "as_shipped").set_backend(be)
            t, y, _ = fn(lsp.termRestrictUpdate, [0.0, 1.0], yflat, opts, sd)
            assert t == to
            assert np.max(np.abs(y - yo)) <= 1e-9 * float(yo.max() - yo.min())
    lsp.engine_for_grid(g, "as_shipped").set_backend(L.BACKEND_AUTO)


@pytest.mark.gpu
def test_device_hjipde_solve_comp_zero(lsp):
    """HJIPDE_solve(compMethod='zero', extraArgs.restrictUpdate=True): the termRestrictUpdate(positive=0) swap the
    driver sets up at hji_solver.py:438-442 (and, as shipped, never uses: the default 'zero' equals 'set', see
    tests/test_gpu_driver.py) on the resident state, against the oracle's restricted odeCFL3 in the same time loop."""
    gold = load_golden("restrict_rk2")
    g, mk, d0 = _case(lsp, gold, "air3d")
    s = mk(lsp)
    sd = lsp.Bundle(dict(grid=g, hamFunc=s.hamiltonian, partialFunc=s.dissipation))
    tau = np.array([0.0, 0.05, 0.1])
    data, tau_out, extra = lsp.HJIPDE_solve(d0, tau, sd, "zero", lsp.Bundle(dict(quiet=True, keepLast=True,
                                                                                 restrictUpdate=True)))
    o = mk(osys)
    osd = orc.OracleSchemeData(grid=g, hamFunc=o.hamiltonian, partialFunc=o.dissipation)
    y, dts = d0.flatten(), []
    for i in range(1, len(tau)):
        t = float(tau[i - 1])
        while t < tau[i] - 1e-4:
            t_new, y, _ = orc.ode_cfl3_restricted([t, float(tau[i])], y, osd, False, factor_cfl=0.8, single_step=True)
            dts.append(t_new - t)
            t = t_new
    want = y.reshape(g.shape)
    assert extra.steps == len(dts)
    assert np.max(np.abs(data - want)) <= 1e-9 * float(want.max() - want.min())
    assert (data <= d0 + 1e-12).all()          # ydot <= 0 everywhere: the value function can only decrease
    # the restriction does not leak into later calls on the same context
    yd, _, _ = lsp.termLaxFriedrichs(0.0, np.expand_dims(d0.flatten(), 1), lsp.Bundle(dict(
        grid=g, hamFunc=s.hamiltonian, partialFunc=s.dissipation, dissFunc=lsp.artificialDissipationGLF,
        CoStateCalc=lsp.upwindFirstWENO5a)))
    assert (yd > 0).any() and (yd < 0).any()
