"""SURVEY.md 8(f).2 "next" row: upwindFirstENO2 (upwind_first_eno2.py) and upwindFirstENO3a / upwindFirstENO3
(upwind_first_eno3a.py, upwind_first_eno3.py) as schemeData.CoStateCalc.

CPU part: the numpy oracle against the golden fixture made from the LITERAL reference
(tests/golden/make_golden_eno.py) -- bit-exact -- plus known-answer properties.  GPU part: the device functors
(gather backend; standalone operator and fused stage kernels) against the same fixture: derivatives and ydot within
1e-12 of range (the minimum-modulus choices are discontinuous, so any wrong choice is an O(dx^2) error, far above
that), stepBound and every t identical, fields within 1e-9 of range (north_star)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import hj_oracle as orc
from oracle import systems as osys

SCHEMES = ("eno2", "eno3a")


def _case(lsp, gold, name):
    pd = [int(i) for i in np.nonzero(gold[name + "_periodic"])[0]]
    g = lsp.createGrid(gold[name + "_grid_min"], gold[name + "_grid_max"], gold[name + "_grid_N"], pdDims=pd if pd else None)
    if name == "air3d":
        mk = lambda m: m.DubinsVehicleRel(g, float(gold["air3d_u_bound"]), float(gold["air3d_w_bound"]))
    else:
        mk = lambda m: m.DoubleIntegrator(g, float(gold["dint_u_bound"]))
    return g, mk, gold[name + "_data0"]


@pytest.mark.parametrize("name", ["air3d", "dint"])
@pytest.mark.parametrize("scheme", SCHEMES)
def test_oracle_eno_golden_bit_exact(lsp, name, scheme):
    gold = load_golden("eno_schemes")
    g, mk, d0 = _case(lsp, gold, name)
    fn = orc.upwind_first_eno2 if scheme == "eno2" else orc.upwind_first_eno3a
    for d in range(g.dim):
        L, R = fn(g, d0, d)
        assert np.array_equal(L, gold["%s_%s_L%d" % (name, scheme, d)])
        assert np.array_equal(R, gold["%s_%s_R%d" % (name, scheme, d)])
    s = mk(osys)
    osd = orc.OracleSchemeData(grid=g, hamFunc=s.hamiltonian, partialFunc=s.dissipation)
    y = np.expand_dims(d0.flatten(), 1)
    ydot, sb = orc.term_lax_friedrichs(0.0, y, osd, scheme)
    assert np.array_equal(ydot, gold["%s_%s_ydot" % (name, scheme)])
    assert sb == float(gold["%s_%s_stepBound" % (name, scheme)])
    t = 0.0
    for k in range(3):
        t, y, _ = orc.ode_cfl3([t, 1.0], y, osd, factor_cfl=0.8, single_step=True, weno=scheme)
        assert t == gold["%s_%s_t" % (name, scheme)][k]
    assert np.array_equal(y, gold["%s_%s_y" % (name, scheme)])


@pytest.mark.parametrize("scheme,degree", [("eno2", 2), ("eno3a", 3)])
def test_oracle_eno_polynomial_exactness(lsp, scheme, degree):
    """Known answer: every candidate of an order-p ENO scheme differentiates polynomials of degree <= p exactly, so the
    chosen one does too -- away from extrapolated ghost cells (periodic data would not be polynomial: interior only)."""
    g = lsp.createGrid(np.array([-1.0, -1.0]), np.array([1.0, 1.0]), np.array([41, 9]))
    x = np.asarray(g.vs[0]).reshape(-1, 1) + 0 * np.asarray(g.vs[1]).reshape(1, -1)
    c = [0.3, -1.1, 0.7, 0.45]
    f = sum(c[k] * x ** k for k in range(degree + 1))
    df = sum(k * c[k] * x ** (k - 1) for k in range(1, degree + 1))
    fn = orc.upwind_first_eno2 if scheme == "eno2" else orc.upwind_first_eno3a
    L, R = fn(g, f, 0)
    inner = slice(4, -4)
    assert np.max(np.abs(L[inner] - df[inner])) < 1e-12
    assert np.max(np.abs(R[inner] - df[inner])) < 1e-12


def test_costate_names_select_the_scheme(lsp):
    """Host logic (no device): CoStateCalc is recognised by name, like every other callable of the bundle."""
    from levelsetpy_b200.engine import weno_mode_of
    for fn, want in ((lsp.upwindFirstENO2, "eno2"), (lsp.upwindFirstENO3a, "eno3a"), (lsp.upwindFirstENO3, "eno3a"),
                     (lsp.upwindFirstWENO5a, "as_shipped"), (lsp.upwindFirstWENO5, "as_shipped")):
        assert weno_mode_of(lsp.Bundle(dict(CoStateCalc=fn))) == want
    assert weno_mode_of(lsp.Bundle(dict(CoStateCalc=lsp.upwindFirstWENO5a, wenoMode="intended"))) == "intended"
    assert weno_mode_of(lsp.Bundle(dict(CoStateCalc=lsp.upwindFirstENO2, wenoMode="intended"))) == "eno2"
    with pytest.raises(ValueError):
        weno_mode_of(lsp.Bundle(dict(CoStateCalc=lsp.upwindFirstWENO5a, wenoMode="eno2")))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["air3d", "dint"])
@pytest.mark.parametrize("scheme", SCHEMES)
def test_device_eno_vs_golden(lsp, name, scheme):
    gold = load_golden("eno_schemes")
    g, mk, d0 = _case(lsp, gold, name)
    fn = lsp.upwindFirstENO2 if scheme == "eno2" else lsp.upwindFirstENO3a
    rng_of = lambda a: float(np.max(a) - np.min(a)) or 1.0
    for d in range(g.dim):
        L, R = fn(g, d0, d)
        for got, key in ((L, "L"), (R, "R")):
            want = gold["%s_%s_%s%d" % (name, scheme, key, d)]
            assert got.shape == want.shape
            assert np.max(np.abs(got - want)) <= 1e-12 * rng_of(want), (key, d)
    if scheme == "eno3a":
        L3, _ = lsp.upwindFirstENO3(g, d0, 0)
        assert np.array_equal(L3, fn(g, d0, 0)[0])
    s = mk(lsp)
    sd = lsp.Bundle(dict(grid=g, hamFunc=s.hamiltonian, partialFunc=s.dissipation,
                         dissFunc=lsp.artificialDissipationGLF, CoStateCalc=fn))
    y = np.expand_dims(d0.flatten(), 1)
    ydot, sb, _ = lsp.termLaxFriedrichs(0.0, y, sd)
    want = gold["%s_%s_ydot" % (name, scheme)]
    assert ydot.shape == want.shape
    assert np.max(np.abs(ydot - want)) <= 1e-12 * rng_of(want)
    assert sb == float(gold["%s_%s_stepBound" % (name, scheme)])
    opts = lsp.odeCFLset(lsp.Bundle(dict(factorCFL=0.8, singleStep="on")))
    t = 0.0
    for k in range(3):
        t, y, _ = lsp.odeCFL3(lsp.termLaxFriedrichs, [t, 1.0], y, opts, sd)
        assert t == gold["%s_%s_t" % (name, scheme)][k]
    want = gold["%s_%s_y" % (name, scheme)]
    assert np.max(np.abs(y - want)) <= 1e-9 * rng_of(want)
    assert np.mean(np.sign(y) == np.sign(want)) >= 0.9999


@pytest.mark.gpu
@pytest.mark.parametrize("scheme", ["eno2", "eno3a"])
def test_device_eno_on_both_backends(lsp, scheme):
    """The ENO functors run in the plane-ring (TMA) kernel too for whole 3-D systems: the same device functions build the
    divided-difference tables and make the minimum-modulus choices, so the two backends agree to rounding of the
    Hamiltonian / stage algebra (a flipped choice would be an O(dx^2) difference), and both match the
    reference-generated golden after three odeCFL3 steps."""
    from levelsetpy_b200 import _lib as L
    gold = load_golden("eno_schemes")
    g, mk, d0 = _case(lsp, gold, "air3d")
    s = mk(lsp)
    fn = lsp.upwindFirstENO2 if scheme == "eno2" else lsp.upwindFirstENO3a
    sd = lsp.Bundle(dict(grid=g, hamFunc=s.hamiltonian, partialFunc=s.dissipation,
                         dissFunc=lsp.artificialDissipationGLF, CoStateCalc=fn))
    opts = lsp.odeCFLset(lsp.Bundle(dict(factorCFL=0.8, singleStep="on")))
    eng = lsp.engine_for_grid(g, scheme)
    out = {}
    try:
        for be in (L.BACKEND_GATHER, L.BACKEND_TMA):
            eng.set_backend(be)
            t, y = 0.0, np.expand_dims(d0.flatten(), 1)
            for k in range(3):
                t, y, _ = lsp.odeCFL3(lsp.termLaxFriedrichs, [t, 1.0], y, opts, sd)
                assert t == gold["air3d_%s_t" % scheme][k]
            out[be] = np.array(y, copy=True)
    finally:
        eng.set_backend(L.BACKEND_AUTO)
    want = gold["air3d_%s_y" % scheme]
    rng = float(want.max() - want.min())
    assert np.max(np.abs(out[L.BACKEND_TMA] - want)) <= 1e-9 * rng
    assert np.max(np.abs(out[L.BACKEND_GATHER] - out[L.BACKEND_TMA])) <= 1e-12 * rng
