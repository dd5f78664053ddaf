"""Dimension-split path for product systems (hj_vec_kernel.cuh: pass 1 = trailing block with the plane-ring kernel,
pass 2 = leading block on the strided dims) against the numpy oracle and against the single-pass gather backend.
SURVEY.md 8d configs 3 / 4 at oracle-sized grids; tolerances as in test_gpu_parity.py."""
import numpy as np
import pytest

from oracle import hj_oracle as orc
from oracle import systems as osys
from test_gpu_parity import FIELD_TOL, assert_close, scheme

pytestmark = pytest.mark.gpu


def _grid4(lsp, N, pd=None, tz=False):
    g = lsp.createGrid(np.array([-1, -1, -1, -1.]), np.array([1, 1, 1, 1.]), np.array(N), pdDims=pd)
    if tz:
        g.bdryData = [lsp.Bundle(dict(towardZero=True)) for _ in range(4)]
    return g


def _grid6(lsp, N, pd=(2, 5)):
    gmin = [-6, -10, 0, -6, -10, 0.]
    gmax = [20, 10, 2 * np.pi, 20, 10, 2 * np.pi]
    for d in pd:
        gmax[d] = gmin[d] + (gmax[d] - gmin[d]) * (1 - 1 / N[d])
    return lsp.createGrid(np.array(gmin), np.array(gmax), np.array(N), pdDims=list(pd) if pd else None)


def _data(g, rng, kind):
    x = np.meshgrid(*[v.reshape(-1) for v in g.vs], indexing="ij")
    if kind == 4:
        d = np.sqrt((x[0] - x[2]) ** 2 + (x[1] - x[3]) ** 2) - 0.2 + 0.1 * np.sin(3 * x[0] + 2 * x[3])
    else:
        d = np.minimum(np.sqrt(x[0] ** 2 + x[1] ** 2) - 5, np.sqrt(x[3] ** 2 + x[4] ** 2) - 5) + 0.3 * np.sin(x[2] + x[5])
    return np.ascontiguousarray(d + 0.02 * rng.standard_normal(g.shape))


def _systems(lsp, g, kind):
    if kind == 4:
        return (lsp.ProductSystem(g, [lsp.DoubleIntegrator(g, 1.0), lsp.DoubleIntegrator(g, 0.6)]),
                osys.ProductSystem([osys.DoubleIntegrator(g, 1.0, dims=(0, 1)), osys.DoubleIntegrator(g, 0.6, dims=(2, 3))]))
    return (lsp.ProductSystem(g, [lsp.DubinsVehicleRel(g, 5, 1), lsp.DubinsVehicleRel(g, 4, 1.2)]),
            osys.ProductSystem([osys.DubinsVehicleRel(g, 5, 1, dims=(0, 1, 2)), osys.DubinsVehicleRel(g, 4, 1.2, dims=(3, 4, 5))]))


CASES = [
    (4, [20, 19, 37, 45], None, False),        # partial tiles in every dim, odd innermost extent (pad column)
    (4, [9, 40, 56, 30], [1, 3], False),       # periodic dims in both blocks
    (4, [17, 8, 12, 60], [0], True),           # periodic marching dim of pass 2, towardZero extrapolation
    (6, [9, 10, 11, 12, 9, 13], (2, 5), False),
    (6, [8, 16, 7, 9, 15, 44], (), False),     # all extrapolate; two tiles along dims 1 and 5
    (6, [7, 9, 15, 8, 14, 10], (0, 1, 2, 3, 4, 5), False),
]


@pytest.mark.parametrize("kind,N,pd,tz", CASES)
@pytest.mark.parametrize("weno", ["as_shipped", "intended"])
def test_split_path_vs_oracle_and_gather(lsp, kind, N, pd, tz, weno):
    from levelsetpy_b200 import _lib as L
    rng = np.random.default_rng(7)
    g = _grid4(lsp, N, pd, tz) if kind == 4 else _grid6(lsp, N, pd)
    d0 = _data(g, rng, kind)
    s, o = _systems(lsp, g, kind)
    sd = scheme(lsp, g, s, weno)
    osd = orc.OracleSchemeData(grid=g, hamFunc=o.hamiltonian, partialFunc=o.dissipation)
    opts = lsp.odeCFLset(lsp.Bundle(dict(factorCFL=0.8, singleStep="off")))
    y0 = d0.reshape(-1, 1)
    # three CFL steps so that all three RK buffers are recycled through both passes
    sb = orc.term_lax_friedrichs(0.0, y0, osd, weno)[1]
    tspan = [0.0, 2.5 * 0.8 * sb]
    to, yo, _ = orc.ode_cfl3(tspan, y0, osd, factor_cfl=0.8, single_step=False, weno=weno)
    out = {}
    for name, be in (("gather", L.BACKEND_GATHER), ("tma", L.BACKEND_TMA)):
        lsp.engine_for_grid(g, weno).set_backend(be)
        t, y, _ = lsp.odeCFL3(lsp.termLaxFriedrichs, tspan, y0, opts, sd)
        assert t == to
        assert_close(y, yo, FIELD_TOL, "%s vs oracle" % name)
        out[name] = y
    lsp.engine_for_grid(g, weno).set_backend(L.BACKEND_AUTO)
    assert_close(out["tma"], out["gather"], 1e-11, "split vs gather")
    sign_same = np.mean((out["tma"] > 0) == (yo > 0))
    assert sign_same >= 0.9999


@pytest.mark.parametrize("kind,N,pd", [(4, [20, 19, 37, 45], None), (6, [9, 10, 11, 12, 9, 13], (2, 5))])
def test_split_path_reductions_and_epilogues(lsp, kind, N, pd):
    """Reduction record (both passes write disjoint dims of the same record) and the stage-3 driver epilogues."""
    from levelsetpy_b200 import _lib as L
    from levelsetpy_b200.term import prepare_scheme
    rng = np.random.default_rng(3)
    g = _grid4(lsp, N, pd) if kind == 4 else _grid6(lsp, N, pd)
    d0 = _data(g, rng, kind)
    s, o = _systems(lsp, g, kind)
    sd = scheme(lsp, g, s)
    osd = orc.OracleSchemeData(grid=g, hamFunc=o.hamiltonian, partialFunc=o.dissipation)
    _, _, info = orc.term_lax_friedrichs(0.0, d0.reshape(-1, 1), osd, "as_shipped", full=True)
    target = d0 + 0.05
    obstacle = 0.3 - np.abs(d0)
    res = {}
    for be in (L.BACKEND_GATHER, L.BACKEND_TMA):
        eng, ad = prepare_scheme(sd)
        eng.set_backend(be)
        eng.set_system(ad.system_id, ad.block(), list(enumerate(ad.tables(g))))
        eng.upload(d0)
        eng.step(0.0, 1e-3, None, L.COMP_NONE, False, want_reduce=True)
        rec = eng.step_reductions()[0]
        assert np.array_equal(rec["alphaMax"], np.array(info["alphaMax"]))
        assert np.allclose(rec["derivMin"], info["derivMin"], rtol=1e-11, atol=1e-12)
        assert np.allclose(rec["derivMax"], info["derivMax"], rtol=1e-11, atol=1e-12)
        assert not rec["nan"]
        for comp, obs in ((L.COMP_MIN_OVER_TIME, False), (L.COMP_MAX_WITH_AUX, True)):
            eng.upload(d0)
            eng.upload(target, L.FIELD_AUX)
            eng.upload(obstacle, L.FIELD_OBSTACLE)
            eng.step(0.0, 2e-3, None, comp, obs)
            res[(be, comp)] = eng.download().reshape(g.shape)
        eng.set_backend(L.BACKEND_AUTO)
    for comp in (L.COMP_MIN_OVER_TIME, L.COMP_MAX_WITH_AUX):
        assert_close(res[(L.BACKEND_TMA, comp)], res[(L.BACKEND_GATHER, comp)], 1e-12, "epilogue comp=%d" % comp)
    # minVOverTime never increases; the obstacle clamps from below
    assert np.all(res[(L.BACKEND_TMA, L.COMP_MIN_OVER_TIME)] <= d0 + 1e-15)
    assert np.all(res[(L.BACKEND_TMA, L.COMP_MAX_WITH_AUX)] >= -obstacle - 1e-15)


def test_ghost_warp_rings_are_deterministic(lsp):
    """The ghost-warp rings (hj_tma_kernel.cuh / hj_vec_kernel.cuh) order a warp's plain stores of ghost cells into a slot
    before the consumers' loads of them with mbarrier release / acquire only -- no CTA-wide barrier.  compute-sanitizer's
    racecheck does not model that (profiles/README.md), so this is the empirical check: the same TVD-RK3 step repeated 40
    times from the same state, on grids whose every tile touches the boundary, must give the same bits every time (a
    real race on a ghost cell would show up as a run-to-run difference) -- and those bits are the oracle's to 1e-9."""
    from levelsetpy_b200 import _lib as L
    from levelsetpy_b200.integration import rk3_step_resident
    from levelsetpy_b200.term import prepare_scheme
    N = [9, 8, 10, 11, 41, 41]            # trailing planes 41 x 41: the production pass-1 tile (42 x 21, two ghost warps)
    lo = [-6, -10, 0, -6, -10, 0.]
    hi = [20, 10, 2 * np.pi * (1 - 1 / N[2]), 20, 10, 2 * np.pi * (1 - 1 / N[5])]
    g = lsp.createGrid(np.array(lo), np.array(hi), np.array(N), pdDims=[2, 5])
    x = np.meshgrid(*[np.asarray(v).reshape(-1) for v in g.vs], indexing="ij")
    d0 = np.ascontiguousarray(np.minimum(np.sqrt(x[0] ** 2 + x[1] ** 2) - 5, np.sqrt(x[3] ** 2 + x[4] ** 2) - 5)
                              + 0.3 * np.sin(x[2] + x[5]))
    s = lsp.ProductSystem(g, [lsp.DubinsVehicleRel(g, 5, 1), lsp.DubinsVehicleRel(g, 4, 1.2)])
    sd = lsp.Bundle(dict(grid=g, hamFunc=s.hamiltonian, partialFunc=s.dissipation,
                         dissFunc=lsp.artificialDissipationGLF, CoStateCalc=lsp.upwindFirstWENO5a))
    eng, ad = prepare_scheme(sd)
    eng.set_backend(L.BACKEND_TMA)
    first = None
    for rep in range(40):
        eng.upload(d0)
        rk3_step_resident(eng, ad, g, 0.0, 1.0, 0.8, np.finfo(np.float64).max, L.COMP_MIN_OVER_TIME)
        out = eng.download(shape=g.shape)
        if first is None:
            first = out
        else:
            assert np.array_equal(out, first), "run %d differs from run 0" % rep
    eng.set_backend(L.BACKEND_GATHER)
    eng.upload(d0)
    rk3_step_resident(eng, ad, g, 0.0, 1.0, 0.8, np.finfo(np.float64).max, L.COMP_MIN_OVER_TIME)
    ref = eng.download(shape=g.shape)
    eng.set_backend(L.BACKEND_AUTO)
    assert np.max(np.abs(first - ref)) <= 1e-10 * float(ref.max() - ref.min())
