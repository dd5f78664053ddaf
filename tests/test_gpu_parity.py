"""Parity of the CUDA path (through the C-ABI) against the golden fixtures (= literal reference outputs) and the
numpy oracle.  Tolerances, from BASELINE.json's north_star:
  * fields: max |diff| <= 1e-9 * value range (we hold derivatives/ydot to 1e-12 of their range per evaluation);
  * dt / t sequences: identical (==);
  * zero-level-set sign mask identical on >= 99.99 % of nodes;
  * ghost cells: bit-exact.
"""
import numpy as np
import pytest

from conftest import load_golden, make_grid
from oracle import hj_oracle as orc
from oracle import systems as osys

pytestmark = pytest.mark.gpu

FIELD_TOL = 1e-9     # north_star: relative to value range, after the full horizon
EVAL_TOL = 1e-12     # single operator evaluation, relative to output range


def rng_of(a):
    return float(np.max(a) - np.min(a)) or 1.0


def assert_close(got, want, tol, what):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, what
    err = float(np.max(np.abs(got - want)))
    assert err <= tol * rng_of(want), "%s: max abs err %.3e > %.1e * range %.3e" % (what, err, tol, rng_of(want))


def system_for(lsp, name, g, gold):
    if name.startswith("air3d"):
        return lsp.DubinsVehicleRel(g, float(gold["u_bound"]), float(gold["w_bound"]))
    if name.startswith("dint"):
        return lsp.DoubleIntegrator(g, float(gold["u_bound"]))
    if name.startswith("flock"):
        birds = [lsp.Bird(g, float(gold["u_bound"]), float(gold["w_bounds"][j]),
                          init_xyw=np.array([gold["init_xyw"][j]]).T.copy(), label=j, neigh_rad=3) for j in range(4)]
        return lsp.Flock(g, birds)
    raise KeyError(name)


def scheme(lsp, g, system, weno="as_shipped"):
    return lsp.Bundle(dict(grid=g, hamFunc=system.hamiltonian, partialFunc=system.dissipation,
                           dissFunc=lsp.artificialDissipationGLF, CoStateCalc=lsp.upwindFirstWENO5a, wenoMode=weno))


CASES = ["air3d_21x17x13", "air3d_cyl_21x17x13", "dint_51x51", "dint_33x20", "flock4_15x15x15"]


def test_ghost_cells_bit_exact(lsp):
    gold = load_golden("ghost_cells_6x5x7")
    a = gold["a"]
    for d in range(3):
        for tz in (False, True):
            out = lsp.addGhostExtrapolate(a, d, 3, lsp.Bundle(dict(towardZero=tz)))
            assert np.array_equal(out, gold["extrap_d%d_tz%d" % (d, int(tz))])
        assert np.array_equal(lsp.addGhostPeriodic(a, d, 3, None), gold["periodic_d%d" % d])
    # width 1 default and width 2
    assert np.array_equal(lsp.addGhostPeriodic(a, 1), orc.add_ghost_periodic(a, 1, 1))
    assert np.array_equal(lsp.addGhostExtrapolate(a, 2, 2), orc.add_ghost_extrapolate(a, 2, 2))
    with pytest.raises(ValueError):
        lsp.addGhostExtrapolate(a, 1, 9)


@pytest.mark.parametrize("name", CASES)
def test_upwind_first_weno5a(lsp, name):
    gold = load_golden(name)
    g = make_grid(lsp, gold)
    assert np.array_equal(np.asarray(g.dx).reshape(-1), gold["grid_dx"])
    for d in range(g.dim):
        L, R = lsp.upwindFirstWENO5a(g, gold["data0"], d)
        scale = max(rng_of(gold["derivL%d" % d]), rng_of(gold["derivR%d" % d]), 1e-300)
        assert float(np.max(np.abs(L - gold["derivL%d" % d]))) <= EVAL_TOL * scale + 1e-13
        assert float(np.max(np.abs(R - gold["derivR%d" % d]))) <= EVAL_TOL * scale + 1e-13
        L2, R2 = lsp.upwindFirstWENO5(g, gold["data0"], d)
        assert np.array_equal(L, L2) and np.array_equal(R, R2)
        Li, Ri = lsp.upwindFirstWENO5a(g, gold["data0"], d, wenoMode="intended")
        scale = max(rng_of(gold["intended_derivL%d" % d]), 1e-300)
        assert float(np.max(np.abs(Li - gold["intended_derivL%d" % d]))) <= 1e-11 * scale + 1e-13
        assert float(np.max(np.abs(Ri - gold["intended_derivR%d" % d]))) <= 1e-11 * scale + 1e-13


@pytest.mark.parametrize("name", CASES)
def test_term_lax_friedrichs(lsp, name):
    gold = load_golden(name)
    g = make_grid(lsp, gold)
    sd = scheme(lsp, g, system_for(lsp, name, g, gold))
    y0 = gold["data0"].reshape(-1, 1)
    ydot, sb, _ = lsp.termLaxFriedrichs(0.0, y0, sd)
    assert ydot.shape == gold["ydot"].shape
    assert sb == float(gold["stepBound"]), "stepBound must be identical (dt sequence)"
    assert_close(ydot, gold["ydot"], EVAL_TOL, name + " ydot")


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("backend", ["gather", "auto"])
def test_ode_cfl3_steps(lsp, name, backend):
    from levelsetpy_b200 import _lib as L
    gold = load_golden(name)
    g = make_grid(lsp, gold)
    sd = scheme(lsp, g, system_for(lsp, name, g, gold))
    lsp.engine_for_grid(g).set_backend(L.BACKEND_GATHER if backend == "gather" else L.BACKEND_AUTO)
    opts = lsp.odeCFLset(lsp.Bundle(dict(factorCFL=0.8, singleStep="on")))
    t, y = 0.0, gold["data0"].reshape(-1, 1)
    for k, t_want in enumerate(gold["t_steps"]):
        t, y, _ = lsp.odeCFL3(lsp.termLaxFriedrichs, [t, 1.0], y, opts, sd)
        assert t == t_want, "t after step %d: %r != %r" % (k, t, t_want)
    assert_close(y, gold["y_final"], FIELD_TOL, name + " y after %d steps" % len(gold["t_steps"]))
    lsp.engine_for_grid(g).set_backend(L.BACKEND_AUTO)


@pytest.mark.parametrize("name", ["air3d_21x17x13", "dint_51x51"])
def test_ode_cfl3_intended(lsp, name):
    gold = load_golden(name)
    g = make_grid(lsp, gold)
    sd = scheme(lsp, g, system_for(lsp, name, g, gold), weno="intended")
    opts = lsp.odeCFLset(lsp.Bundle(dict(factorCFL=0.8, singleStep="off")))
    t, y, _ = lsp.odeCFL3(lsp.termLaxFriedrichs, [0.0, float(gold["t_steps"][-1])], gold["data0"].reshape(-1, 1), opts, sd)
    assert t == gold["t_steps"][-1]
    assert_close(y, gold["intended_y_final"], FIELD_TOL, name + " intended")


def test_hjipde_solve_matches_reference(lsp):
    gold = load_golden("hji_air3d_21x17x13")
    g = lsp.createGrid(gold["grid_min"], gold["grid_max"], gold["grid_N"], pdDims=2)
    sysd = lsp.DubinsVehicleRel(g, 5, 1)
    sd = lsp.Bundle(dict(grid=g, hamFunc=sysd.hamiltonian, partialFunc=sysd.dissipation))
    data, tau, extra = lsp.HJIPDE_solve(gold["data0"], gold["tau"], sd, "minVOverTime",
                                        lsp.Bundle(dict(quiet=True, keepLast=True)))
    assert list(extra.dts) == list(gold["dts"]), "dt sequence must be identical"
    assert_close(data, gold["data"], FIELD_TOL, "HJIPDE_solve data")
    same_sign = np.mean(np.sign(data) == np.sign(gold["data"]))
    assert same_sign >= 0.9999


def test_torch_tensor_roundtrip(lsp):
    torch = pytest.importorskip("torch")
    gold = load_golden("air3d_21x17x13")
    g = make_grid(lsp, gold)
    sd = scheme(lsp, g, system_for(lsp, "air3d", g, gold))
    y0 = torch.from_numpy(gold["data0"].reshape(-1, 1)).cuda()
    ydot, sb, _ = lsp.termLaxFriedrichs(0.0, y0, sd)
    assert ydot.is_cuda and sb == float(gold["stepBound"])
    assert_close(ydot.cpu().numpy(), gold["ydot"], EVAL_TOL, "ydot (torch)")


def test_product_systems_vs_oracle(lsp):
    """4-D double-integrator pair and 6-D relative-Dubins pair at oracle-sized grids."""
    rng = np.random.default_rng(5)
    # 4-D
    g4 = lsp.createGrid(np.array([-1, -1, -1, -1.]), np.array([1, 1, 1, 1.]), np.array([13, 11, 12, 10]))
    x = np.meshgrid(*[v.reshape(-1) for v in g4.vs], indexing="ij")
    d4 = np.sqrt((x[0] - x[2]) ** 2 + (x[1] - x[3]) ** 2) - 0.2 + 0.02 * rng.standard_normal(g4.shape)
    s4 = lsp.ProductSystem(g4, [lsp.DoubleIntegrator(g4, 1.0), lsp.DoubleIntegrator(g4, 0.6)])
    o4 = osys.ProductSystem([osys.DoubleIntegrator(g4, 1.0, dims=(0, 1)), osys.DoubleIntegrator(g4, 0.6, dims=(2, 3))])
    # 6-D
    N6 = [7, 8, 9, 8, 7, 10]
    g6 = lsp.createGrid(np.array([-6, -10, 0, -6, -10, 0.]),
                        np.array([20, 10, 2 * np.pi * (1 - 1 / N6[2]), 20, 10, 2 * np.pi * (1 - 1 / N6[5])]),
                        np.array(N6), pdDims=[2, 5])
    x = np.meshgrid(*[v.reshape(-1) for v in g6.vs], indexing="ij")
    d6 = np.minimum(np.sqrt(x[0] ** 2 + x[1] ** 2) - 5, np.sqrt(x[3] ** 2 + x[4] ** 2) - 5) \
        + 0.02 * rng.standard_normal(g6.shape)
    s6 = lsp.ProductSystem(g6, [lsp.DubinsVehicleRel(g6, 5, 1), lsp.DubinsVehicleRel(g6, 4, 1.2)])
    o6 = osys.ProductSystem([osys.DubinsVehicleRel(g6, 5, 1, dims=(0, 1, 2)), osys.DubinsVehicleRel(g6, 4, 1.2, dims=(3, 4, 5))])
    for g, d0, s, o in ((g4, d4, s4, o4), (g6, d6, s6, o6)):
        for weno in ("as_shipped", "intended"):
            sd = scheme(lsp, g, s, weno)
            osd = orc.OracleSchemeData(grid=g, hamFunc=o.hamiltonian, partialFunc=o.dissipation)
            y0 = d0.reshape(-1, 1)
            ydot, sb, _ = lsp.termLaxFriedrichs(0.0, y0, sd)
            oydot, osb = orc.term_lax_friedrichs(0.0, y0, osd, weno)
            assert sb == osb
            assert_close(ydot, oydot, 1e-11, "%d-D ydot %s" % (g.dim, weno))
            opts = lsp.odeCFLset(lsp.Bundle(dict(factorCFL=0.8, singleStep="on")))
            t, y, _ = lsp.odeCFL3(lsp.termLaxFriedrichs, [0.0, 1.0], y0, opts, sd)
            to, yo, _ = orc.ode_cfl3([0.0, 1.0], y0, osd, factor_cfl=0.8, single_step=True, weno=weno)
            assert t == to
            assert_close(y, yo, FIELD_TOL, "%d-D step %s" % (g.dim, weno))


def _air_case(lsp, N, pd, seed=11, tz=False):
    rng = np.random.default_rng(seed)
    gmax = [20.0, 10.0, 2 * np.pi]
    gmin = [-6.0, -10.0, 0.0]
    for d in pd:
        gmax[d] = gmin[d] + (gmax[d] - gmin[d]) * (1 - 1 / N[d])
    g = lsp.createGrid(np.array(gmin), np.array(gmax), np.array(N), pdDims=pd if pd else None)
    if tz:
        g.bdryData = [lsp.Bundle(dict(towardZero=True)) for _ in range(3)]
    x = np.meshgrid(*[v.reshape(-1) for v in g.vs], indexing="ij")
    d0 = np.sqrt(x[0] ** 2 + x[1] ** 2) - 5 + 0.4 * np.sin(x[2] + 0.2 * x[0]) + 0.05 * rng.standard_normal(g.shape)
    return g, np.ascontiguousarray(d0)


@pytest.mark.parametrize("N,pd,tz", [([70, 40, 50], [2], False), ([37, 45, 34], [], True), ([40, 33, 66], [0, 1, 2], False),
                                      ([36, 70, 31], [1], False)])
@pytest.mark.parametrize("weno", ["as_shipped", "intended"])
def test_tma_ring_kernel_vs_oracle_and_gather(lsp, N, pd, tz, weno):
    """Multi-chunk, partial-tile, every-BC-combination check of the TMA plane-ring kernel against the oracle and
    against the gather backend (both through hj_step)."""
    from levelsetpy_b200 import _lib as L
    g, d0 = _air_case(lsp, N, pd, tz=tz)
    s = lsp.DubinsVehicleRel(g, 5, 1)
    sd = scheme(lsp, g, s, weno)
    o = osys.DubinsVehicleRel(g, 5, 1)
    osd = orc.OracleSchemeData(grid=g, hamFunc=o.hamiltonian, partialFunc=o.dissipation)
    opts = lsp.odeCFLset(lsp.Bundle(dict(factorCFL=0.8, singleStep="on")))
    y0 = d0.reshape(-1, 1)
    to, yo, _ = orc.ode_cfl3([0.0, 1.0], y0, osd, factor_cfl=0.8, single_step=True, weno=weno)
    out = {}
    for name, be in (("gather", L.BACKEND_GATHER), ("tma", L.BACKEND_TMA)):
        lsp.engine_for_grid(g, weno).set_backend(be)
        t, y, _ = lsp.odeCFL3(lsp.termLaxFriedrichs, [0.0, 1.0], y0, opts, sd)
        assert t == to
        assert_close(y, yo, FIELD_TOL, "%s vs oracle" % name)
        out[name] = y
    lsp.engine_for_grid(g, weno).set_backend(L.BACKEND_AUTO)
    assert_close(out["tma"], out["gather"], 1e-12, "tma vs gather")


def test_step_reductions_match_oracle(lsp):
    """derivMin/derivMax/alphaMax/NaN record of the fused stage kernels (artificial_diss_glf.py:82-88,104)."""
    from levelsetpy_b200 import _lib as L
    from levelsetpy_b200.term import prepare_scheme
    g, d0 = _air_case(lsp, [40, 36, 34], [2])
    s = lsp.DubinsVehicleRel(g, 5, 1)
    sd = scheme(lsp, g, s)
    o = osys.DubinsVehicleRel(g, 5, 1)
    osd = orc.OracleSchemeData(grid=g, hamFunc=o.hamiltonian, partialFunc=o.dissipation)
    _, _, info = orc.term_lax_friedrichs(0.0, d0.reshape(-1, 1), osd, "as_shipped", full=True)
    for be in (L.BACKEND_GATHER, L.BACKEND_TMA):
        eng, ad = prepare_scheme(sd)
        eng.set_backend(be)
        eng.set_system(ad.system_id, ad.block(), list(enumerate(ad.tables(g))))
        eng.upload(d0)
        eng.step(0.0, 1e-3, None, L.COMP_NONE, False, want_reduce=True)
        rec = eng.step_reductions()[0]            # stage 1 sees y0
        assert np.array_equal(rec["alphaMax"], np.array(info["alphaMax"]))
        assert np.allclose(rec["derivMin"], info["derivMin"], rtol=1e-11, atol=1e-12)
        assert np.allclose(rec["derivMax"], info["derivMax"], rtol=1e-11, atol=1e-12)
        assert not rec["nan"]
        eng.set_backend(L.BACKEND_AUTO)


@pytest.mark.parametrize("N,pd", [([70, 40, 50], [2]), ([131, 36, 34], [1, 2]), ([64, 33, 40], [])])
def test_pipelined_host_step_is_bit_identical(lsp, N, pd):
    """hj_ode_cfl3_single with a HOST buffer runs as a chunked H2D / wavefront-of-stages / D2H pipeline (3-D grids with
    N0 >= 64 and an even innermost extent); it must give the bits of the resident three-launch step, for every
    compMethod epilogue, and the oracle's t."""
    from levelsetpy_b200 import _lib as L
    from levelsetpy_b200.term import prepare_scheme
    g, d0 = _air_case(lsp, N, pd)
    s = lsp.DubinsVehicleRel(g, 5, 1)
    sd = scheme(lsp, g, s)
    eng, ad = prepare_scheme(sd)
    eng.set_backend(L.BACKEND_TMA)
    eng.set_system(ad.system_id, ad.block(), list(enumerate(ad.tables(g))))
    o = osys.DubinsVehicleRel(g, 5, 1)
    osd = orc.OracleSchemeData(grid=g, hamFunc=o.hamiltonian, partialFunc=o.dissipation)
    to, yo, _ = orc.ode_cfl3([0.0, 1.0], d0.reshape(-1, 1), osd, factor_cfl=0.8, single_step=True)
    for comp in (L.COMP_NONE, L.COMP_MIN_OVER_TIME):
        y = np.ascontiguousarray(d0.reshape(-1)).copy()
        t1, _, dt = eng.ode_cfl3_single(0.0, 1.0, 0.8, np.finfo(np.float64).max, y, comp)
        eng.upload(d0)
        eng.step(0.0, dt, None, comp, False)
        want = eng.download().reshape(-1)
        assert t1 == to
        assert np.array_equal(y, want), "comp=%d: max diff %.3e" % (comp, np.max(np.abs(y - want)))
        if comp == L.COMP_NONE:
            assert_close(y.reshape(-1, 1), yo, FIELD_TOL, "pipelined step vs oracle")
        # the chunk height is a tuning knob only (hj_set_pipeline_planes): every chunking gives the same bits, down to the
        # smallest one the +-3-plane stencil allows, and with a ragged last chunk
        for planes in (1, 3, 5, 11, 32, N[0]):
            eng.set_pipeline_planes(planes)
            y2 = np.ascontiguousarray(d0.reshape(-1)).copy()
            t2, _, _ = eng.ode_cfl3_single(0.0, 1.0, 0.8, np.finfo(np.float64).max, y2, comp)
            assert t2 == to and np.array_equal(y2, want), "planes=%d comp=%d" % (planes, comp)
        eng.set_pipeline_planes(0)
    eng.set_backend(L.BACKEND_AUTO)
