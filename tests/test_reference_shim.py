"""CPU, build container only: the numpy oracle against the LITERAL reference (imported from /root/reference through
oracle/ref_shim.py) on fresh seeded inputs that are NOT among the committed fixtures.  Skipped where the reference
tree is absent (the GPU box)."""
import numpy as np
import pytest

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ref():
    ref_shim.install()
    import LevelSetPy.BoundaryCondition as BC
    import LevelSetPy.DynamicalSystems as DS
    import LevelSetPy.ExplicitIntegration as EI
    import LevelSetPy.Grids as G
    import LevelSetPy.SpatialDerivative as SD
    import LevelSetPy.Utilities as U
    return dict(BC=BC, DS=DS, EI=EI, G=G, SD=SD, U=U)


def col(x, dt=np.float64):
    return np.asarray(x, dtype=dt).reshape(-1, 1)


def test_oracle_matches_reference_on_fresh_inputs(ref):
    from oracle import hj_oracle as orc
    from oracle import systems as osys
    N = [14, 19, 11]
    g = ref["G"].createGrid(col([-4, -7, 0]), col([9, 6, 2 * np.pi * (1 - 1 / N[2])]), col(N, np.int64), pdDims=2)
    rng = np.random.default_rng(2024)
    data = np.ascontiguousarray(np.sqrt(g.xs[0] ** 2 + g.xs[1] ** 2) - 3 + 0.3 * np.cos(g.xs[2]) + 0.1 * rng.standard_normal(g.shape))
    for d in range(3):
        L, R = ref["SD"].upwindFirstWENO5a(g, data, d)
        oL, oR = orc.upwind_first_weno5a(g, data, d, "as_shipped")
        assert np.array_equal(np.asarray(L), oL) and np.array_equal(np.asarray(R), oR)
    B = ref["U"].Bundle
    rs = ref["DS"].DubinsVehicleRel(g, 3, 1.5)
    sd = B(dict(grid=g, hamFunc=rs.hamiltonian, partialFunc=rs.dissipation,
                dissFunc=ref["EI"].artificialDissipationGLF, CoStateCalc=ref["SD"].upwindFirstWENO5a))
    os_ = osys.DubinsVehicleRel(g, 3, 1.5)
    osd = orc.OracleSchemeData(grid=g, hamFunc=os_.hamiltonian, partialFunc=os_.dissipation)
    opts = ref["EI"].odeCFLset(B({"factorCFL": 0.8, "singleStep": "on"}))
    y0 = data.reshape(-1, 1)
    t, y, _ = ref["EI"].odeCFL3(ref["EI"].termLaxFriedrichs, [0.0, 1.0], y0, opts, sd)
    to, yo, _ = orc.ode_cfl3([0.0, 1.0], y0, osd, factor_cfl=0.8, single_step=True)
    assert t == to and np.array_equal(np.asarray(y), yo)


def test_reference_weno5a_is_fixed_weight_as_shipped(ref):
    """SURVEY.md fact 4, checked on the literal reference: max |WENO5a - (.1 d0 + .6 d1 + .3 d2)| is at ulp level."""
    g = ref["G"].createGrid(col([0, 0]), col([1, 2]), col([25, 31], np.int64))
    rng = np.random.default_rng(5)
    data = rng.standard_normal(g.shape)
    for d in range(2):
        L, R = ref["SD"].upwindFirstWENO5a(g, data, d)
        dL, dR, _ = ref["SD"].upwindFirstENO3aHelper(g, data, d, False, False)
        wL = 0.1 * np.asarray(dL[0]) + 0.6 * np.asarray(dL[1]) + 0.3 * np.asarray(dL[2])
        assert np.max(np.abs(np.asarray(L) - wL)) <= 16 * np.finfo(float).eps * np.max(np.abs(wL))


def test_oracle_eno_matches_reference_on_fresh_inputs(ref):
    """upwindFirstENO2 / upwindFirstENO3a, incl. towardZero ghost cells and the fused RHS."""
    from oracle import hj_oracle as orc
    from oracle import systems as osys
    N = [16, 12, 10]
    g = ref["G"].createGrid(col([-4, -7, 0]), col([9, 6, 2 * np.pi * (1 - 1 / N[2])]), col(N, np.int64), pdDims=2)
    rng = np.random.default_rng(77)
    data = np.ascontiguousarray(np.sqrt(g.xs[0] ** 2 + g.xs[1] ** 2) - 3 + 0.5 * np.sin(2 * g.xs[2]) + 0.2 * rng.standard_normal(g.shape))
    pairs = ((ref["SD"].upwindFirstENO2, orc.upwind_first_eno2, "eno2"), (ref["SD"].upwindFirstENO3a, orc.upwind_first_eno3a, "eno3a"))
    for rfn, ofn, tag in pairs:
        for d in range(3):
            L, R = rfn(g, data, d)
            oL, oR = ofn(g, data, d)
            assert np.array_equal(np.asarray(L), oL) and np.array_equal(np.asarray(R), oR), (tag, d)
        B = ref["U"].Bundle
        rs = ref["DS"].DubinsVehicleRel(g, 3, 1.5)
        sd = B(dict(grid=g, hamFunc=rs.hamiltonian, partialFunc=rs.dissipation,
                    dissFunc=ref["EI"].artificialDissipationGLF, CoStateCalc=rfn))
        os_ = osys.DubinsVehicleRel(g, 3, 1.5)
        osd = orc.OracleSchemeData(grid=g, hamFunc=os_.hamiltonian, partialFunc=os_.dissipation)
        ydot, sb, _ = ref["EI"].termLaxFriedrichs(0.0, data.reshape(-1, 1), sd)
        oydot, osb = orc.term_lax_friedrichs(0.0, data.reshape(-1, 1), osd, tag)
        assert sb == osb and np.array_equal(np.asarray(ydot), oydot)


def test_reference_objects_resolve_to_device_functors(ref):
    """INTEGRATION.md section 1 promises that the reference's OWN grid and DynamicalSystems objects are recognised
    (duck-typed by class name + attributes): hand them to functors.resolve / engine.grid_signature and compare the
    functor ids, parameter blocks and trig tables with what this package's own objects (and the oracle's) give."""
    import levelsetpy_b200 as lsp
    from levelsetpy_b200 import _lib as L
    from levelsetpy_b200.engine import grid_signature
    from levelsetpy_b200.functors import resolve
    N = [14, 19, 11]
    gmin, gmax = col([-4, -7, 0]), col([9, 6, 2 * np.pi * (1 - 1 / N[2])])
    rg = ref["G"].createGrid(gmin, gmax, col(N, np.int64), pdDims=2)
    og = lsp.createGrid(gmin, gmax, np.array(N), pdDims=2)
    rs, os_ = grid_signature(rg), grid_signature(og)
    assert rs[0] == os_[0] and rs[1] == os_[1] and rs[3] == os_[3] and rs[4] == os_[4]
    assert rs[2] == os_[2], "grid.dx must be bit-identical"
    for a, b in zip(rs[5], os_[5]):
        assert np.array_equal(a, b), "grid.vs must be bit-identical"
    # relative Dubins
    rsys, osys_ = ref["DS"].DubinsVehicleRel(rg, 3, 1.5), lsp.DubinsVehicleRel(og, 3, 1.5)
    ra, oa = resolve(rsys.hamiltonian, rsys.dissipation, rg), resolve(osys_.hamiltonian, osys_.dissipation, og)
    assert ra.system_id == oa.system_id == L.SYS_DUBINS_REL
    assert np.array_equal(ra.block(), oa.block())
    for a, b in zip(ra.tables(rg), oa.tables(og)):
        assert np.array_equal(a, b)
    assert np.array_equal(ra.tables(rg)[0], np.cos(np.asarray(rg.vs[2]).reshape(-1)))   # dubins_relative.py:81
    # double integrator
    rg2 = ref["G"].createGrid(col([-1, -1]), col([1, 1]), col([21, 17], np.int64))
    rdi = ref["DS"].DoubleIntegrator(rg2, 0.7)
    ad = resolve(rdi.hamiltonian, rdi.dissipation, rg2)
    assert ad.system_id == L.SYS_DOUBLE_INT and np.array_equal(ad.block(), np.array([0.7]))
    # flock of 4 birds: the adapter drives the REFERENCE's own _housekeeping; blocks must equal the ones built from
    # this package's Flock over three consecutive RHS evaluations (flock.py:213 mutates the headings every call)
    n = 9
    rg3 = ref["G"].createGrid(col([-1, -1, -np.pi]), col([1, 1, np.pi * (1 - 2 / n)]), col([n, n, n], np.int64), pdDims=2)
    og3 = lsp.createGrid(col([-1, -1, -np.pi]), col([1, 1, np.pi * (1 - 2 / n)]), np.array([n, n, n]), pdDims=2)
    wb = [0.8, 1.0, 1.3, 0.6]
    xyw = [[0.1 * j - 0.05, 0.2 * j - 0.3, 0.3 * j + 0.1] for j in range(4)]
    rf = ref["DS"].Flock(rg3, [ref["DS"].Bird(rg3, 1.0, wb[j], init_xyw=np.array([xyw[j]]).T.copy(), label=j, neigh_rad=3)
                               for j in range(4)], label=1)
    of = lsp.Flock(og3, [lsp.Bird(og3, 1.0, wb[j], init_xyw=np.array([xyw[j]]).T.copy(), label=j, neigh_rad=3)
                         for j in range(4)])
    ra, oa = resolve(rf.hamiltonian, rf.dissipation, rg3), resolve(of.hamiltonian, of.dissipation, og3)
    assert ra.system_id == oa.system_id == L.SYS_FLOCK and ra.time_varying
    for _ in range(3):
        br, bo = ra.block(), oa.block()
        assert np.array_equal(br, bo)
        assert ra.alphas(br) == oa.alphas(bo)
    # a reference Bird on its own, both method pairs
    rb = rf.vehicles[1]
    for pair in (("hamiltonian", "dissipation"), ("hamiltonian_abs", "dissipation_abs")):
        ad = resolve(getattr(rb, pair[0]), getattr(rb, pair[1]), rg3)
        ab = resolve(getattr(of.vehicles[1], pair[0]), getattr(of.vehicles[1], pair[1]), og3)
        assert np.array_equal(ad.block(), ab.block())
    # anything unregistered is refused, never evaluated on the CPU
    with pytest.raises(NotImplementedError):
        resolve(lambda *a: 0, lambda *a: 0, rg)


def test_reference_driver_zero_is_set_and_full_horizon(ref):
    """The literal HJIPDE_solve: (1) compMethod 'zero' integrates exactly like 'set' (its time loop hard-codes
    termLaxFriedrichs, hji_solver.py:542) and 'minWithZero' ends in error() (:599) -- the as-shipped behaviour
    levelsetpy_b200.HJIPDE_solve reproduces by default; (2) over a full horizon (t in [0, 1], minVOverTime) the
    oracle's restatement stays bit-identical to it, dt sequence included."""
    from oracle import hj_oracle as orc
    from oracle import systems as osys
    import LevelSetPy.ValueFuncs as VF
    B = ref["U"].Bundle
    N = [15, 13, 11]
    g = ref["G"].createGrid(col([-6, -10, 0]), col([20, 10, 2 * np.pi * (1 - 1 / N[2])]), col(N, np.int64), pdDims=2)
    d0 = np.ascontiguousarray(np.sqrt(g.xs[0] ** 2 + g.xs[1] ** 2) - 5.0)

    def sd():
        s = ref["DS"].DubinsVehicleRel(g, 5, 1)
        return B(dict(grid=g, hamFunc=s.hamiltonian, partialFunc=s.dissipation,
                      dissFunc=ref["EI"].artificialDissipationGLF, CoStateCalc=ref["SD"].upwindFirstWENO5a))
    o = osys.DubinsVehicleRel(g, 5, 1)
    osd = orc.OracleSchemeData(grid=g, hamFunc=o.hamiltonian, partialFunc=o.dissipation)
    tau = np.array([0.0, 0.1, 0.2])
    q = lambda: B(dict(quiet=True, keepLast=True))
    a = np.asarray(VF.HJIPDE_solve(d0.copy(), tau, sd(), "zero", q())[0])
    b = np.asarray(VF.HJIPDE_solve(d0.copy(), tau, sd(), "set", q())[0])
    assert np.array_equal(a, b), "as shipped, 'zero' == 'set'"
    assert np.array_equal(a, orc.hji_solve(d0, tau, osd, "zero")[0])
    with pytest.raises(Exception):
        VF.HJIPDE_solve(d0.copy(), tau, sd(), "minWithZero", q())
    tau = np.linspace(0.0, 1.0, 4)
    full = np.asarray(VF.HJIPDE_solve(d0.copy(), tau, sd(), "minVOverTime", q())[0])
    want, dts, _ = orc.hji_solve(d0, tau, osd, "minVOverTime")
    assert len(dts) >= 8
    assert np.array_equal(full, want), "oracle == literal reference after the full horizon (%d steps)" % len(dts)


def test_reference_llf_as_shipped(ref):
    """artificialDissipationLLF as shipped (diss_local_laxfried.py:126-134): `stepBoundInv` is summed un-maximised, so with
    an ARRAY alpha `(1 / stepBoundInv).get().item()` raises ValueError -- the behaviour this package reproduces
    (levelsetpy_b200/dissipation.py, term.py) -- while GLF on the same inputs runs; and the reference's host-side
    validation of this package's schemeData for LLF gives the same verdict without touching a GPU."""
    from LevelSetPy.ExplicitIntegration.Dissipation import artificialDissipationLLF
    import levelsetpy_b200 as lsp
    N = [13, 11, 9]
    g = ref["G"].createGrid(col([-6, -10, 0]), col([20, 10, 2 * np.pi * (1 - 1 / N[2])]), col(N, np.int64), pdDims=2)
    d0 = np.ascontiguousarray(np.sqrt(g.xs[0] ** 2 + g.xs[1] ** 2) - 5 + 0.2 * np.sin(g.xs[2]))
    B = ref["U"].Bundle
    s = ref["DS"].DubinsVehicleRel(g, 5, 1)
    sd = B(dict(grid=g, hamFunc=s.hamiltonian, partialFunc=s.dissipation, dissFunc=ref["EI"].artificialDissipationGLF,
                CoStateCalc=ref["SD"].upwindFirstWENO5a))
    _, sb, _ = ref["EI"].termLaxFriedrichs(0.0, d0.reshape(-1, 1), sd)
    assert sb > 0
    sd.dissFunc = artificialDissipationLLF
    with pytest.raises(ValueError, match="size 1"):
        ref["EI"].termLaxFriedrichs(0.0, d0.reshape(-1, 1), sd)
    # this package, same bundle shape: the LLF token with an array-alpha system is refused with the same error, before
    # any device call (prepare_scheme resolves the functor on the host)
    from levelsetpy_b200.term import prepare_scheme
    mine = lsp.DubinsVehicleRel(g, 5, 1)
    with pytest.raises(ValueError, match="size 1"):
        prepare_scheme(lsp.Bundle(dict(grid=g, hamFunc=mine.hamiltonian, partialFunc=mine.dissipation,
                                       dissFunc=lsp.artificialDissipationLLF, CoStateCalc=lsp.upwindFirstWENO5a)))
