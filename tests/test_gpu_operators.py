"""The single hooks of the reference's operator API as standalone device operators: hamFunc, partialFunc
and dissFunc called on dense arrays -- what the reference's own termLaxFriedrichs does with them
(term_lax_friedrich.py:107-128) -- against the numpy oracle, and composed hook by hook against the fused kernel."""
import numpy as np
import pytest

from oracle import hj_oracle as orc
from oracle import systems as osys

pytestmark = pytest.mark.gpu


def rng_of(a):
    return float(np.max(a) - np.min(a)) or 1.0


def _air3d(lsp, N=(21, 17, 13)):
    g = lsp.createGrid(np.array([-6.0, -10.0, 0.0]), np.array([20.0, 10.0, 2 * np.pi * (1 - 1 / N[2])]), np.array(N), pdDims=2)
    x = np.meshgrid(*[v.reshape(-1) for v in g.vs], indexing="ij")
    d0 = np.sqrt(x[0] ** 2 + x[1] ** 2) - 5 + 0.3 * np.sin(x[2] + 0.2 * x[0]) + 0.05 * np.random.default_rng(1).standard_normal(g.shape)
    return g, np.ascontiguousarray(d0)


def _dint(lsp):
    g = lsp.createGrid(-np.ones(2), np.ones(2), np.array([33, 20]))
    x = np.meshgrid(*[v.reshape(-1) for v in g.vs], indexing="ij")
    return g, np.ascontiguousarray(np.sqrt(x[0] ** 2 + x[1] ** 2) - 0.4 + 0.1 * np.sin(3 * x[0]))


@pytest.mark.parametrize("which", ["dubins", "dint"])
def test_hooks_one_by_one(lsp, which):
    if which == "dubins":
        g, d0 = _air3d(lsp)
        s, o = lsp.DubinsVehicleRel(g, 5, 1), osys.DubinsVehicleRel(g, 5, 1)
    else:
        g, d0 = _dint(lsp)
        s, o = lsp.DoubleIntegrator(g, 0.7), osys.DoubleIntegrator(g, 0.7)
    sd = lsp.Bundle(dict(grid=g, hamFunc=s.hamiltonian, partialFunc=s.dissipation, dissFunc=lsp.artificialDissipationGLF,
                         CoStateCalc=lsp.upwindFirstWENO5a))
    osd = orc.OracleSchemeData(grid=g, hamFunc=o.hamiltonian, partialFunc=o.dissipation)
    oy, osb, full = orc.term_lax_friedrichs(0.0, d0.reshape(-1, 1), osd, "as_shipped", full=True)
    L_, R_ = full["derivL"], full["derivR"]
    C_ = [0.5 * (a + b) for a, b in zip(L_, R_)]
    # hamFunc(t, data, derivC, schemeData)
    ham = s.hamiltonian(0.0, d0, C_, sd)
    want = np.asarray(o.hamiltonian(0.0, d0, C_, osd))
    assert ham.shape == tuple(g.shape)
    assert float(np.max(np.abs(ham - want))) <= 1e-13 * rng_of(want)
    # partialFunc(t, data, derivMin, derivMax, schemeData, dim): state-only alphas, bit for bit
    for d in range(g.dim):
        a = s.dissipation(0.0, d0, full["derivMin"], full["derivMax"], sd, d)
        wa = o.dissipation(0.0, d0, full["derivMin"], full["derivMax"], osd, d)
        wa = np.broadcast_to(np.asarray(wa, dtype=np.float64), g.shape)
        assert np.array_equal(np.asarray(a), wa), d
    # dissFunc(t, data, derivL, derivR, schemeData): the reference's operation order, so bit for bit
    diss, sb = lsp.artificialDissipationGLF(0.0, d0, L_, R_, sd)
    wdiss, wsb, _, _, _ = orc.artificial_dissipation_glf(0.0, d0, L_, R_, osd)
    assert sb == wsb == osb
    assert np.array_equal(diss, np.broadcast_to(wdiss, g.shape))
    # ... and hook by hook they compose to what the fused kernel returns (term_lax_friedrich.py:107-128)
    dl, dr = zip(*[lsp.upwindFirstWENO5a(g, d0, d) for d in range(g.dim)])
    ydot = -(s.hamiltonian(0.0, d0, [0.5 * (a + b) for a, b in zip(dl, dr)], sd)
             - lsp.artificialDissipationGLF(0.0, d0, list(dl), list(dr), sd)[0])
    fused, fsb, _ = lsp.termLaxFriedrichs(0.0, d0.reshape(-1, 1), sd)
    assert fsb == sb
    assert float(np.max(np.abs(ydot.reshape(-1, 1) - fused))) <= 1e-12 * rng_of(fused)
    assert float(np.max(np.abs(fused - oy))) <= 1e-12 * rng_of(oy)
    # LLF as shipped: an array alpha makes `(1 / stepBoundInv).get().item()` raise (diss_local_laxfried.py:126-134)
    with pytest.raises(ValueError, match="size 1"):
        lsp.artificialDissipationLLF(0.0, d0, L_, R_, sd)
    sd.dissFunc = lsp.artificialDissipationLLF
    with pytest.raises(ValueError, match="size 1"):
        lsp.termLaxFriedrichs(0.0, d0.reshape(-1, 1), sd)


def test_llf_equals_glf_for_scalar_alphas(lsp):
    """A Bird's alphas are scalars (bird.py:339-344): LLF runs and is GLF (diss_local_laxfried.py:117-134)."""
    n = 15
    g = lsp.createGrid(np.array([-1.0, -1.0, -np.pi]), np.array([1.0, 1.0, np.pi * (1 - 2 / n)]), np.array([n, n + 2, n]), pdDims=2)
    x = np.meshgrid(*[v.reshape(-1) for v in g.vs], indexing="ij")
    d0 = np.ascontiguousarray(np.sqrt(x[0] ** 2 + x[1] ** 2) - 0.3 + 0.05 * np.sin(2 * x[2] + x[0]))

    def bird(mod):
        mk = lambda k, w, xyw: mod.Bird(g, 1.0, w, init_xyw=np.array(xyw, dtype=np.float64).reshape(3, 1), label=k, neigh_rad=3)
        b = mk(0, 0.9, [0.15, -0.2, 0.4])
        for nb in (mk(1, 1.2, [0.3, 0.1, -0.7]), mk(2, 0.5, [-0.2, 0.25, 1.1])):
            b.update_neighbor(nb)
        return b

    b, ob = bird(lsp), bird(osys)
    y0 = d0.reshape(-1, 1)
    out = {}
    for name in ("artificialDissipationGLF", "artificialDissipationLLF"):
        sd = lsp.Bundle(dict(grid=g, hamFunc=b.hamiltonian, partialFunc=b.dissipation, dissFunc=getattr(lsp, name),
                             CoStateCalc=lsp.upwindFirstWENO5a))
        out[name] = lsp.termLaxFriedrichs(0.0, y0, sd)[:2]
    assert np.array_equal(out["artificialDissipationGLF"][0], out["artificialDissipationLLF"][0])
    assert out["artificialDissipationGLF"][1] == out["artificialDissipationLLF"][1]
    osd = orc.OracleSchemeData(grid=g, hamFunc=ob.hamiltonian, partialFunc=ob.dissipation)
    oy, osb, full = orc.term_lax_friedrichs(0.0, y0, osd, "as_shipped", full=True)
    diss, sb = lsp.artificialDissipationLLF(0.0, d0, full["derivL"], full["derivR"],
                                            lsp.Bundle(dict(grid=g, partialFunc=b.dissipation)))
    wdiss, wsb, _, _, _ = orc.artificial_dissipation_glf(0.0, d0, full["derivL"], full["derivR"], osd)
    assert sb == wsb == osb
    assert np.array_equal(diss, np.broadcast_to(wdiss, g.shape))
    assert b.dissipation(0.0, d0, None, None, None, 1) == ob.dissipation(0.0, d0, None, None, osd, 1)


def test_generate_all_candidates_bit_exact(lsp):
    """upwindFirstWENO5a / upwindFirstENO3a(..., generateAll=True) = the six candidates of upwindFirstENO3aHelper
    (upwind_first_weno5a.py:73-75, ENO3aHelper.py:116-189): un-fused arithmetic in the reference's order, so bit for bit."""
    g, d0 = _air3d(lsp, (21, 17, 13))
    for d in range(3):
        for fn in (lsp.upwindFirstWENO5a, lsp.upwindFirstENO3a):
            dL, dR = fn(g, d0, d, True)
            wL, wR, _, _ = orc.eno3a_helper(g, d0, d)
            assert len(dL) == 3 and len(dR) == 3
            for k in range(3):
                assert np.array_equal(np.asarray(dL[k]), wL[k]), (d, k)
                assert np.array_equal(np.asarray(dR[k]), wR[k]), (d, k)
    with pytest.raises(ValueError):
        lsp.upwindFirstWENO5a(g, d0, 4, True)
