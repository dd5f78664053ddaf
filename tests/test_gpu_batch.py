"""Batched small grids: BatchSolver (hj_create_batch / hj_step_batch, one launch per RK stage
for the whole batch) against the numpy oracle run grid by grid, and against this package's own per-grid odeCFL3.
Tolerances as in test_gpu_parity.py: identical t / dt per grid, fields within 1e-9 of the value range."""
import numpy as np
import pytest

from oracle import hj_oracle as orc
from oracle import systems as osys

pytestmark = pytest.mark.gpu


def _flock_case(lsp, j, N):
    """flockGrid-style box shifted by 0.2 j, a 4-bird flock whose initial states depend on j."""
    sh = 0.2 * j
    g = lsp.createGrid(np.array([-1 + sh, -1 + sh, -np.pi]), np.array([1 + sh, 1 + sh, np.pi * (1 - 2 / N[2])]),
                       np.array(N), pdDims=2)
    wb = [0.8, 1.0, 1.3 + 0.05 * j, 0.6]
    xyw = [[0.1 * k - 0.05 + 0.01 * j, 0.2 * k - 0.3, 0.3 * k + 0.1 * (j + 1)] for k in range(4)]
    birds = [lsp.Bird(g, 1.0, wb[k], init_xyw=np.array([xyw[k]]).T.copy(), label=k, neigh_rad=3) for k in range(4)]
    obirds = [osys.Bird(g, 1.0, wb[k], init_xyw=np.array(xyw[k]), label=k, neigh_rad=3) for k in range(4)]
    rng = np.random.default_rng(100 + j)
    d0 = lsp.shapeCylinder(g, 2, np.array([[sh], [sh], [0.0]]), 0.3) + 0.02 * rng.standard_normal(g.shape)
    return g, lsp.Flock(g, birds), osys.Flock(g, obirds), np.ascontiguousarray(d0)


@pytest.mark.parametrize("nb,N", [(5, [21, 19, 24]), (2, [36, 40, 31]), (9, [15, 15, 15])])
def test_batch_matches_per_grid_oracle(lsp, nb, N):
    from levelsetpy_b200 import _lib as L
    cases = [_flock_case(lsp, j, N) for j in range(nb)]
    sds = [lsp.Bundle(dict(grid=g, hamFunc=f.hamiltonian, partialFunc=f.dissipation,
                           dissFunc=lsp.artificialDissipationGLF, CoStateCalc=lsp.upwindFirstWENO5a))
           for g, f, _, _ in cases]
    bs = lsp.BatchSolver(sds)
    bs.upload([c[3] for c in cases])
    nsteps = 3
    ts, dts = [], []
    for _ in range(nsteps):
        t, dt = bs.step(1.0, 0.8, comp=L.COMP_MIN_OVER_TIME)
        ts.append(t)
        dts.append(dt)
    got = bs.download()
    assert got.shape == (nb,) + tuple(N)
    for j, (g, _, of, d0) in enumerate(cases):
        osd = orc.OracleSchemeData(grid=g, hamFunc=of.hamiltonian, partialFunc=of.dissipation)
        to, yo = 0.0, d0.reshape(-1, 1)
        for k in range(nsteps):
            y_last = yo
            to, yo, _ = orc.ode_cfl3([to, 1.0], yo, osd, factor_cfl=0.8, single_step=True)
            yo = np.minimum(yo, y_last)
            assert ts[k][j] == to, "grid %d step %d: t %r vs oracle %r" % (j, k, ts[k][j], to)
        want = yo.reshape(g.shape)
        err = float(np.max(np.abs(got[j] - want)))
        assert err <= 1e-9 * float(want.max() - want.min()), "grid %d: %.3e" % (j, err)
        assert np.mean((got[j] > 0) == (want > 0)) >= 0.9999
    # the grids really had different time steps (per-grid dt is exercised)
    assert len({float(x) for x in dts[0]}) > 1


def test_batch_matches_own_per_grid_path(lsp):
    """Same batch through this package's per-grid odeCFL3 (one context per grid): bit-identical fields."""
    nb, N = 3, [24, 18, 20]
    cases = [_flock_case(lsp, j, N) for j in range(nb)]
    mk = lambda g, f: lsp.Bundle(dict(grid=g, hamFunc=f.hamiltonian, partialFunc=f.dissipation,
                                      dissFunc=lsp.artificialDissipationGLF, CoStateCalc=lsp.upwindFirstWENO5a))
    bs = lsp.BatchSolver([mk(g, f) for g, f, _, _ in cases])
    bs.upload([c[3] for c in cases])
    tb, _ = bs.step(1.0, 0.8)
    got = bs.download()
    opts = lsp.odeCFLset(lsp.Bundle(dict(factorCFL=0.8, singleStep="on")))
    for j in range(nb):
        g, f, _, d0 = _flock_case(lsp, j, N)               # fresh flock: the batch step mutated the first one's headings
        t, y, _ = lsp.odeCFL3(lsp.termLaxFriedrichs, [0.0, 1.0], d0.reshape(-1, 1), opts, mk(g, f))
        assert t == tb[j]
        assert np.array_equal(y.reshape(g.shape), got[j])


def test_batch_rejects_what_it_cannot_run(lsp):
    g, f, _, _ = _flock_case(lsp, 0, [15, 15, 15])
    g2, f2, _, _ = _flock_case(lsp, 1, [15, 16, 15])
    mk = lambda g, f, **kw: lsp.Bundle(dict(grid=g, hamFunc=f.hamiltonian, partialFunc=f.dissipation, **kw))
    with pytest.raises(ValueError):
        lsp.BatchSolver([mk(g, f), mk(g2, f2)])            # different shapes
    with pytest.raises(NotImplementedError):
        lsp.BatchSolver([mk(g, f, wenoMode="intended")])
    s = lsp.DubinsVehicleRel(g, 5, 1)
    with pytest.raises(NotImplementedError):
        lsp.BatchSolver([lsp.Bundle(dict(grid=g, hamFunc=s.hamiltonian, partialFunc=s.dissipation))])
