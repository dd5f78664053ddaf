"""CPU: the host side of the batched flock path.  FlockBatchPlanner evaluates the flock
bookkeeping (Flock._housekeeping, flock.py:147-188) and the per-flock parameter blocks for the whole batch at once; it
must reproduce the per-flock pass (batch_step_plan) bit for bit -- blocks, dt, t and the birds' mutated headings --
and must step aside (build -> None) for batches it does not cover."""
import numpy as np


def _flocks(lsp, nb, nbirds=4, neigh_rad=3, n=9):
    sds = []
    for j in range(nb):
        sh = 0.2 * j
        g = lsp.createGrid(np.array([-1 + sh, -1 + sh, -np.pi]), np.array([1 + sh, 1 + sh, np.pi * (1 - 2 / n)]),
                           np.array([n, n, n]), pdDims=2, low_mem=True)
        birds = [lsp.Bird(g, 1.0 + 0.01 * k, 0.6 + 0.1 * k + 0.03 * j,
                          init_xyw=np.array([[0.1 * j + 0.01 * k], [0.2 * j - 0.02 * k], [0.3 * j + 0.1 * k - 1.0]]),
                          label=k, neigh_rad=neigh_rad) for k in range(nbirds)]
        f = lsp.Flock(g, birds)
        sds.append(lsp.Bundle(dict(grid=g, hamFunc=f.hamiltonian, partialFunc=f.dissipation)))
    return sds


def _adapters(lsp, sds):
    from levelsetpy_b200.functors import resolve
    return [resolve(sd.hamFunc, sd.partialFunc, sd.grid) for sd in sds], [np.asarray(sd.grid.dx).reshape(-1) for sd in sds]


def test_planner_is_bit_identical_to_per_flock_pass(lsp):
    from levelsetpy_b200.batch import FlockBatchPlanner, batch_step_plan
    for nbirds, rad in ((4, 3), (5, 2), (3, 3), (6, 4)):
        nb = 7
        sa, sb = _flocks(lsp, nb, nbirds, rad), _flocks(lsp, nb, nbirds, rad)
        aa, dxs = _adapters(lsp, sa)
        ab, _ = _adapters(lsp, sb)
        pl = FlockBatchPlanner.build(ab)
        assert pl is not None
        ta = tb = np.zeros(nb)
        te = np.linspace(0.01, 1.0, nb)          # some grids are limited by t_end - t, others by the CFL bound
        for _ in range(3):
            pa, da, ta = batch_step_plan(aa, dxs, ta, te, 0.8)
            pb, db, tb = pl.plan(dxs, tb, te, 0.8)
            assert pa.shape == pb.shape and np.array_equal(pa, pb)
            assert np.array_equal(da, db) and np.array_equal(ta, tb)
            wa = np.array([b.w_e for sd in sa for b in sd.hamFunc.__self__.vehicles], dtype=np.float64)
            wb = np.array([b.w_e for sd in sb for b in sd.hamFunc.__self__.vehicles], dtype=np.float64)
            assert np.array_equal(wa, wb)


def test_planner_steps_aside_for_non_uniform_batches(lsp):
    from levelsetpy_b200.batch import FlockBatchPlanner
    mixed = _flocks(lsp, 2, 4, 3) + _flocks(lsp, 1, 5, 3)
    assert FlockBatchPlanner.build(_adapters(lsp, mixed)[0]) is None           # different flock sizes
    radii = _flocks(lsp, 2, 4, 3) + _flocks(lsp, 1, 4, 2)
    assert FlockBatchPlanner.build(_adapters(lsp, radii)[0]) is None           # different neighbour lists
    pair = _flocks(lsp, 1, 4, 3)
    ads = _adapters(lsp, pair)[0]
    outsider = lsp.Bird(pair[0].grid, 1.0, 1.0, init_xyw=np.array([[0.0], [0.0], [0.0]]), label=1, neigh_rad=3)
    pair[0].hamFunc.__self__.vehicles[0].update_neighbor(outsider)
    assert FlockBatchPlanner.build(ads) is None                                # a neighbour outside the flock
    g = pair[0].grid
    s = lsp.DubinsVehicleRel(g, 5, 1)
    from levelsetpy_b200.functors import resolve
    assert FlockBatchPlanner.build([resolve(s.hamiltonian, s.dissipation, g)]) is None
