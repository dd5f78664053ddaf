"""GPU: the slab (stored-halo, HJ_BC_HALO) contexts.  Several slabs of one grid are stepped in lock-step on ONE
device through LocalWorld -- the same SlabSolver code, halo layout, edge-halo fill and reductions the multi-process
NCCL path uses -- and compared with the single-context result and the numpy oracle.  Tolerance: 1e-9 of the value
range after the steps (north_star), identical t sequence."""
import numpy as np
import pytest

from oracle import hj_oracle as orc
from oracle import systems as osys

pytestmark = pytest.mark.gpu


def _grid3(lsp, N, pd):
    gmin, gmax = [-6.0, -10.0, 0.0], [20.0, 10.0, 2 * np.pi]
    for d in pd:
        gmax[d] = gmin[d] + (gmax[d] - gmin[d]) * (1 - 1 / N[d])
    g = lsp.createGrid(np.array(gmin), np.array(gmax), np.array(N), pdDims=pd if pd else None)
    rng = np.random.default_rng(17)
    x = np.meshgrid(*[np.asarray(v).reshape(-1) for v in g.vs], indexing="ij")
    d0 = np.sqrt(x[0] ** 2 + x[1] ** 2) - 5 + 0.4 * np.sin(x[2] + 0.2 * x[0]) + 0.05 * rng.standard_normal(g.shape)
    return g, np.ascontiguousarray(d0)


@pytest.mark.parametrize("world", [1, 2, 3])
@pytest.mark.parametrize("pd", [[2], [0, 2]])
@pytest.mark.parametrize("weno", ["as_shipped", "intended"])
@pytest.mark.parametrize("backend", ["gather", "tma"])
@pytest.mark.parametrize("transport", ["peer", "p2p"])
def test_slabs_match_single_domain(lsp, world, pd, weno, backend, transport):
    from levelsetpy_b200 import _lib as L
    from levelsetpy_b200.slab import LocalWorld
    g, d0 = _grid3(lsp, [26, 37, 34], pd)
    s = lsp.DubinsVehicleRel(g, 5, 1)
    sd = lsp.Bundle(dict(grid=g, hamFunc=s.hamiltonian, partialFunc=s.dissipation, wenoMode=weno))
    o = osys.DubinsVehicleRel(g, 5, 1)
    osd = orc.OracleSchemeData(grid=g, hamFunc=o.hamiltonian, partialFunc=o.dissipation)
    be = L.BACKEND_GATHER if backend == "gather" else L.BACKEND_TMA
    w = LocalWorld(sd, world, backend=be, transport=transport)   # peer: hj_halo_push / hj_halo_wait between contexts
    assert w.peer == (transport == "peer" and world > 1)
    w.poison_halos = True     # ranged mode (TMA, as_shipped): the interior range runs on NaN halos and must not read them
    w.upload(d0)
    t, to, yo = 0.0, 0.0, d0.reshape(-1, 1)
    for _ in range(2):
        t, dt = w.step(t, 1.0, 0.8, comp=L.COMP_MIN_OVER_TIME)
        y_last = yo
        to, yo, _ = orc.ode_cfl3([to, 1.0], yo, osd, factor_cfl=0.8, single_step=True, weno=weno)
        yo = np.minimum(yo, y_last)
        assert t == to
    got, want = w.download(), yo.reshape(g.shape)
    assert all(s.ranged() for s in w.slabs) == (backend == "tma" and weno == "as_shipped")
    assert np.max(np.abs(got - want)) <= 1e-9 * (want.max() - want.min())
    assert np.mean(np.sign(got) == np.sign(want)) >= 0.9999


def test_stage_range_union_equals_stage(lsp):
    """hj_stage_range over disjoint ranges covering the marched dim == hj_stage, bit for bit (incl. in-place stage 3
    and the accumulate-only reduction mode); refused on gather / product / intended contexts."""
    from levelsetpy_b200 import _lib as L
    from levelsetpy_b200.term import prepare_scheme
    g, d0 = _grid3(lsp, [41, 37, 34], [2])
    s = lsp.DubinsVehicleRel(g, 5, 1)
    sd = lsp.Bundle(dict(grid=g, hamFunc=s.hamiltonian, partialFunc=s.dissipation,
                         dissFunc=lsp.artificialDissipationGLF, CoStateCalc=lsp.upwindFirstWENO5a))
    outs, reds = [], []
    # the third partition starts ranges at planes 1 and 2 (ghost planes below the grid are needed by a range that does
    # not start at plane 0) and ends ranges 1 and 2 planes below the top, with single-plane ranges in between: what the
    # pipelined host-buffer step (hj_ode_cfl3_step) launches
    partitions = [None, [(3, 38), (0, 3), (38, 41)], [(2, 3), (0, 1), (1, 2), (3, 20), (20, 39), (40, 41), (39, 40)]]
    for part in partitions:
        eng, ad = prepare_scheme(sd)
        eng.set_backend(L.BACKEND_TMA)
        eng.upload(d0)
        eng.set_system(ad.system_id, ad.block(), list(enumerate(ad.tables(g))))
        for stage in (1, 2, 3):
            if part:
                assert eng.supports_range()
                for k, (a, b) in enumerate(part):
                    eng.stage_range(stage, a, b, 0.0, 1e-3, None, L.COMP_MIN_OVER_TIME, False, 1 if k == 0 else 2)
            else:
                eng.stage(stage, 0.0, 1e-3, None, L.COMP_MIN_OVER_TIME, want_reduce=True)
        reds.append(eng.step_reductions())
        outs.append(eng.download(shape=g.shape))
    for k in (1, 2):
        assert np.array_equal(outs[0], outs[k]), "partition %d" % k
        for a, b in zip(reds[0], reds[k]):
            for key in a:
                assert np.array_equal(np.asarray(a[key]), np.asarray(b[key])), key
    with pytest.raises(Exception):
        eng.stage_range(1, 5, 5, 0.0, 1e-3)
    eng.set_backend(L.BACKEND_GATHER)
    assert not eng.supports_range()


def test_slabs_4d_pair_with_obstacle(lsp):
    """4-D double-integrator pair (config 3 shape, oracle-sized): slabs along dim 0 are a *slow* dim for the TMA
    kernel (halo planes reached through global loads), with the obstacle epilogue fused into stage 3."""
    from levelsetpy_b200 import _lib as L
    from levelsetpy_b200.slab import LocalWorld
    g = lsp.createGrid(-np.ones(4), np.ones(4), np.array([12, 9, 18, 34]))
    x = np.meshgrid(*[np.asarray(v).reshape(-1) for v in g.vs], indexing="ij")
    rng = np.random.default_rng(2)
    d0 = np.sqrt((x[0] - x[2]) ** 2 + (x[1] - x[3]) ** 2) - 0.2 + 0.02 * rng.standard_normal(g.shape)
    obs = 0.3 - np.sqrt(x[0] ** 2 + x[2] ** 2)
    s = lsp.ProductSystem(g, [lsp.DoubleIntegrator(g, 1.0), lsp.DoubleIntegrator(g, 0.6)])
    sd = lsp.Bundle(dict(grid=g, hamFunc=s.hamiltonian, partialFunc=s.dissipation))
    o = osys.ProductSystem([osys.DoubleIntegrator(g, 1.0, dims=(0, 1)), osys.DoubleIntegrator(g, 0.6, dims=(2, 3))])
    osd = orc.OracleSchemeData(grid=g, hamFunc=o.hamiltonian, partialFunc=o.dissipation)
    for be in (L.BACKEND_GATHER, L.BACKEND_TMA):
        w = LocalWorld(sd, 3, backend=be)
        w.upload(d0)
        w.upload(obs, L.FIELD_OBSTACLE)
        t, dt = w.step(0.0, 1.0, 0.8, comp=L.COMP_MIN_OVER_TIME, use_obstacle=True)
        to, yo, _ = orc.ode_cfl3([0.0, 1.0], d0.reshape(-1, 1), osd, factor_cfl=0.8, single_step=True)
        yo = np.maximum(np.minimum(yo, d0.reshape(-1, 1)), -obs.reshape(-1, 1))
        assert t == to
        want = yo.reshape(g.shape)
        assert np.max(np.abs(w.download() - want)) <= 1e-9 * (want.max() - want.min())


def test_slabs_6d_pair_split_path(lsp):
    """6-D relative-Dubins pair on 2 slabs: the dimension-split path with stored halo planes on dim 0, which is a TILED
    dim of pass 2 (one tile covers the thin slab) and is not touched by pass 1."""
    from levelsetpy_b200 import _lib as L
    from levelsetpy_b200.slab import LocalWorld
    N = [10, 9, 8, 7, 9, 12]
    gmin = [-6, -10, 0, -6, -10, 0.]
    gmax = [20, 10, 2 * np.pi * (1 - 1 / N[2]), 20, 10, 2 * np.pi * (1 - 1 / N[5])]
    g = lsp.createGrid(np.array(gmin), np.array(gmax), np.array(N), pdDims=[2, 5])
    x = np.meshgrid(*[np.asarray(v).reshape(-1) for v in g.vs], indexing="ij")
    rng = np.random.default_rng(4)
    d0 = np.minimum(np.sqrt(x[0] ** 2 + x[1] ** 2) - 5, np.sqrt(x[3] ** 2 + x[4] ** 2) - 5) \
        + 0.3 * np.sin(x[2] + x[5]) + 0.02 * rng.standard_normal(g.shape)
    s = lsp.ProductSystem(g, [lsp.DubinsVehicleRel(g, 5, 1), lsp.DubinsVehicleRel(g, 4, 1.2)])
    sd = lsp.Bundle(dict(grid=g, hamFunc=s.hamiltonian, partialFunc=s.dissipation))
    o = osys.ProductSystem([osys.DubinsVehicleRel(g, 5, 1, dims=(0, 1, 2)), osys.DubinsVehicleRel(g, 4, 1.2, dims=(3, 4, 5))])
    osd = orc.OracleSchemeData(grid=g, hamFunc=o.hamiltonian, partialFunc=o.dissipation)
    to, yo = 0.0, d0.reshape(-1, 1)
    for _ in range(2):
        y_last = yo
        to, yo, _ = orc.ode_cfl3([to, 1.0], yo, osd, factor_cfl=0.8, single_step=True)
        yo = np.minimum(yo, y_last)
    want = yo.reshape(g.shape)
    got = {}
    for be, pieces in ((L.BACKEND_GATHER, 1), (L.BACKEND_TMA, 1), (L.BACKEND_TMA, 3), (L.BACKEND_TMA, "auto"),
                       (L.BACKEND_TMA, "fused"), (L.BACKEND_TMA, "hybrid")):
        # pieces > 1: pass 2 in column pieces (hj_stage_pass_cols), each piece of the edge planes pushed to the neighbour
        # as soon as it is computed (hj_halo_push with columns) and awaited piece by piece in the next stage
        # "fused": pass 2 stores its edge planes into the neighbour's halo planes itself (hj_halo_set_fused + hj_halo_signal)
        # "hybrid": that for the upper neighbour only, the planes for the lower one go through the copy engines
        fz = {"fused": True, "hybrid": "hybrid"}.get(pieces, False)
        w = LocalWorld(sd, 2, backend=be, pieces=1 if fz else pieces, fused=fz)
        w.poison_halos = True        # pass 1 (hj_stage_pass) runs on NaN halos: it must not read them
        w.upload(d0)
        t = 0.0
        for _ in range(2):
            t, dt = w.step(t, 1.0, 0.8, comp=L.COMP_MIN_OVER_TIME)
        assert w.slabs[0].two_pass() == (be == L.BACKEND_TMA)   # the two-kernel protocol is what ran on the TMA backend
        assert (w.slabs[0].pieces() is not None) == (be == L.BACKEND_TMA and pieces not in (1, "fused", "hybrid"))
        assert w.slabs[0].fused() == (pieces in ("fused", "hybrid"))
        assert t == to
        got[(be, pieces)] = w.download()
        assert np.max(np.abs(got[(be, pieces)] - want)) <= 1e-9 * (want.max() - want.min())
    assert np.array_equal(got[(L.BACKEND_TMA, 1)], got[(L.BACKEND_TMA, 3)])      # the pieces change nothing, bit for bit
    assert np.array_equal(got[(L.BACKEND_TMA, 1)], got[(L.BACKEND_TMA, "auto")])
    assert np.array_equal(got[(L.BACKEND_TMA, 1)], got[(L.BACKEND_TMA, "fused")])
    assert np.array_equal(got[(L.BACKEND_TMA, 1)], got[(L.BACKEND_TMA, "hybrid")])


def test_stage_pass_equals_stage(lsp):
    """hj_stage == hj_stage_pass(1); hj_stage_pass(2), bit for bit, on a product system; other contexts refuse."""
    from levelsetpy_b200 import _lib as L
    from levelsetpy_b200.term import prepare_scheme
    g = lsp.createGrid(-np.ones(4), np.ones(4), np.array([12, 9, 18, 34]))
    x = np.meshgrid(*[np.asarray(v).reshape(-1) for v in g.vs], indexing="ij")
    d0 = np.sqrt((x[0] - x[2]) ** 2 + (x[1] - x[3]) ** 2) - 0.2 + 0.1 * np.sin(3 * x[1] + x[3])
    s = lsp.ProductSystem(g, [lsp.DoubleIntegrator(g, 1.0), lsp.DoubleIntegrator(g, 0.6)])
    sd = lsp.Bundle(dict(grid=g, hamFunc=s.hamiltonian, partialFunc=s.dissipation,
                         dissFunc=lsp.artificialDissipationGLF, CoStateCalc=lsp.upwindFirstWENO5a))
    outs = []
    for two in (0, 1, 2):
        eng, ad = prepare_scheme(sd)
        eng.set_backend(L.BACKEND_TMA)
        eng.upload(d0)
        eng.set_system(ad.system_id, ad.block(), list(enumerate(ad.tables(g))))
        assert eng.is_split()
        V, q = eng.split_cols()
        assert V == 18 * 34 and q > 0 and q % 2 == 0
        for stage in (1, 2, 3):
            if two == 1:
                eng.stage(stage, 0.0, 1e-3, None, L.COMP_MIN_OVER_TIME, which_pass=1)
                eng.stage(stage, 0.0, 1e-3, None, L.COMP_MIN_OVER_TIME, which_pass=2)
            elif two == 2:          # pass 2 in three uneven column pieces, last piece first (hj_stage_pass_cols)
                eng.stage(stage, 0.0, 1e-3, None, L.COMP_MIN_OVER_TIME, which_pass=1)
                for a, e in ((5 * q, V), (0, 2 * q), (2 * q, 5 * q)):
                    eng.stage_cols(stage, a, e, 0.0, 1e-3, None, L.COMP_MIN_OVER_TIME)
            else:
                eng.stage(stage, 0.0, 1e-3, None, L.COMP_MIN_OVER_TIME)
        outs.append(eng.download(shape=g.shape))
        with pytest.raises(Exception):
            eng.stage_cols(1, 1, V, 0.0, 1e-3)              # not a multiple of the quantum
    assert np.array_equal(outs[0], outs[1])
    assert np.array_equal(outs[0], outs[2])
    g3, d3 = _grid3(lsp, [26, 37, 34], [2])
    s3 = lsp.DubinsVehicleRel(g3, 5, 1)
    eng, ad = prepare_scheme(lsp.Bundle(dict(grid=g3, hamFunc=s3.hamiltonian, partialFunc=s3.dissipation,
                                             dissFunc=lsp.artificialDissipationGLF, CoStateCalc=lsp.upwindFirstWENO5a)))
    eng.upload(d3)
    eng.set_system(ad.system_id, ad.block(), list(enumerate(ad.tables(g3))))
    assert not eng.is_split()
    with pytest.raises(Exception):
        eng.stage(1, 0.0, 1e-3, None, which_pass=1)
