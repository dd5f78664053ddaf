"""CPU: the N>1 host logic -- slab partitioning, halo send/recv pairing (incl. the 2-rank periodic ring), edge-halo
extrapolation, alpha max-allreduce and the identical-dt rule -- exercised with the REAL SlabSolver over
``torch.distributed`` (gloo, world_size 2 and 3) and over the in-process LocalWorld, with the numpy oracle standing
in for the per-slab CUDA context (tests/slab_oracle_engine.py).  Reference result: the single-domain oracle."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case(periodic0):
    import levelsetpy_b200 as lsp
    N = [14, 9, 8]
    pd = [0, 2] if periodic0 else [2]
    gmin, gmax = [-6.0, -10.0, 0.0], [20.0, 10.0, 2 * np.pi * (1 - 1 / N[2])]
    if periodic0:
        gmax[0] = gmin[0] + (gmax[0] - gmin[0]) * (1 - 1 / N[0])
    g = lsp.createGrid(np.array(gmin), np.array(gmax), np.array(N), pdDims=pd)
    rng = np.random.default_rng(3)
    x = np.meshgrid(*[np.asarray(v).reshape(-1) for v in g.vs], indexing="ij")
    d0 = np.sqrt(x[0] ** 2 + x[1] ** 2) - 5 + 0.3 * np.sin(x[2]) + 0.05 * rng.standard_normal(g.shape)
    s = lsp.DubinsVehicleRel(g, 5, 1)
    sd = lsp.Bundle(dict(grid=g, hamFunc=s.hamiltonian, partialFunc=s.dissipation))
    return lsp, g, np.ascontiguousarray(d0), sd


def _oracle(g, d0, nsteps, comp):
    from oracle import hj_oracle as orc
    from oracle import systems as osys
    o = osys.DubinsVehicleRel(g, 5, 1)
    osd = orc.OracleSchemeData(grid=g, hamFunc=o.hamiltonian, partialFunc=o.dissipation)
    t, y, ts = 0.0, d0.reshape(-1, 1), []
    for _ in range(nsteps):
        y_last = y
        t, y, _ = orc.ode_cfl3([t, 1.0], y, osd, factor_cfl=0.8, single_step=True)
        if comp:
            y = np.minimum(y, y_last)
        ts.append(t)
    return ts, y.reshape(g.shape)


def _worker(rank, world, port, periodic0, q, two_pass=False):
    import torch.distributed as dist
    from slab_oracle_engine import OracleSlabEngine, TwoPassOracleSlabEngine, RangedOracleSlabEngine
    from levelsetpy_b200.slab import SlabSolver
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        lsp, g, d0, sd = _case(periodic0)
        factory = {False: OracleSlabEngine, True: TwoPassOracleSlabEngine, "ranged": RangedOracleSlabEngine}[two_pass]
        sol = SlabSolver(sd, device=0, engine_factory=factory)
        if two_pass == "ranged":
            # the interior range of every stage must not read the halo planes: poison them before each step's first use
            orig_post = sol.comm.post

            def poisoned_post(sends, recvs):
                for t_, _, _ in recvs:
                    t_.fill_(float("nan"))
                return orig_post(sends, recvs)
            sol.comm.post = poisoned_post
        sol.upload(d0[sol.lo:sol.hi])
        t, ts = 0.0, []
        for _ in range(2):
            t, dt = sol.step(t, 1.0, 0.8, comp=1)
            ts.append(t)
        if two_pass == "ranged":
            n0 = sol.n0
            assert sol.ranged() and not sol.overlapped()
            assert sol.eng.log == [(s, a, b) for _ in range(2) for s in (1, 2, 3)
                                   for a, b in ((3, n0 - 3), (0, 3), (n0 - 3, n0))]
        elif two_pass:      # the exchange of every stage was posted before pass 1 and awaited before pass 2
            assert sol.overlapped()
            assert sol.eng.log == [(p, s) for _ in range(2) for s in (1, 2, 3) for p in ("pass1", "pass2")]
        q.put((rank, sol.lo, sol.hi, ts, sol.download()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,periodic0,two_pass", [(2, False, False), (2, True, False), (3, False, False),
                                                      (2, False, True), (2, True, True),
                                                      (2, False, "ranged"), (2, True, "ranged")])
def test_slab_solver_over_gloo_matches_single_domain_oracle(world, periodic0, two_pass):
    """two_pass: the overlapped protocol of product systems (halo exchange posted, pass 1, wait, pass 2);
    "ranged": the protocol of whole 3-D systems (exchange posted, interior planes, wait, the two 3-plane edge ranges),
    with the receiving halo planes NaN-poisoned when the exchange is posted."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, periodic0, q, two_pass)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    _, g, d0, _ = _case(periodic0)
    ts, want = _oracle(g, d0, 2, True)
    got = np.concatenate([r[4] for r in res], axis=0)
    assert [r[1:3] for r in res][0][0] == 0 and res[-1][2] == g.shape[0]
    for r in res:
        assert r[3] == ts, "every rank must produce the single-domain dt / t sequence"
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) <= 1e-12 * (want.max() - want.min())


@pytest.mark.parametrize("world,periodic0", [(1, True), (1, False), (4, True), (3, False)])
def test_local_world_matches_single_domain_oracle(world, periodic0):
    from slab_oracle_engine import OracleSlabEngine
    from levelsetpy_b200.slab import LocalWorld
    lsp, g, d0, sd = _case(periodic0)
    w = LocalWorld(sd, world, engine_factory=OracleSlabEngine)
    w.upload(d0)
    t, ts = 0.0, []
    for _ in range(2):
        t, dt = w.step(t, 1.0, 0.8, comp=1)
        ts.append(t)
    want_ts, want = _oracle(g, d0, 2, True)
    assert ts == want_ts
    assert np.max(np.abs(w.download() - want)) <= 1e-12 * (want.max() - want.min())


def test_partition():
    from levelsetpy_b200.slab import partition
    assert partition(41, 8) == [(0, 6), (6, 11), (11, 16), (16, 21), (21, 26), (26, 31), (31, 36), (36, 41)]
    assert partition(161, 2) == [(0, 81), (81, 161)]
    assert partition(12, 4) == [(0, 3), (3, 6), (6, 9), (9, 12)]
    with pytest.raises(ValueError):
        partition(16, 8)


@pytest.mark.parametrize("mode", ["whole", "pieces", "fused", "hybrid"])
@pytest.mark.parametrize("world,periodic0", [(2, False), (3, True), (4, False)])
def test_peer_halo_protocols_on_cpu(world, periodic0, mode):
    """The host side of the peer-memory halo protocols (levelsetpy_b200/slab.py) with a CPU stand-in for
    the per-slab context (tests/slab_oracle_engine.py::PeerOracleSlabEngine): whole-plane pushes, pass 2 in column
    pieces with early pushes, halo planes stored by pass 2 itself (both sides / hybrid).  A wait that precedes the push
    it depends on trips an assertion in the stand-in; the result must equal the single-domain oracle."""
    from slab_oracle_engine import PeerOracleSlabEngine
    from levelsetpy_b200.slab import LocalWorld
    lsp, g, d0, sd = _case(periodic0)
    kw = {"whole": dict(pieces=1, fused=False), "pieces": dict(pieces=3, fused=False),
          "fused": dict(pieces=1, fused=True), "hybrid": dict(pieces=1, fused="hybrid")}[mode]
    w = LocalWorld(sd, world, engine_factory=PeerOracleSlabEngine, **kw)
    assert w.peer
    w.upload(d0)
    t, ts = 0.0, []
    for _ in range(2):
        t, dt = w.step(t, 1.0, 0.8, comp=1)
        ts.append(t)
    want_ts, want = _oracle(g, d0, 2, True)
    assert ts == want_ts
    assert np.max(np.abs(w.download() - want)) <= 1e-12 * (want.max() - want.min())
    s0 = w.slabs[0]
    kinds = [c[0] for c in s0.eng.calls]
    assert s0.fused() == (mode in ("fused", "hybrid")) and (s0.pieces() is not None) == (mode == "pieces")
    if mode == "fused":          # one priming push, then only signals: no copy is queued in the steady state
        assert kinds.count("push") == 1 and kinds.count("signal") == 6
    if mode == "hybrid":         # every stage: signal towards the upper neighbour, copy-engine push towards the lower one
        assert [c[3] for c in s0.eng.calls if c[0] == "push"][1:] == [1] * 6
        assert [c[2] for c in s0.eng.calls if c[0] == "signal"] == [2] * 6
    if mode == "pieces":
        assert kinds.count("push") == 3 * (1 + 6) and kinds.count("wait") == 3 * 6
    # a second upload drops what was pushed ahead for the old state and the march still matches
    w.upload(d0)
    t2, _ = w.step(0.0, 1.0, 0.8, comp=1)
    assert t2 == want_ts[0]
