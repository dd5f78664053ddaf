"""Small cases of every kernel family for compute-sanitizer (profiles/run_gpu_sanitize.sh): one TVD-RK3 step each of
air3D 21x17x13 (plane-ring kernel, both WENO modes, reductions), the 4-D and 6-D product systems (dimension-split path
with ghost warps), a Flock batch, a 2-slab LocalWorld over the peer-memory halo transport (pieces protocol) and the
gather backend.  Prints one line per case; any sanitizer finding fails the run."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import levelsetpy_b200 as lsp
    from levelsetpy_b200 import _lib as L
    from levelsetpy_b200.integration import rk3_step_resident
    from levelsetpy_b200.slab import LocalWorld
    from levelsetpy_b200.term import prepare_scheme
    rng = np.random.default_rng(5)
    fmax = np.finfo(np.float64).max

    def bundle(g, s, weno="as_shipped"):
        return lsp.Bundle(dict(grid=g, hamFunc=s.hamiltonian, partialFunc=s.dissipation, wenoMode=weno,
                               dissFunc=lsp.artificialDissipationGLF, CoStateCalc=lsp.upwindFirstWENO5a))

    def air3d(N):
        g = lsp.createGrid(np.array([-6.0, -10.0, 0.0]), np.array([20.0, 10.0, 2 * np.pi * (1 - 1 / N[2])]), np.array(N), pdDims=2)
        x = np.meshgrid(*[np.asarray(v).reshape(-1) for v in g.vs], indexing="ij")
        return g, np.sqrt(x[0] ** 2 + x[1] ** 2) - 5 + 0.3 * np.sin(x[2]) + 0.05 * rng.standard_normal(g.shape)

    def run(name, sd, g, d0, backend=L.BACKEND_AUTO, steps=1):
        eng, ad = prepare_scheme(sd)
        eng.set_backend(backend)
        eng.upload(d0)
        t = 0.0
        for _ in range(steps):
            t, dt = rk3_step_resident(eng, ad, g, t, 1.0, 0.8, fmax, L.COMP_MIN_OVER_TIME)
        y = eng.download(shape=g.shape)
        assert np.all(np.isfinite(y))
        eng.set_backend(L.BACKEND_AUTO)
        print("%-34s t=%.6g  |y|max=%.6g" % (name, t, float(np.abs(y).max())), flush=True)

    g, d0 = air3d([21, 17, 13])
    # genericHam / genericPartial over a device dynSys: reduce-only pre-pass (hj_deriv_range) + the GenericF stage kernels
    car = lsp.Bundle(dict(grid=g, dynSys=lsp.DubinsCar(1.3, 0.9, [0.15, 0.25, 0.1]), hamFunc=lsp.genericHam,
                          partialFunc=lsp.genericPartial, dissFunc=lsp.artificialDissipationGLF,
                          CoStateCalc=lsp.upwindFirstWENO5a))
    for be, nm in ((L.BACKEND_TMA, "tma"), (L.BACKEND_GATHER, "gather")):
        run("generic DubinsCar 21x17x13 " + nm, car, g, d0, be)
    if "--generic-only" in sys.argv:
        return
    for weno in ("as_shipped", "intended"):
        run("air3d 21x17x13 tma " + weno, bundle(g, lsp.DubinsVehicleRel(g, 5, 1), weno), g, d0, L.BACKEND_TMA)
    run("air3d 21x17x13 gather", bundle(g, lsp.DubinsVehicleRel(g, 5, 1)), g, d0, L.BACKEND_GATHER)
    g, d0 = air3d([40, 37, 70])
    run("air3d 40x37x70 tma (multi-tile)", bundle(g, lsp.DubinsVehicleRel(g, 5, 1)), g, d0, L.BACKEND_TMA)

    g4 = lsp.createGrid(-np.ones(4), np.ones(4), np.array([12, 9, 18, 34]))
    x = np.meshgrid(*[np.asarray(v).reshape(-1) for v in g4.vs], indexing="ij")
    d4 = np.sqrt((x[0] - x[2]) ** 2 + (x[1] - x[3]) ** 2) - 0.2 + 0.02 * rng.standard_normal(g4.shape)
    s4 = lsp.ProductSystem(g4, [lsp.DoubleIntegrator(g4, 1.0), lsp.DoubleIntegrator(g4, 0.6)])
    run("dint4d 12x9x18x34 split", bundle(g4, s4), g4, d4, L.BACKEND_TMA)

    N = [10, 9, 8, 7, 9, 12]
    lo = [-6, -10, 0, -6, -10, 0.]
    hi = [20, 10, 2 * np.pi * (1 - 1 / N[2]), 20, 10, 2 * np.pi * (1 - 1 / N[5])]
    g6 = lsp.createGrid(np.array(lo), np.array(hi), np.array(N), pdDims=[2, 5])
    x = np.meshgrid(*[np.asarray(v).reshape(-1) for v in g6.vs], indexing="ij")
    d6 = np.minimum(np.sqrt(x[0] ** 2 + x[1] ** 2) - 5, np.sqrt(x[3] ** 2 + x[4] ** 2) - 5) + 0.3 * np.sin(x[2] + x[5])
    s6 = lsp.ProductSystem(g6, [lsp.DubinsVehicleRel(g6, 5, 1), lsp.DubinsVehicleRel(g6, 4, 1.2)])
    run("dubins6d 10x9x8x7x9x12 split", bundle(g6, s6), g6, d6, L.BACKEND_TMA)

    w = LocalWorld(bundle(g6, s6), 2, backend=L.BACKEND_TMA, pieces=3)
    w.upload(d6)
    t, _ = w.step(0.0, 1.0, 0.8, comp=L.COMP_MIN_OVER_TIME)
    assert np.all(np.isfinite(w.download()))
    print("%-34s t=%.6g  (peer halos, 3 pieces)" % ("dubins6d 2 slabs LocalWorld", t), flush=True)

    import bench
    sds, data = bench.flock_batch_setup(lsp, 3, 21)
    bs = lsp.BatchSolver(sds, device=0)
    bs.upload(np.stack(data))
    tt = bs.step(1e9, 0.8, L.COMP_MIN_OVER_TIME)[0]
    print("%-34s t=%s" % ("flock batch 3 x 21^3", np.asarray(tt)), flush=True)


if __name__ == "__main__":
    main()
