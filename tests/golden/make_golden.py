#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the LITERAL reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports the reference through oracle/ref_shim.py (fake cupy / matplotlib / skimage),
evaluates the hot path on small grids, and stores inputs + outputs as .npz fixtures.
While doing so it also checks the numpy oracle (oracle/hj_oracle.py, weno='as_shipped')
against the reference bit-for-bit and aborts on any mismatch -- so a committed fixture is
both a reference output and an oracle output.

The one deviation from stock reference code: ``Flock.dissipation`` ends with
``np.maximum.reduce(alphas, dtype=object)`` on a ragged list, which raises on numpy >= 1.24
(flock.py:257).  For the Flock fixture only, that method is replaced by a version that ends
with the scalar maximum the old numpy produced; everything else in the reference is untouched.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import ref_shim  # noqa: E402

ref_shim.install()

from LevelSetPy.Utilities import Bundle  # noqa: E402
from LevelSetPy.Grids import createGrid  # noqa: E402
from LevelSetPy.InitialConditions import shapeCylinder  # noqa: E402
from LevelSetPy.SpatialDerivative import upwindFirstWENO5a, upwindFirstWENO5  # noqa: E402
from LevelSetPy.BoundaryCondition import addGhostExtrapolate, addGhostPeriodic  # noqa: E402
from LevelSetPy.ExplicitIntegration import (  # noqa: E402
    odeCFL3, odeCFLset, termLaxFriedrichs, artificialDissipationGLF)
from LevelSetPy.DynamicalSystems import DubinsVehicleRel, DoubleIntegrator, Bird, Flock  # noqa: E402
from LevelSetPy.ValueFuncs import HJIPDE_solve  # noqa: E402

from oracle import hj_oracle as orc  # noqa: E402
from oracle import systems as osys  # noqa: E402


def col(x):
    return np.asarray(x, dtype=np.float64).reshape(-1, 1)


def icol(x):
    return np.asarray(x, dtype=np.int64).reshape(-1, 1)


def same(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape or not np.array_equal(a, b):
        raise SystemExit("ORACLE != REFERENCE for %s (max abs diff %.3e)" % (
            what, float(np.max(np.abs(a - b))) if a.shape == b.shape else float("nan")))


def perturb(grid, base, seed):
    """SDF + smooth wave + seeded noise so every stencil branch is exercised."""
    rng = np.random.default_rng(seed)
    wave = 0.0
    for d in range(grid.dim):
        wave = wave + 0.3 * np.sin(1.7 * np.asarray(grid.xs[d]) + 0.4 * d)
    return np.ascontiguousarray(base + wave + 0.05 * rng.standard_normal(base.shape))


def ref_sd(grid, system):
    return Bundle(dict(grid=grid, hamFunc=system.hamiltonian, partialFunc=system.dissipation,
                       dissFunc=artificialDissipationGLF, CoStateCalc=upwindFirstWENO5a))


def orc_sd(grid, system):
    return orc.OracleSchemeData(grid=grid, hamFunc=system.hamiltonian, partialFunc=system.dissipation)


def run_case(name, grid, ref_system_factory, orc_system_factory, data0, nsteps, t_end, extra=None):
    """derivL/R per dim, one RHS, and ``nsteps`` single-step odeCFL3 calls, reference vs oracle."""
    out = dict(extra or {})
    out["data0"] = data0
    out["grid_min"] = np.asarray(grid.min).reshape(-1)
    out["grid_max"] = np.asarray(grid.max).reshape(-1)
    out["grid_N"] = np.asarray(grid.N).reshape(-1).astype(np.int64)
    out["grid_dx"] = np.asarray(grid.dx).reshape(-1)
    out["periodic"] = np.array([grid.bdry[d].__name__ == "addGhostPeriodic" for d in range(grid.dim)])
    for d in range(grid.dim):
        out["vs%d" % d] = np.asarray(grid.vs[d]).reshape(-1)
        L, R = upwindFirstWENO5a(grid, data0, d)
        L2, R2 = upwindFirstWENO5(grid, data0, d)          # the alias, upwind_first_weno5.py:11-48
        same(L, L2, "%s WENO5 alias L%d" % (name, d))
        oL, oR = orc.upwind_first_weno5a(grid, data0, d, "as_shipped")
        same(L, oL, "%s derivL[%d]" % (name, d))
        same(R, oR, "%s derivR[%d]" % (name, d))
        out["derivL%d" % d], out["derivR%d" % d] = np.asarray(L), np.asarray(R)
        iL, iR = orc.upwind_first_weno5a(grid, data0, d, "intended")   # oracle-only (documented in the test)
        out["intended_derivL%d" % d], out["intended_derivR%d" % d] = iL, iR

    y0 = np.expand_dims(data0.flatten(), 1)
    rsys, osys_ = ref_system_factory(), orc_system_factory()
    ydot, sb, _ = termLaxFriedrichs(0.0, y0, ref_sd(grid, rsys))
    oydot, osb = orc.term_lax_friedrichs(0.0, y0, orc_sd(grid, osys_), "as_shipped")
    same(ydot, oydot, name + " ydot")
    same(sb, osb, name + " stepBound")
    out["ydot"], out["stepBound"] = np.asarray(ydot), np.float64(sb)

    rsys, osys_ = ref_system_factory(), orc_system_factory()
    opts = odeCFLset(Bundle({"factorCFL": 0.8, "singleStep": "on"}))
    sd_r, sd_o = ref_sd(grid, rsys), orc_sd(grid, osys_)
    t, y, to, yo = 0.0, y0, 0.0, y0
    ts = []
    for k in range(nsteps):
        t, y, _ = odeCFL3(termLaxFriedrichs, [t, t_end], y, opts, sd_r)
        to, yo, _ = orc.ode_cfl3([to, t_end], yo, sd_o, factor_cfl=0.8, single_step=True)
        same(t, to, "%s t after step %d" % (name, k))
        same(y, yo, "%s y after step %d" % (name, k))
        ts.append(float(t))
    out["t_steps"] = np.array(ts)
    out["y_final"] = np.asarray(y)
    rsys, osys_ = ref_system_factory(), orc_system_factory()
    _, yi, _ = orc.ode_cfl3([0.0, ts[-1]], y0, orc_sd(grid, osys_), factor_cfl=0.8, weno="intended")
    out["intended_y_final"] = yi                                       # oracle-only
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote %-28s steps=%d t=%s  (oracle == reference, bit-exact)" % (name + ".npz", nsteps, ts))


def main():
    # ---- 1. air3D-like 3-D, dims 0,1 extrapolate, dim 2 periodic (SURVEY.md Appendix A verification grid)
    N = [21, 17, 13]
    g = createGrid(col([-6, -10, 0]), col([20, 10, 2 * np.pi * (1 - 1 / N[2])]), icol(N), pdDims=2)
    base = shapeCylinder(g, 2, np.zeros((3, 1)), 5)
    run_case("air3d_21x17x13", g, lambda: DubinsVehicleRel(g, 5, 1), lambda: osys.DubinsVehicleRel(g, 5, 1),
             perturb(g, base, 0), nsteps=3, t_end=1.0, extra=dict(u_bound=5.0, w_bound=1.0))
    # the unperturbed cylinder: derivative along dim 2 is identically zero -> eps == 1e-99 path
    run_case("air3d_cyl_21x17x13", g, lambda: DubinsVehicleRel(g, 5, 1), lambda: osys.DubinsVehicleRel(g, 5, 1),
             np.ascontiguousarray(base), nsteps=2, t_end=1.0, extra=dict(u_bound=5.0, w_bound=1.0))

    # ---- 2. double integrator 2-D, all extrapolate; closed-form stepBound 1/(max|x2|/dx0 + |u|/dx1)
    g2 = createGrid(col([-1, -1]), col([1, 1]), icol([51, 51]))
    base2 = shapeCylinder(g2, [], np.zeros((2, 1)), 0.3)
    run_case("dint_51x51", g2, lambda: DoubleIntegrator(g2, 1), lambda: osys.DoubleIntegrator(g2, 1),
             perturb(g2, base2, 1), nsteps=3, t_end=1.0, extra=dict(u_bound=1.0))
    g2b = createGrid(col([-1.5, -0.7]), col([1.2, 0.9]), icol([33, 20]))
    base2b = shapeCylinder(g2b, [], np.zeros((2, 1)), 0.4)
    run_case("dint_33x20", g2b, lambda: DoubleIntegrator(g2b, 0.7), lambda: osys.DoubleIntegrator(g2b, 0.7),
             perturb(g2b, base2b, 2), nsteps=2, t_end=1.0, extra=dict(u_bound=0.7))

    # ---- 3. Flock of 4 birds on one 3-D grid (dim 2 periodic)
    def patched_dissipation(self, t, data, derivMin, derivMax, schemeData, dim):
        vehicles = [x for x in self.vehicles if x is not self.vehicles[self.attacked_idx]]
        alphas = [v.dissipation_abs(t, data, derivMin, derivMax, schemeData, dim) for v in vehicles]
        alphas.append(self.vehicles[self.attacked_idx].dissipation(t, data, derivMin, derivMax, schemeData, dim))
        import cupy as cp                                                    # the shim's fake
        return cp.asarray(max(float(np.asarray(a).reshape(-1)[0]) for a in alphas))   # scalarised flock.py:257-258
    Flock.dissipation = patched_dissipation
    Nf = 15
    gf = createGrid(col([-1, -1, -np.pi]), col([1, 1, np.pi * (1 - 2 / Nf)]), icol([Nf, Nf, Nf]), pdDims=2)
    wb = [0.8, 1.0, 1.3, 0.6]
    xyw = [[0.1 * j - 0.05, 0.2 * j - 0.3, 0.3 * j + 0.1] for j in range(4)]

    def ref_flock():
        birds = [Bird(gf, 1.0, wb[j], init_xyw=np.array([xyw[j]]).T.copy(), label=j, neigh_rad=3) for j in range(4)]
        return Flock(gf, birds, label=1)

    def orc_flock():
        birds = [osys.Bird(gf, 1.0, wb[j], init_xyw=np.array(xyw[j]), label=j, neigh_rad=3) for j in range(4)]
        return osys.Flock(gf, birds)
    basef = shapeCylinder(gf, 2, np.zeros((3, 1)), 0.3)
    run_case("flock4_15x15x15", gf, ref_flock, orc_flock, perturb(gf, basef, 3), nsteps=3, t_end=1.0,
             extra=dict(w_bounds=np.array(wb), init_xyw=np.array(xyw), u_bound=1.0))

    # ---- 4. boundary-condition fixtures (addGhostExtrapolate incl. towardZero and sign(0)=0, addGhostPeriodic)
    rng = np.random.default_rng(7)
    a = rng.standard_normal((6, 5, 7))
    a[0, 2, :] = 0.0                     # sign(edge) == 0 -> zero slope (add_ghost_extrapolate.py:96)
    a[:, 4, 6] = 0.0
    bc = dict(a=a)
    for d in range(3):
        for tz in (False, True):
            r = np.asarray(addGhostExtrapolate(a, d, 3, Bundle(dict(towardZero=tz))))
            same(r, orc.add_ghost_extrapolate(a, d, 3, tz), "addGhostExtrapolate d%d tz%d" % (d, tz))
            bc["extrap_d%d_tz%d" % (d, int(tz))] = r
        r = np.asarray(addGhostPeriodic(a, d, 3, None))
        same(r, orc.add_ghost_periodic(a, d, 3), "addGhostPeriodic d%d" % d)
        bc["periodic_d%d" % d] = r
    np.savez_compressed(os.path.join(HERE, "ghost_cells_6x5x7.npz"), **bc)
    print("wrote ghost_cells_6x5x7.npz")

    # ---- 5. the driver: HJIPDE_solve(keepLast, minVOverTime) on the air3D grid
    rsys = DubinsVehicleRel(g, 5, 1)
    sd = ref_sd(g, rsys)
    tau = np.array([0.0, 0.05, 0.1])
    data0 = np.ascontiguousarray(base)
    res = HJIPDE_solve(data0, tau, sd, "minVOverTime", Bundle(dict(quiet=True, keepLast=True)))
    data_ref = np.asarray(res[0])
    data_orc, dts, ts = orc.hji_solve(data0, tau, orc_sd(g, osys.DubinsVehicleRel(g, 5, 1)), "minVOverTime")
    same(data_ref, data_orc, "HJIPDE_solve data")
    np.savez_compressed(os.path.join(HERE, "hji_air3d_21x17x13.npz"), data0=data0, tau=tau, data=data_ref,
                        dts=np.array(dts), ts=np.array(ts), grid_min=np.asarray(g.min).reshape(-1),
                        grid_max=np.asarray(g.max).reshape(-1), grid_N=np.asarray(g.N).reshape(-1))
    print("wrote hji_air3d_21x17x13.npz  steps=%d  (oracle == reference, bit-exact)" % len(dts))


if __name__ == "__main__":
    main()
