#!/usr/bin/env python
"""Golden fixture for the "next" row SURVEY.md 8(f).2: upwindFirstENO2 / upwindFirstENO3a (upwindFirstENO3 is an alias)
as schemeData.CoStateCalc, produced by the LITERAL reference imported from /root/reference through oracle/ref_shim.py,
with the numpy oracle (oracle/hj_oracle.py: upwind_first_eno2, upwind_first_eno3a) asserted bit-identical.

    python tests/golden/make_golden_eno.py     (only where /root/reference exists)

Recorded per case (air3D 21x17x13 with a periodic dim, double integrator 33x20, both with towardZero variants of the
extrapolated ghost cells off/on): derivL/derivR per dim, one termLaxFriedrichs RHS + stepBound, three single-step
odeCFL3 calls.  upwindFirstFirst (hji_solver's 'low' accuracy) raises IndexError in the reference as shipped
(upwind_first_first.py:60-62 assigns into an empty list) and is therefore not part of the path.
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
from LevelSetPy.SpatialDerivative import upwindFirstENO2, upwindFirstENO3a, upwindFirstENO3  # noqa: E402
from LevelSetPy.ExplicitIntegration import odeCFL3, odeCFLset, termLaxFriedrichs, artificialDissipationGLF  # noqa: E402
from LevelSetPy.DynamicalSystems import DubinsVehicleRel, DoubleIntegrator  # noqa: E402

from oracle import hj_oracle as orc  # noqa: E402
from oracle import systems as osys  # noqa: E402
from make_golden import col, icol, same, perturb, orc_sd  # noqa: E402

SCHEMES = (("eno2", upwindFirstENO2, orc.upwind_first_eno2), ("eno3a", upwindFirstENO3a, orc.upwind_first_eno3a))


def main():
    out = {}
    N = [21, 17, 13]
    g = createGrid(col([-6, -10, 0]), col([20, 10, 2 * np.pi * (1 - 1 / N[2])]), icol(N), pdDims=2)
    d3 = perturb(g, shapeCylinder(g, 2, np.zeros((3, 1)), 5), 31)
    g2 = createGrid(col([-1, -1]), col([1, 1]), icol([33, 20]))
    d2 = perturb(g2, np.sqrt(np.asarray(g2.xs[0]) ** 2 + np.asarray(g2.xs[1]) ** 2) - 0.4, 32)
    cases = (("air3d", g, d3, lambda: DubinsVehicleRel(g, 5, 1), lambda: osys.DubinsVehicleRel(g, 5, 1), dict(u_bound=5.0, w_bound=1.0)),
             ("dint", g2, d2, lambda: DoubleIntegrator(g2, 0.7), lambda: osys.DoubleIntegrator(g2, 0.7), dict(u_bound=0.7)))
    opts = odeCFLset(Bundle({"factorCFL": 0.8, "singleStep": "on"}))
    for name, grid, data0, rf, of, extra in cases:
        for k, v in extra.items():
            out["%s_%s" % (name, k)] = v
        out[name + "_data0"] = data0
        out[name + "_grid_min"] = np.asarray(grid.min).reshape(-1)
        out[name + "_grid_max"] = np.asarray(grid.max).reshape(-1)
        out[name + "_grid_N"] = np.asarray(grid.N).reshape(-1).astype(np.int64)
        out[name + "_periodic"] = np.array([grid.bdry[d].__name__ == "addGhostPeriodic" for d in range(grid.dim)])
        for tag, ref_fn, orc_fn in SCHEMES:
            for d in range(grid.dim):
                L, R = ref_fn(grid, data0, d)
                Lo, Ro = orc_fn(grid, data0, d)
                same(L, Lo, "%s %s derivL dim %d" % (name, tag, d))
                same(R, Ro, "%s %s derivR dim %d" % (name, tag, d))
                out["%s_%s_L%d" % (name, tag, d)] = np.asarray(L)
                out["%s_%s_R%d" % (name, tag, d)] = np.asarray(R)
            rs, os_ = rf(), of()
            rsd = Bundle(dict(grid=grid, hamFunc=rs.hamiltonian, partialFunc=rs.dissipation,
                              dissFunc=artificialDissipationGLF, CoStateCalc=ref_fn))
            osd = orc_sd(grid, os_)
            y = np.expand_dims(data0.flatten(), 1)
            ydot, sb, _ = termLaxFriedrichs(0.0, y, rsd)
            oydot, osb = orc.term_lax_friedrichs(0.0, y, osd, tag)
            same(ydot, oydot, "%s %s ydot" % (name, tag))
            same(sb, osb, "%s %s stepBound" % (name, tag))
            out["%s_%s_ydot" % (name, tag)] = np.asarray(ydot)
            out["%s_%s_stepBound" % (name, tag)] = float(sb)
            t, to, yo, ts = 0.0, 0.0, y, []
            for k in range(3):
                t, y, _ = odeCFL3(termLaxFriedrichs, [t, 1.0], y, opts, rsd)
                to, yo, _ = orc.ode_cfl3([to, 1.0], yo, osd, factor_cfl=0.8, single_step=True, weno=tag)
                same(t, to, "%s %s t step %d" % (name, tag, k))
                same(y, yo, "%s %s y step %d" % (name, tag, k))
                ts.append(float(t))
            out["%s_%s_t" % (name, tag)] = np.array(ts)
            out["%s_%s_y" % (name, tag)] = np.asarray(y)
        # the alias
        L3, R3 = upwindFirstENO3(grid, data0, 0)
        same(L3, out["%s_eno3a_L0" % name], name + " upwindFirstENO3 alias")
    path = os.path.join(HERE, "eno_schemes.npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%d arrays, %.1f kB): oracle == literal reference bit for bit" % (path, len(out), os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    main()
