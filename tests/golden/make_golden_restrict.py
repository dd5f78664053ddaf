#!/usr/bin/env python
"""Golden fixture for the "next" row SURVEY.md 8(f).1: termRestrictUpdate + odeCFL2 (and odeCFL3 on the restricted
term), produced by the LITERAL reference imported from /root/reference through oracle/ref_shim.py, with the numpy
oracle (oracle/hj_oracle.py: term_restrict_update, ode_cfl2, ode_cfl3_restricted) asserted bit-identical.

    python tests/golden/make_golden_restrict.py     (only where /root/reference exists)

Reference behaviour recorded here: termRestrictUpdate squeezes its ydot to (n,) (term_restrict_update.py:92,:94), so the
integrators are driven with y of shape (n,) -- with (n,1) the reference's ``y + deltaT*ydot`` broadcasts to (n,n).
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
from LevelSetPy.SpatialDerivative import upwindFirstWENO5a  # noqa: E402
from LevelSetPy.ExplicitIntegration import (  # noqa: E402
    odeCFL2, odeCFL3, odeCFLset, termLaxFriedrichs, termRestrictUpdate, artificialDissipationGLF)
from LevelSetPy.DynamicalSystems import DubinsVehicleRel, DoubleIntegrator  # noqa: E402

from oracle import hj_oracle as orc  # noqa: E402
from oracle import systems as osys  # noqa: E402
from make_golden import col, icol, same, perturb, ref_sd, orc_sd  # noqa: E402


def main():
    out = {}
    N = [21, 17, 13]
    g = createGrid(col([-6, -10, 0]), col([20, 10, 2 * np.pi * (1 - 1 / N[2])]), icol(N), pdDims=2)
    base = shapeCylinder(g, 2, np.zeros((3, 1)), 5)
    d3 = perturb(g, base, 21)
    g2 = createGrid(col([-1, -1]), col([1, 1]), icol([33, 20]))
    d2 = perturb(g2, np.sqrt(np.asarray(g2.xs[0]) ** 2 + np.asarray(g2.xs[1]) ** 2) - 0.4, 22)
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
        yflat = data0.flatten()                                   # (n,): see the module docstring
        y1 = np.expand_dims(yflat, 1)
        for positive in (True, False):
            tag = "%s_%s" % (name, "pos" if positive else "neg")
            rsd = Bundle(dict(innerFunc=termLaxFriedrichs, innerData=ref_sd(grid, rf()), positive=positive))
            osd = orc_sd(grid, of())
            # one RHS
            ydot, sb, _ = termRestrictUpdate(0.0, yflat, rsd)
            oydot, osb = orc.term_restrict_update(0.0, yflat, osd, positive)
            same(ydot, oydot, tag + " restricted ydot")
            same(sb, osb, tag + " stepBound")
            out[tag + "_ydot"] = np.asarray(ydot)
            # odeCFL2 and odeCFL3 on the restricted term, 3 single steps each
            for order, fn in ((2, odeCFL2), (3, odeCFL3)):
                rsd = Bundle(dict(innerFunc=termLaxFriedrichs, innerData=ref_sd(grid, rf()), positive=positive))
                t, y, to, yo, ts = 0.0, yflat, 0.0, yflat, []
                for k in range(3):
                    t, y, _ = fn(termRestrictUpdate, [t, 1.0], y, opts, rsd)
                    if order == 2:
                        to, yo, _ = orc.ode_cfl2([to, 1.0], yo, osd, factor_cfl=0.8, single_step=True, restrict=positive)
                    else:
                        to, yo, _ = orc.ode_cfl3_restricted([to, 1.0], yo, osd, positive, factor_cfl=0.8, single_step=True)
                    same(t, to, "%s rk%d t step %d" % (tag, order, k))
                    same(y, yo, "%s rk%d y step %d" % (tag, order, k))
                    ts.append(float(t))
                out["%s_rk%d_t" % (tag, order)] = np.array(ts)
                out["%s_rk%d_y" % (tag, order)] = np.asarray(y)
        # plain odeCFL2(termLaxFriedrichs) with y (n,1)
        t, y, to, yo, ts = 0.0, y1, 0.0, y1, []
        sd_r, sd_o = ref_sd(grid, rf()), orc_sd(grid, of())
        for k in range(3):
            t, y, _ = odeCFL2(termLaxFriedrichs, [t, 1.0], y, opts, sd_r)
            to, yo, _ = orc.ode_cfl2([to, 1.0], yo, sd_o, factor_cfl=0.8, single_step=True)
            same(t, to, "%s plain rk2 t step %d" % (name, k))
            same(y, yo, "%s plain rk2 y step %d" % (name, k))
            ts.append(float(t))
        out[name + "_plain_rk2_t"] = np.array(ts)
        out[name + "_plain_rk2_y"] = np.asarray(y)
    np.savez_compressed(os.path.join(HERE, "restrict_rk2.npz"), **out)
    print("wrote restrict_rk2.npz (oracle == reference, bit-exact)")


if __name__ == "__main__":
    main()
