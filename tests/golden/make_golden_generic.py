#!/usr/bin/env python
"""Golden fixture for SURVEY.md 8(f).4: genericHam / genericPartial (Hamiltonians/generic_ham.py, generic_partial.py) over a
user dynSys, produced by the LITERAL reference imported from /root/reference through oracle/ref_shim.py, with the numpy
restatement (oracle/generic.py) asserted bit-identical.

    python tests/golden/make_golden_generic.py     (only where /root/reference exists)

No class of the reference implements the dynSys API these functions call (get_opt_u / get_opt_v / dynamics), so the dynSys
is the one the API was written for, a Dubins car with disturbances (oracle/generic.py: DubinsCar, plain numpy).  Recorded
per mode case on an air3D-shaped 21x17x13 grid: ham and the three alphas of one RHS, termLaxFriedrichs ydot + stepBound,
three single-step odeCFL3 and odeCFL2 calls (t and y), odeCFL2 over termRestrictUpdate, and one HJIPDE_solve through schemeData.dynSys (hji_solver.py:413-415).
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
from LevelSetPy.ExplicitIntegration import (odeCFL2, odeCFL3, odeCFLset, termLaxFriedrichs, termRestrictUpdate,  # noqa: E402
                                             artificialDissipationGLF)
from LevelSetPy.Hamiltonians import genericHam, genericPartial  # noqa: E402
from LevelSetPy.ValueFuncs import HJIPDE_solve  # noqa: E402

from oracle import hj_oracle as orc  # noqa: E402
from oracle import generic as ogen  # noqa: E402
from make_golden import col, icol, same, perturb  # noqa: E402

CASES = (("default", {}), ("umax_dmin", dict(uMode="max", dMode="min")), ("forward", dict(tMode="forward")))
DYN = dict(speed=1.3, wMax=0.9, dMax=[0.15, 0.25, 0.1])


def main():
    out = {}
    N = [21, 17, 13]
    g = createGrid(col([-6, -10, 0]), col([20, 10, 2 * np.pi * (1 - 1 / N[2])]), icol(N), pdDims=2)
    data0 = perturb(g, shapeCylinder(g, 2, np.zeros((3, 1)), 5), 41)
    out["data0"] = data0
    out["grid_min"] = np.asarray(g.min).reshape(-1)
    out["grid_max"] = np.asarray(g.max).reshape(-1)
    out["grid_N"] = np.asarray(g.N).reshape(-1).astype(np.int64)
    out["periodic"] = np.array([g.bdry[d].__name__ == "addGhostPeriodic" for d in range(g.dim)])
    out["speed"], out["wMax"], out["dMax"] = DYN["speed"], DYN["wMax"], np.array(DYN["dMax"])
    opts = odeCFLset(Bundle({"factorCFL": 0.8, "singleStep": "on"}))
    y0 = np.expand_dims(data0.flatten(), 1)
    for tag, modes in CASES:
        dyn = ogen.DubinsCar(**DYN)
        rsd = Bundle(dict(grid=g, dynSys=dyn, hamFunc=genericHam, partialFunc=genericPartial,
                          dissFunc=artificialDissipationGLF, CoStateCalc=upwindFirstWENO5a, **modes))
        osd = orc.OracleSchemeData(grid=g, dynSys=dyn, hamFunc=ogen.generic_ham, partialFunc=ogen.generic_partial, **modes)
        # the hooks one by one, on the derivatives of data0
        dL, dR, dC = [], [], []
        for d in range(3):
            L, R = upwindFirstWENO5a(g, data0, d)
            dL.append(np.asarray(L)); dR.append(np.asarray(R)); dC.append(0.5 * (np.asarray(L) + np.asarray(R)))
        lo = [min(np.min(dL[d]), np.min(dR[d])) for d in range(3)]
        hi = [max(np.max(dL[d]), np.max(dR[d])) for d in range(3)]
        import copy
        r1, o1 = copy.copy(rsd), copy.copy(osd)
        ham = genericHam(0.0, data0, dC, r1)
        same(ham, ogen.generic_ham(0.0, data0, dC, o1), tag + " ham")
        out[tag + "_ham"] = np.asarray(ham)
        out[tag + "_derivMin"], out[tag + "_derivMax"] = np.array(lo), np.array(hi)
        for d in range(3):
            a = genericPartial(0.0, data0, lo, hi, r1, d)
            same(a, ogen.generic_partial(0.0, data0, lo, hi, o1, d), tag + " alpha %d" % d)
            out[tag + "_alpha%d" % d] = np.asarray(a, dtype=np.float64)
        ydot, sb, _ = termLaxFriedrichs(0.0, y0, rsd)
        oydot, osb = orc.term_lax_friedrichs(0.0, y0, osd)
        same(ydot, oydot, tag + " ydot")
        same(sb, osb, tag + " stepBound")
        out[tag + "_ydot"], out[tag + "_stepBound"] = np.asarray(ydot), float(sb)
        t, to, y, yo, ts = 0.0, 0.0, y0, y0, []
        for k in range(3):
            t, y, _ = odeCFL3(termLaxFriedrichs, [t, 1.0], y, opts, rsd)
            to, yo, _ = orc.ode_cfl3([to, 1.0], yo, osd, factor_cfl=0.8, single_step=True)
            same(t, to, tag + " t step %d" % k)
            same(y, yo, tag + " y step %d" % k)
            ts.append(float(t))
        out[tag + "_t"], out[tag + "_y"] = np.array(ts), np.asarray(y)
        # odeCFL2 (ode_cfl_2.py) over the same hooks: three single steps
        t, to, y, yo, ts = 0.0, 0.0, y0, y0, []
        for k in range(3):
            t, y, _ = odeCFL2(termLaxFriedrichs, [t, 1.0], y, opts, rsd)
            to, yo, _ = orc.ode_cfl2([to, 1.0], yo, osd, factor_cfl=0.8, single_step=True)
            same(t, to, tag + " rk2 t step %d" % k)
            same(y, yo, tag + " rk2 y step %d" % k)
            ts.append(float(t))
        out[tag + "_rk2_t"], out[tag + "_rk2_y"] = np.array(ts), np.asarray(y)
    # termRestrictUpdate(positive=False) around the generic term, odeCFL2 -- what the RCBRT notebooks drive (y of shape (n,))
    dyn = ogen.DubinsCar(**DYN)
    inner = Bundle(dict(grid=g, dynSys=dyn, hamFunc=genericHam, partialFunc=genericPartial,
                        dissFunc=artificialDissipationGLF, CoStateCalc=upwindFirstWENO5a))
    rsd = Bundle(dict(innerFunc=termLaxFriedrichs, innerData=inner, positive=False))
    osd = orc.OracleSchemeData(grid=g, dynSys=dyn, hamFunc=ogen.generic_ham, partialFunc=ogen.generic_partial)
    yflat = data0.flatten()
    t, to, y, yo, ts = 0.0, 0.0, yflat, yflat, []
    for k in range(3):
        t, y, _ = odeCFL2(termRestrictUpdate, [t, 1.0], y, opts, rsd)
        to, yo, _ = orc.ode_cfl2([to, 1.0], yo, osd, factor_cfl=0.8, single_step=True, restrict=False)
        same(t, to, "restricted rk2 t step %d" % k)
        same(y, yo, "restricted rk2 y step %d" % k)
        ts.append(float(t))
    out["restrict_neg_rk2_t"], out["restrict_neg_rk2_y"] = np.array(ts), np.asarray(y)
    # the driver: schemeData.dynSys alone makes HJIPDE_solve install genericHam / genericPartial (hji_solver.py:413-415)
    dyn = ogen.DubinsCar(**DYN)
    tau = np.array([0.0, 0.15, 0.3])
    rsd = Bundle(dict(grid=g, dynSys=dyn, uMode="min", dMode="max", CoStateCalc=upwindFirstWENO5a))
    data, _, _ = HJIPDE_solve(data0, tau, rsd, "minVOverTime", Bundle(dict(quiet=True, keepLast=True)))
    osd = orc.OracleSchemeData(grid=g, dynSys=dyn, hamFunc=ogen.generic_ham, partialFunc=ogen.generic_partial, uMode="min", dMode="max")
    odata, dts, _ = orc.hji_solve(data0, tau, osd, "minVOverTime")
    same(np.asarray(data).reshape(g.shape), np.asarray(odata).reshape(g.shape), "HJIPDE_solve(dynSys)")
    out["hji_tau"], out["hji_data"], out["hji_dts"] = tau, np.asarray(odata).reshape(g.shape), np.array(dts)
    path = os.path.join(HERE, "generic_dyn.npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%d arrays, %.1f kB): oracle == literal reference bit for bit" % (path, len(out), os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    main()
