"""Times the LITERAL reference (robotsorcerer/LevelSetPy, read-only at /root/reference, driven through oracle/ref_shim.py:
fake cupy = numpy) on BASELINE.json's configs[0] -- one odeCFL3(termLaxFriedrichs) TVD-RK3 step of air3D 101^3 per call --
next to the numpy oracle port on the same box, and writes profiles/r02_reference_literal_cpu.json.  Runs only where the
reference checkout exists (the build container); bench.py --impl reference quotes the committed file, because the
reference cannot travel to the GPU box.     python tools/time_reference_literal.py [n=101] [steps=3]"""
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import ref_shim  # noqa: E402

ref_shim.install()

from LevelSetPy.Utilities import Bundle  # noqa: E402
from LevelSetPy.Grids import createGrid  # noqa: E402
from LevelSetPy.SpatialDerivative import upwindFirstWENO5a  # noqa: E402
from LevelSetPy.ExplicitIntegration import odeCFL3, odeCFLset, termLaxFriedrichs, artificialDissipationGLF  # noqa: E402
from LevelSetPy.DynamicalSystems import DubinsVehicleRel  # noqa: E402

from oracle import hj_oracle as orc  # noqa: E402
from oracle import systems as osys  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 101
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    g = createGrid(np.array([[-6.0, -10.0, 0.0]]).T, np.array([[20.0, 10.0, 2 * np.pi * (1 - 1 / n)]]).T,
                   np.array([[n, n, n]]).T, 2)
    data0 = np.sqrt(np.asarray(g.xs[0]) ** 2 + np.asarray(g.xs[1]) ** 2) - 5.0
    s = DubinsVehicleRel(g, 5, 1)
    sd = Bundle(dict(grid=g, hamFunc=s.hamiltonian, partialFunc=s.dissipation, dissFunc=artificialDissipationGLF,
                     CoStateCalc=upwindFirstWENO5a))
    opts = odeCFLset(Bundle(dict(factorCFL=0.8, singleStep="on")))
    y = data0.reshape(-1, 1)
    t = 0.0
    t0 = time.perf_counter()
    for _ in range(steps):
        t, y, _ = odeCFL3(termLaxFriedrichs, [t, 1e9], y, opts, sd)
    ref_s = (time.perf_counter() - t0) / steps
    o = osys.DubinsVehicleRel(g, 5, 1)
    osd = orc.OracleSchemeData(grid=g, hamFunc=o.hamiltonian, partialFunc=o.dissipation)
    yo, to = data0.reshape(-1, 1), 0.0
    t0 = time.perf_counter()
    for _ in range(steps):
        to, yo, _ = orc.ode_cfl3([to, 1e9], yo, osd, factor_cfl=0.8, single_step=True)
    port_s = (time.perf_counter() - t0) / steps
    out = {"what": "one odeCFL3(termLaxFriedrichs) TVD-RK3 step, air3D %d^3, WENO5a + GLF, factorCFL 0.8" % n,
           "literal_reference_s_per_step": ref_s, "literal_reference_point_steps_per_s": n ** 3 / ref_s,
           "oracle_port_s_per_step": port_s, "oracle_port_point_steps_per_s": n ** 3 / port_s,
           "port_over_literal": ref_s / port_s, "identical_result": bool(np.array_equal(np.asarray(y), yo) and t == to),
           "steps": steps, "cores_used": 1, "host_cores": os.cpu_count(),
           "where": "build container (no GPU): the literal reference runs on numpy through oracle/ref_shim.py",
           "numpy": np.__version__}
    with open(os.path.join(ROOT, "profiles", "r02_reference_literal_cpu.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
