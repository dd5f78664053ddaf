#!/usr/bin/env python
"""Developer tool: what the two-pass stage of genericHam / genericPartial costs at full size.

    python tools/time_generic.py [N] [steps]          (GPU box)

Steps a DubinsCar dynSys (csrc/hj_systems.cuh: GenericF<DubinsCarDyn>) and, beside it, DubinsVehicleRel on the same
N^3 air3D grid with the state resident, and prints one JSON line per system: ms per TVD-RK3 step (wall clock around
synchronised steps, the generic path has three host round trips per step by construction), point-steps/s, and for the
generic path the time of the reduce-only pre-pass (hj_deriv_range) on its own.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import levelsetpy_b200 as lsp  # noqa: E402
from levelsetpy_b200 import _lib as L  # noqa: E402
from levelsetpy_b200.integration import rk3_step_resident  # noqa: E402
from levelsetpy_b200.term import prepare_scheme  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    g = lsp.createGrid(np.array([-6.0, -10.0, 0.0]), np.array([20.0, 10.0, 2 * np.pi * (1 - 1 / N)]), np.array([N, N, N]),
                       pdDims=2)
    x = [np.asarray(v).reshape(-1) for v in g.vs]
    d0 = (np.sqrt(x[0][:, None, None] ** 2 + x[1][None, :, None] ** 2) - 5.0
          + 0.1 * np.sin(x[2][None, None, :] + 0.3 * x[0][:, None, None]))
    car = lsp.DubinsCar(speed=5.0, wMax=1.0, dMax=[0.0, 0.0, 1.0])
    rel = lsp.DubinsVehicleRel(g, 5, 1)
    cases = (
        ("generic(DubinsCar)", dict(dynSys=car, hamFunc=lsp.genericHam, partialFunc=lsp.genericPartial)),
        ("DubinsVehicleRel", dict(hamFunc=rel.hamiltonian, partialFunc=rel.dissipation)),
    )
    for name, hooks in cases:
        sd = lsp.Bundle(dict(grid=g, dissFunc=lsp.artificialDissipationGLF, CoStateCalc=lsp.upwindFirstWENO5a, **hooks))
        eng, ad = prepare_scheme(sd)
        eng.upload(d0.reshape(-1, 1))
        t = 0.0
        for _ in range(3):
            t, _ = rk3_step_resident(eng, ad, g, t, 1e9, 0.8, 1e9, L.COMP_MIN_OVER_TIME)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            t, dt = rk3_step_resident(eng, ad, g, t, 1e9, 0.8, 1e9, L.COMP_MIN_OVER_TIME)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / steps
        rec = dict(system=name, grid=[N, N, N], steps=steps, ms_per_step=ms, point_steps_per_s=N ** 3 / (ms * 1e-3), dt=dt)
        if ad.dynamic:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(steps):
                eng.deriv_range(None, 1)
            rec["deriv_range_ms"] = (time.perf_counter() - t0) * 1e3 / steps
            t0 = time.perf_counter()
            for _ in range(steps):
                eng.alpha_max()
            rec["alpha_max_ms"] = (time.perf_counter() - t0) * 1e3 / steps
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
