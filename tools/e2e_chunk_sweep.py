"""Developer tooling: the pipelined host-buffer step (hj_ode_cfl3_step; what levelsetpy_b200.odeCFL3(..., singleStep='on')
calls for a host array) at air3D 512^3 for several chunk heights (hj_set_pipeline_planes), next to the bare full-duplex
transfer of one field each way (the floor of any per-step host round trip).  Every chunking must give the same bits.

    python tools/e2e_chunk_sweep.py [--n 512] [--planes 32,16,8,6,4,3] [--steps 6]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--planes", default="32,16,8,4,2,0")
    ap.add_argument("--steps", type=int, default=6)
    a = ap.parse_args()
    import torch
    import levelsetpy_b200 as lsp
    import bench
    from levelsetpy_b200.term import prepare_scheme

    g, data0 = bench.air3d_setup(lsp, a.n, a.n, a.n)
    sd = bench.scheme_for(lsp, g, "as_shipped")
    eng, ad = prepare_scheme(sd)
    eng.set_system(ad.system_id, ad.block(), list(enumerate(ad.tables(g))))
    nbytes = data0.size * 8
    out = {"grid": [a.n] * 3, "field_bytes": nbytes}

    # floor: one field up and one field down at the same time, pinned, two streams
    hin = torch.from_numpy(data0.reshape(-1).copy()).pin_memory()
    hout = torch.empty_like(hin).pin_memory()
    din = torch.empty(data0.size, dtype=torch.float64, device="cuda")
    dout = torch.zeros(data0.size, dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for name, up, down in (("h2d_only", True, False), ("d2h_only", False, True), ("duplex", True, True)):
        ts = []
        for _ in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if up:
                with torch.cuda.stream(s1):
                    din.copy_(hin, non_blocking=True)
            if down:
                with torch.cuda.stream(s2):
                    hout.copy_(dout, non_blocking=True)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        out[name + "_ms"] = 1e3 * min(ts[1:])
    del din, dout, hout
    print(json.dumps(out), flush=True)

    big = np.finfo(np.float64).max
    ref = None
    for spec in a.planes.split(","):
        planes, flag = int(spec), 0
        eng.set_pipeline_planes(planes)
        y = eng.pinned_out()
        y[:] = data0.reshape(-1)
        t = 0.0
        ts = []
        for k in range(2 + a.steps):
            yo = eng.pinned_out()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            t, _ = eng.ode_cfl3_step(t, 1e9, 0.8, big, y, yo)
            ts.append(time.perf_counter() - t0)
            y = yo
        chk = float(y.sum()), float(y.min()), float(y.max())
        if ref is None and not flag:
            ref = y.copy()
        row = {"planes": planes, "switch": flag, "ms_per_step_min": 1e3 * min(ts[2:]), "ms_per_step_mean": 1e3 * float(np.mean(ts[2:])),
               "bit_identical_to_first": bool(np.array_equal(ref, y)) if not flag else None, "checksum": chk}
        print(json.dumps(row), flush=True)
    eng.set_pipeline_planes(0)

    # the same through the reference-facing call (what bench.py's e2e times)
    opts = lsp.odeCFLset(lsp.Bundle(dict(factorCFL=0.8, singleStep="on")))
    y_np = np.ascontiguousarray(data0.reshape(-1, 1))
    te = 0.0
    ts = []
    for k in range(3 + a.steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        te, y_np, _ = lsp.odeCFL3(lsp.termLaxFriedrichs, [te, 1e9], y_np, opts, sd)
        ts.append(time.perf_counter() - t0)
    print(json.dumps({"api": "odeCFL3 default chunking", "ms_per_step_min": 1e3 * min(ts[3:]),
                      "ms_per_step_mean": 1e3 * float(np.mean(ts[3:]))}), flush=True)


if __name__ == "__main__":
    main()
