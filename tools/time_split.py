"""Developer tooling: per-kernel times of the dimension-split path (6-D pair: pass 1 / pass 2 of each RK stage), of the
4-D pair and of the Flock batch, for one build of the library.   python tools/time_split.py [--lib path.so] [--planes0 8]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=None)
    ap.add_argument("--planes0", type=int, default=8)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--what", default="6d,4d,fb")
    a = ap.parse_args()
    from levelsetpy_b200 import _lib as L
    if a.lib:
        L.SO_PATH = os.path.abspath(a.lib)
    import torch
    import levelsetpy_b200 as lsp
    import bench
    from levelsetpy_b200.term import prepare_scheme
    out = {"lib": os.path.basename(L.SO_PATH)}

    def ev_time(fn, reps):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    for kind, key, planes0 in (("dubins6d", "6d", a.planes0), ("dint4d", "4d", None)):
        if key not in a.what.split(","):
            continue
        g, system, fill = bench.product_setup(lsp, kind, {"dubins6d": 41, "dint4d": 161}[kind], planes0)
        sd = bench.scheme_for(lsp, g, "as_shipped", system)
        eng, ad = prepare_scheme(sd)
        bench.fill_resident(eng, g, fill)
        eng.set_system(ad.system_id, ad.block(), list(enumerate(ad.tables(g))))
        dt = 0.8 * eng.alpha_max()[1]
        pts = float(np.prod(np.asarray(g.N, dtype=np.float64)))
        rec = {"nodes": pts}
        tot = 0.0
        for stage in (1, 2, 3):
            for p in (1, 2):
                ms = ev_time(lambda: eng.stage(stage, 0.0, dt, None, L.COMP_MIN_OVER_TIME, False, which_pass=p), a.reps)
                rec["s%dp%d_ms" % (stage, p)] = round(ms, 3)
                tot += ms
        rec["step_ms"] = round(tot, 3)
        rec["Gpts_per_s"] = round(pts / tot * 1e-6, 2)
        out[key] = rec
        eng.close()
        lsp.clear_engine_cache() if hasattr(lsp, "clear_engine_cache") else None
        torch.cuda.empty_cache()
    if "fb" in a.what.split(","):
        sds, data = bench.flock_batch_setup(lsp, 64, 101)
        bs = lsp.BatchSolver(sds, device=0)
        bs.upload(np.stack(data))
        ms = ev_time(lambda: bs.step(1e9, 0.8, L.COMP_MIN_OVER_TIME), 5)
        out["fb"] = {"grids": 64, "step_ms": round(ms, 3), "Gpts_per_s": round(64 * 101.0 ** 3 / ms * 1e-6, 2)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
