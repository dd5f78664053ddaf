"""One rank of a REAL multi-process slab solve (one process per GPU, torch.distributed / NCCL), launched by
tests/test_gpu_multirank.py through ``python -m torch.distributed.run``:  every rank advances its slab with
``SlabSolver`` (both halo transports: peer-memory pushes and NCCL send/recv; overlap protocols on and off), rank 0
also advances the single-domain problem on one context, and every slab is compared with the matching planes of it.
The slab path runs the same kernels on the same values, so the bar is bit-identity; the dt sequence must be equal.
Prints one JSON line per case on rank 0 and exits non-zero on any mismatch.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cases(lsp):
    rng = np.random.default_rng(11)

    def air3d(N, pd):
        g = lsp.createGrid(np.array([-6.0, -10.0, 0.0]), np.array([20.0, 10.0, 2 * np.pi * (1 - 1 / N[2])]), np.array(N),
                           pdDims=pd)
        x = np.meshgrid(*[np.asarray(v).reshape(-1) for v in g.vs], indexing="ij")
        d0 = np.sqrt(x[0] ** 2 + x[1] ** 2) - 5 + 0.4 * np.sin(x[2] + 0.2 * x[0]) + 0.05 * rng.standard_normal(g.shape)
        return g, lsp.DubinsVehicleRel(g, 5, 1), np.ascontiguousarray(d0)

    def dint4d(N):
        g = lsp.createGrid(-np.ones(4), np.ones(4), np.array(N))
        x = np.meshgrid(*[np.asarray(v).reshape(-1) for v in g.vs], indexing="ij")
        d0 = np.sqrt((x[0] - x[2]) ** 2 + (x[1] - x[3]) ** 2) - 0.2 + 0.02 * rng.standard_normal(g.shape)
        return g, lsp.ProductSystem(g, [lsp.DoubleIntegrator(g, 1.0), lsp.DoubleIntegrator(g, 0.6)]), np.ascontiguousarray(d0)

    def dubins6d(N, pd):
        lo = [-6, -10, 0, -6, -10, 0.]
        hi = [20, 10, 2 * np.pi * (1 - 1 / N[2]), 20, 10, 2 * np.pi * (1 - 1 / N[5])]
        g = lsp.createGrid(np.array(lo), np.array(hi), np.array(N), pdDims=pd)
        x = np.meshgrid(*[np.asarray(v).reshape(-1) for v in g.vs], indexing="ij")
        d0 = np.minimum(np.sqrt(x[0] ** 2 + x[1] ** 2) - 5, np.sqrt(x[3] ** 2 + x[4] ** 2) - 5) \
            + 0.3 * np.sin(x[2] + x[5]) + 0.02 * rng.standard_normal(g.shape)
        return g, lsp.ProductSystem(g, [lsp.DubinsVehicleRel(g, 5, 1), lsp.DubinsVehicleRel(g, 4, 1.2)]), np.ascontiguousarray(d0)

    yield "air3d_48x40x36", air3d([48, 40, 36], [2]), "as_shipped"
    yield "air3d_periodic0_48x40x36", air3d([48, 40, 36], [0, 2]), "as_shipped"
    yield "air3d_intended_30x26x34", air3d([30, 26, 34], [2]), "intended"
    yield "dint4d_16x9x18x34", dint4d([16, 9, 18, 34]), "as_shipped"
    yield "dubins6d_12x9x8x7x9x12", dubins6d([12, 9, 8, 7, 9, 12], [2, 5]), "as_shipped"


def main():
    import torch
    import torch.distributed as dist
    import levelsetpy_b200 as lsp
    from levelsetpy_b200 import _lib as L
    from levelsetpy_b200.integration import rk3_step_resident
    from levelsetpy_b200.slab import SlabSolver
    from levelsetpy_b200.term import prepare_scheme

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nsteps, failures = 3, 0
    for name, (g, system, d0), weno in cases(lsp):
        sd = lsp.Bundle(dict(grid=g, hamFunc=system.hamiltonian, partialFunc=system.dissipation, wenoMode=weno,
                             dissFunc=lsp.artificialDissipationGLF, CoStateCalc=lsp.upwindFirstWENO5a))
        ref, ref_dts = None, []
        if rank == 0:                                  # single-domain answer on one context
            eng, ad = prepare_scheme(sd)
            eng.upload(d0)
            t = 0.0
            for _ in range(nsteps):
                t, dt = rk3_step_resident(eng, ad, g, t, 1.0, 0.8, np.finfo(np.float64).max, L.COMP_MIN_OVER_TIME)
                ref_dts.append(dt)
            ref = eng.download(shape=g.shape)
        # peer + overlap: product systems push their halos from inside pass 2 (fused) or piece by piece through the
        # copy engines (fused off); whole 3-D systems advance the interior range under a whole-plane push either way
        for transport, overlap, fused in (("peer", True, True), ("peer", True, "hybrid"), ("peer", True, False),
                                          ("peer", False, False), ("p2p", True, False), ("p2p", False, False)):
            if True:
                solver = SlabSolver(sd, device=local, transport=transport, overlap=overlap, fused=fused)
                solver.upload(np.ascontiguousarray(d0[solver.lo:solver.hi]))
                t, dts = 0.0, []
                for _ in range(nsteps):
                    t, dt = solver.step(t, 1.0, 0.8, L.COMP_MIN_OVER_TIME)
                    dts.append(dt)
                mine = solver.download()
                parts = [None] * world
                dist.all_gather_object(parts, (solver.lo, solver.hi, mine, dts))
                if rank == 0:
                    err, same_dt = 0.0, True
                    for lo, hi, arr, d in parts:
                        err = max(err, float(np.max(np.abs(arr - ref[lo:hi]))))
                        same_dt = same_dt and list(d) == list(ref_dts)
                    rng_ = float(ref.max() - ref.min())
                    ok = bool(err == 0.0 and same_dt)
                    failures += 0 if ok else 1
                    print(json.dumps({"case": name, "weno": weno, "world": world, "transport": transport,
                                      "overlap": overlap, "fused": (fused if (solver.overlapped() and solver.fused()) else False),
                                      "pieces": len(solver.pieces() or [None]) if solver.overlapped() else 1,
                                      "protocol": "two_pass" if solver.two_pass() else (
                                          "ranged" if solver.ranged() else "exchange_first"),
                                      "max_abs_err": err, "max_rel_err": err / rng_, "bit_identical": err == 0.0,
                                      "dt_identical": same_dt, "steps": nsteps, "ok": ok}), flush=True)
                torch.cuda.synchronize()
                solver.close()
                dist.barrier()
    flag = torch.tensor([failures], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
