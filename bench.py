#!/usr/bin/env python
"""bench.py -- fp64 WENO5 + Lax-Friedrichs + TVD-RK3 grid-point updates/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload air3d512|...]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

One "step" = one full TVD-RK3 step (3 fused RHS+stage kernels) of the hot path over the whole grid.
1 point-step = one grid node advanced by one RK3 step.  N=1 workload: air3D 512^3 (configs[1]).
N>1: the same 512^2 cross-section with 512 planes per GPU, slab-decomposed along dim 0 with a 3-plane halo
exchange per RK stage (weak scaling).  Prints ONE JSON line (rank 0).

The same run also measures BASELINE.json's other configs and reports them under "workloads" in that line:
configs[2] (4-D double-integrator pair 161^4) and configs[3] (6-D relative-Dubins pair 41^6) STRONG-scaled over the
N ranks -- ms per step, point-steps/s, fraction of the aggregate 64-B HBM roofline, halo bytes per step, compute-only
and exchange-only times -- each with a "verify" record (the slabs after a few steps against a single-domain answer:
max_rel_err, dt_identical; 41^6 also against per-plane checksums committed from the N=1 run), and at N=1 configs[4]
(the 256 x 101^3 Flock batch).  ``--blocks none`` skips them.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fp64 WENO5+LF grid-point updates/s (1 point-step = one node advanced one TVD-RK3 step)"
UNIT = "point-steps/s"
ALG_BYTES_PER_POINT_STEP = 64.0   # 16 + 24 + 24 B over the three fused stages


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- workloads
def air3d_setup(lsp, N0, N1, N2):
    """air3D (Dubins relative, reach-avoid cylinder): SURVEY.md 8d configs 1/2."""
    g = lsp.createGrid(np.array([-6.0, -10.0, 0.0]), np.array([20.0, 10.0, 2 * np.pi * (1 - 1 / N2)]),
                       np.array([N0, N1, N2]), pdDims=2, low_mem=True)
    data0 = np.ascontiguousarray(np.broadcast_to(np.sqrt(g.xs[0] ** 2 + g.xs[1] ** 2) - 5.0, g.shape))
    return g, data0


def scheme_for(lsp, g, weno, system=None):
    s = system if system is not None else lsp.DubinsVehicleRel(g, 5, 1)
    return lsp.Bundle(dict(grid=g, hamFunc=s.hamiltonian, partialFunc=s.dissipation, wenoMode=weno,
                           dissFunc=lsp.artificialDissipationGLF, CoStateCalc=lsp.upwindFirstWENO5a))


def product_setup(lsp, kind, n, planes0=None):
    """the 4-D double-integrator pair (161^4) and the 6-D relative-Dubins pair (41^6).
    Returns (grid, system, fill) where fill(view, lo, hi) writes the initial data of dim-0 planes lo..hi into a torch
    view shaped [hi-lo, N1, ..., N_{D-2}, pitch] ON THE DEVICE (a 41^6 field is 38 GB: it is never staged on the host)."""
    import torch
    n0 = planes0 if planes0 is not None else n
    if kind == "dint4d":
        g = lsp.createGrid(np.array([-1.0] * 4), np.array([1.0] * 4), np.array([n0, n, n, n]), low_mem=True)
        system = lsp.ProductSystem(g, [lsp.DoubleIntegrator(g, 1.0), lsp.DoubleIntegrator(g, 1.0)])

        def fill(view, lo, hi):
            v = [torch.as_tensor(g.vs[d].reshape(-1), device=view.device) for d in range(4)]
            x1 = v[0][lo:hi].reshape(-1, 1, 1, 1); v1 = v[1].reshape(1, -1, 1, 1)
            x2 = v[2].reshape(1, 1, -1, 1); v2 = v[3].reshape(1, 1, 1, -1)
            view[..., :n] = torch.sqrt((x1 - x2) ** 2 + (v1 - v2) ** 2) - 0.2
        return g, system, fill
    if kind == "dubins6d":
        lo3, hi3 = [-6.0, -10.0, 0.0], [20.0, 10.0, 2 * np.pi * (1 - 1 / n)]
        g = lsp.createGrid(np.array(lo3 + lo3), np.array(hi3 + hi3), np.array([n0, n, n, n, n, n]), pdDims=[2, 5],
                           low_mem=True)
        system = lsp.ProductSystem(g, [lsp.DubinsVehicleRel(g, 5, 1), lsp.DubinsVehicleRel(g, 5, 1)])

        def fill(view, lo, hi):
            v = [torch.as_tensor(g.vs[d].reshape(-1), device=view.device) for d in range(6)]
            a = torch.sqrt(v[0][lo:hi].reshape(-1, 1) ** 2 + v[1].reshape(1, -1) ** 2) - 5.0
            b = torch.sqrt(v[3].reshape(-1, 1) ** 2 + v[4].reshape(1, -1) ** 2) - 5.0
            view[..., :n] = torch.minimum(a.reshape(hi - lo, n, 1, 1, 1, 1), b.reshape(1, 1, 1, n, n, 1))
        return g, system, fill
    raise SystemExit("unknown workload %r" % kind)


def flock_batch_setup(lsp, nb, n):
    """nb independent n^3 grids (flockGrid-style boxes shifted by 0.2 j), a 4-bird Flock each."""
    sds, data = [], []
    for j in range(nb):
        sh = 0.2 * j
        g = lsp.createGrid(np.array([-1 + sh, -1 + sh, -np.pi]), np.array([1 + sh, 1 + sh, np.pi * (1 - 2 / n)]),
                           np.array([n, n, n]), pdDims=2, low_mem=True)
        birds = [lsp.Bird(g, 1.0, 1.0, init_xyw=np.array([[0.1 * j + 0.01 * k], [0.2 * j - 0.02 * k], [0.3 * j + 0.1 * k]]),
                          label=k, neigh_rad=3) for k in range(4)]
        f = lsp.Flock(g, birds)
        sds.append(lsp.Bundle(dict(grid=g, hamFunc=f.hamiltonian, partialFunc=f.dissipation)))
        x = g.vs[0].reshape(-1, 1, 1) - sh
        y = g.vs[1].reshape(1, -1, 1) - sh
        data.append(np.broadcast_to(np.sqrt(x ** 2 + y ** 2) - 0.3, (n, n, n)))
    return sds, data


def fill_resident(eng, g, fill, planes_per_chunk=None, slab=None):
    """Initial data straight into RK buffer 0 (pitched layout), a few dim-0 planes at a time.  ``slab`` = (lo, hi):
    the context holds only those dim-0 planes of the grid (plus its stored halo planes)."""
    N = [int(x) for x in np.asarray(g.N).reshape(-1)]
    lo0, hi0 = slab if slab is not None else (0, N[0])
    n0 = hi0 - lo0
    pitch = (N[-1] + 1) // 2 * 2
    buf = eng.buffer_tensor(0)
    halo = (buf.numel() - n0 * int(np.prod(N[1:-1])) * pitch) // 2
    body = buf[halo:buf.numel() - halo].view(n0, *N[1:-1], pitch)
    per_plane = int(np.prod(N[1:-1])) * pitch * 8
    step = planes_per_chunk or max(1, int(2e9 // per_plane))
    for lo in range(0, n0, step):
        hi = min(n0, lo + step)
        fill(body[lo:hi], lo0 + lo, lo0 + hi)


# ----------------------------------------------------------------------------- configs[2..4] in the same run
GOLDEN_6D = os.path.join(ROOT, "tests", "golden", "bench_dubins6d_41_checksums.json")
VERIFY_STEPS = 2


def _interior(eng, g, n0):
    """Torch view [n0, N1, ..., N_{D-2}, pitch] of the interior planes of RK buffer 0."""
    N = [int(x) for x in np.asarray(g.N).reshape(-1)]
    pitch = (N[-1] + 1) // 2 * 2
    buf = eng.buffer_tensor(0)
    halo = (buf.numel() - n0 * int(np.prod(N[1:-1])) * pitch) // 2
    return buf[halo:buf.numel() - halo].view(n0, *N[1:-1], pitch)


def _plane_checksums(view, nx):
    """Per dim-0 plane (sum, min, max) of the nodes (pad columns excluded)."""
    out = []
    for p in range(view.shape[0]):
        v = view[p][..., :nx]
        out.append([float(v.sum(dtype=v.dtype).item()), float(v.min().item()), float(v.max().item())])
    return out


def _timed(step, steps, barrier, world, torch):
    """K steps between barriers, device-timed, max over ranks (ms per step)."""
    t = 0.0
    torch.cuda.synchronize()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        t = step(t)
    ev1.record()
    torch.cuda.synchronize()
    barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        import torch.distributed as dist
        tms = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    return ms / steps


def product_block(args, lsp, L, kind, world, rank, local, hbm_peak, comp, backend):
    """configs[2] / [3]: the fixed product grid, one context (N=1) or slab-decomposed along dim 0 over the ranks
    (strong scaling), state made on the device; timing, attribution and verification.  Returns a dict (rank 0)."""
    import torch
    from levelsetpy_b200.engine import Engine, clear_engine_cache
    from levelsetpy_b200.integration import rk3_step_resident
    from levelsetpy_b200.term import prepare_scheme
    n = {"dint4d": 161, "dubins6d": 41}[kind] if args.block_n is None else args.block_n
    steps = max(1, min(args.steps, args.block_steps))
    gg, system, fill = product_setup(lsp, kind, n)
    sd = scheme_for(lsp, gg, args.weno, system)
    N = [int(x) for x in np.asarray(gg.N).reshape(-1)]
    points = float(np.prod(np.asarray(N, dtype=np.float64)))
    fmax = np.finfo(np.float64).max
    out = {"config": ("configs[2]: 4-D double-integrator pair %d^4" if kind == "dint4d" else
                      "configs[3]: 6-D relative-Dubins pair %d^6") % n, "scaling": "strong", "n_gpus": world,
           "steps": steps, "warmup": 3}
    solver = None
    if world > 1:
        import torch.distributed as dist
        from levelsetpy_b200.slab import SlabSolver
        solver = SlabSolver(sd, device=local, backend=backend, transport=args.transport, fused={"auto": "auto", "both": True, "hybrid": "hybrid", "off": False}[args.fused])
        eng, lo, hi = solver.eng, solver.lo, solver.hi
        step = lambda t: solver.step(t, 1e9, 0.8, comp)[0]
        barrier = dist.barrier
        refill = lambda: fill_resident(eng, gg, fill, slab=(lo, hi))
        solver.state_changed()
    else:
        eng, ad = prepare_scheme(sd)
        eng.set_backend(backend)
        lo, hi = 0, N[0]
        step = lambda t: rk3_step_resident(eng, ad, gg, t, 1e9, 0.8, fmax, comp)[0]
        barrier = lambda: None
        refill = lambda: fill_resident(eng, gg, fill)
    refill()
    t = 0.0
    for _ in range(3):
        t = step(t)
    ms = _timed(step, steps, barrier, world, torch)
    out.update({"ms_per_step": ms, "value": points / (ms * 1e-3), "unit": UNIT,
                "frac_of_aggregate_64B_roofline": points * ALG_BYTES_PER_POINT_STEP / (ms * 1e-3) / (world * hbm_peak * 1e9)})
    if solver is not None:
        plane_bytes = eng.plane_elems * 8
        faces = (1 if solver.lo_peer is not None else 0) + (1 if solver.hi_peer is not None else 0)
        out.update({"planes_per_rank": [hi - lo], "transport": ("peer memory (%s)" % (("stores from inside pass 2" + (" towards one neighbour, copy engines towards "
                                                                                      "the other" if solver.hybrid() else ""))
                                                        if (solver.overlapped() and solver.fused()) else "copy engines"))
                    if solver.peer else "NCCL send/recv",
                    "protocol": "two_pass" if solver.two_pass() else ("ranged" if solver.ranged() else "exchange_first"),
                    "halo_bytes_in_per_step_per_rank": 3 * faces * L.HJ_GHOST * plane_bytes})
        # attribution: the same step with the exchange switched off, and the exchange alone (results are discarded:
        # the state is re-made before verification)
        out["fused_halo_push"] = ("hybrid" if solver.hybrid() else "both") if (solver.overlapped() and solver.fused()) else "off"
        out["pieces"] = len(solver.pieces() or [None])
        solver.set_mode("compute")
        out["compute_only_ms"] = _timed(step, steps, barrier, world, torch)
        solver.set_mode("comm")
        out["exchange_only_ms"] = _timed(step, steps, barrier, world, torch)
        solver.set_mode("full")

    if rank == 0:
        print("[bench] %s timing: %s" % (kind, json.dumps(out)), file=sys.stderr, flush=True)
    # ---- verify: VERIFY_STEPS steps from the initial data against a single-domain answer
    refill()
    if solver is not None:
        solver.state_changed()            # the state was rewritten in place: halo pieces pushed ahead are stale
    t, dts = 0.0, []
    for _ in range(VERIFY_STEPS):
        t_new = step(t)
        dts.append(t_new - t)
        t = t_new
    torch.cuda.synchronize()
    mine = _interior(eng, gg, hi - lo)
    sums = _plane_checksums(mine, N[-1])
    ver = {"steps": VERIFY_STEPS}
    if world > 1:
        import torch.distributed as dist
        allsums = [None] * world
        dist.all_gather_object(allsums, (lo, sums, dts))
        allsums.sort(key=lambda x: x[0])
        sums = [p for _, part, _ in allsums for p in part]
        ver["dt_identical_across_ranks"] = all(d == allsums[0][2] for _, _, d in allsums)
    if kind == "dubins6d" and n == 41:
        if args.write_golden and world == 1 and rank == 0:
            with open(args.write_golden, "w") as fh:
                json.dump({"what": "per dim-0 plane (sum, min, max) of the 41^6 field after %d steps (bench.py N=1)" % VERIFY_STEPS,
                           "dts": dts, "planes": sums}, fh)
        if os.path.exists(GOLDEN_6D):
            with open(GOLDEN_6D) as fh:
                gold = json.load(fh)
            gp = np.asarray(gold["planes"])
            sp = np.asarray(sums)
            scale = np.maximum(np.abs(gp[:, 0]), 1.0)
            ver["checksum_max_rel_diff_vs_n1_golden"] = float(np.max(np.abs(sp[:, 0] - gp[:, 0]) / scale))
            ver["plane_minmax_identical_to_n1_golden"] = bool(np.array_equal(sp[:, 1:], gp[:, 1:]))
            ver["dt_identical"] = list(dts) == list(gold["dts"])
        else:
            ver["checksums"] = "no committed N=1 checksums (bench.py --write-golden at N=1)"
    # element-wise against a single-domain run on rank 0: 161^4 on the gather backend (an independent kernel),
    # 41^6 (N > 1 only: two 114 GB contexts do not fit one GPU) on the dimension-split path of one context
    do_elem = kind == "dint4d" or world > 1
    if do_elem:                                         # ... if rank 0 has room for the single-domain context
        need = 3.0 * points * 8 * ((N[-1] + 1) // 2 * 2) / N[-1] + 4e9
        room = torch.tensor([1 if torch.cuda.mem_get_info()[0] > need else 0], device="cuda")
        if world > 1:
            import torch.distributed as dist
            dist.broadcast(room, 0)
        do_elem = bool(room.item())
        if not do_elem:
            ver["elementwise"] = "skipped: no room on rank 0 for a %.0f GB single-domain context next to its slab" % (need * 1e-9)
    if do_elem:
        ref = None
        if rank == 0:
            ref_backend = L.BACKEND_GATHER if kind == "dint4d" else backend
            ref = Engine(gg, args.weno, local, backend=ref_backend)
            ad0 = solver.adapter if solver is not None else ad
            fill_resident(ref, gg, fill)
            tr, rdts = 0.0, []
            for _ in range(VERIFY_STEPS):
                tn = rk3_step_resident(ref, ad0, gg, tr, 1e9, 0.8, fmax, comp)[0]
                rdts.append(tn - tr)
                tr = tn
            torch.cuda.synchronize()
            refv = _interior(ref, gg, N[0])
            ver["reference"] = "single-domain, %s backend, rank 0" % ("gather" if ref_backend == L.BACKEND_GATHER else "plane-ring")
            ver["dt_identical"] = ver.get("dt_identical", True) and list(rdts) == list(dts)
        err = torch.zeros(1, dtype=torch.float64, device="cuda")
        lo_v, hi_v = float("inf"), -float("inf")
        nx = N[-1]

        def compare(got, p):                            # one dim-0 plane at a time: no field-sized temporaries
            nonlocal err, lo_v, hi_v
            want = refv[p][..., :nx]
            err = torch.maximum(err, (got[..., :nx] - want).abs().max().reshape(1))
            lo_v, hi_v = min(lo_v, float(want.min().item())), max(hi_v, float(want.max().item()))

        if world == 1:
            for p in range(N[0]):
                compare(mine[p], p)
        else:
            import torch.distributed as dist
            parts = partition_of(N[0], world)
            stage_t = torch.empty_like(mine[0]) if rank == 0 else None
            for r in range(world):                      # one slab at a time through a plane-sized staging tensor
                rlo, rhi = parts[r]
                for p in range(rlo, rhi):
                    if rank == 0:
                        if r == 0:
                            compare(mine[p - rlo], p)
                        else:
                            dist.recv(stage_t, src=r)
                            compare(stage_t, p)
                    elif rank == r:
                        dist.send(mine[p - rlo].contiguous(), dst=0)
        rng_ = (hi_v - lo_v) if hi_v > lo_v else 1.0
        if rank == 0:
            ver["max_abs_err"] = float(err.item())
            ver["max_rel_err"] = float(err.item()) / rng_
            ver["bit_identical"] = float(err.item()) == 0.0
            ref.close()
    out["verify"] = ver
    if rank == 0:
        print("[bench] %s: %s" % (kind, json.dumps(out)), file=sys.stderr, flush=True)
    if solver is not None:
        torch.cuda.synchronize()
        solver.close()
    else:
        clear_engine_cache()
        eng.close()
    torch.cuda.empty_cache()
    return out


def partition_of(n0, world):
    from levelsetpy_b200.slab import partition
    return partition(n0, world)


def flock_block(args, lsp, L, local, hbm_peak, comp):
    """configs[4] at N=1: 256 independent 101^3 Flock grids as one batch context."""
    import torch
    steps = max(1, min(args.steps, args.block_steps))
    sds, data = flock_batch_setup(lsp, args.batch, 101)
    bsolver = lsp.BatchSolver(sds, device=local)
    bsolver.upload(np.stack(data))
    step = lambda t: float(bsolver.step(1e9, 0.8, comp)[0][0])
    t = 0.0
    for _ in range(3):
        t = step(t)
    ms = _timed(step, steps, lambda: None, 1, torch)
    points = float(args.batch) * 101.0 ** 3
    del bsolver
    torch.cuda.empty_cache()
    return {"config": "configs[4]: batch of %d independent 101^3 Flock grids (4 birds each, per-grid dt)" % args.batch,
            "n_gpus": 1, "steps": steps, "warmup": 3, "ms_per_step": ms, "value": points / (ms * 1e-3), "unit": UNIT,
            "frac_of_aggregate_64B_roofline": points * ALG_BYTES_PER_POINT_STEP / (ms * 1e-3) / (hbm_peak * 1e9),
            "scaling": "replicas only (independent grids, no exchange)"}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arm (oracle port)
def cpu_port_step_rate(sample_n, steps, warmup=0):
    """The numpy oracle (restating the reference's CPU path) on an air3D sample; returns (point-steps/s, secs)."""
    from oracle import hj_oracle as orc
    from oracle import systems as osys
    import levelsetpy_b200.grids as grids
    g = grids.createGrid(np.array([-6.0, -10.0, 0.0]), np.array([20.0, 10.0, 2 * np.pi * (1 - 1 / sample_n)]),
                         np.array([sample_n] * 3), pdDims=2)
    data0 = np.sqrt(g.xs[0] ** 2 + g.xs[1] ** 2) - 5.0
    o = osys.DubinsVehicleRel(g, 5, 1)
    sd = orc.OracleSchemeData(grid=g, hamFunc=o.hamiltonian, partialFunc=o.dissipation)
    y = data0.reshape(-1, 1)
    t = 0.0
    for _ in range(warmup):
        t, y, _ = orc.ode_cfl3([t, 1e9], y, sd, factor_cfl=0.8, single_step=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        y_last = y
        t, y, _ = orc.ode_cfl3([t, 1e9], y, sd, factor_cfl=0.8, single_step=True)
        y = np.minimum(y, y_last)
    dt = time.perf_counter() - t0
    return sample_n ** 3 * steps / dt, dt


def literal_reference_record():
    """The LITERAL reference's timing on configs[0], measured where its checkout exists (tools/time_reference_literal.py,
    build container, through oracle/ref_shim.py) and committed: it cannot travel to the GPU box, so the CPU arm times
    the oracle port live and quotes this next to it (the port is ~3x faster than the literal code, same bits)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_reference_literal_cpu.json")) as fh:
            d = json.load(fh)
        return {k: d[k] for k in ("what", "literal_reference_s_per_step", "literal_reference_point_steps_per_s",
                                  "oracle_port_s_per_step", "port_over_literal", "identical_result", "where")}
    except Exception:
        return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cpu_sample
    rate, secs = cpu_port_step_rate(n, args.steps, args.warmup)
    sample = "air3D %d^3 (same box/ICs/system as the GPU arm), one TVD-RK3 step per bench step, numpy oracle port" % n
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "cpu_sample": sample},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                         "host_cores": os.cpu_count(), "literal_reference": literal_reference_record()},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_name(args):
    n = args.n
    if args.workload == "dint4d":
        return ("4-D double-integrator pair %d^4 fp64 (GLF, WENO5a, odeCFL3, minVOverTime), slab-decomposed along dim 0 "
                "over %d GPU(s)" % (n, args.gpus))
    if args.workload == "flockbatch":
        return ("batch of %d independent %d^3 Flock grids (4 birds each, per-grid dt), one launch per RK stage for the "
                "whole batch" % (args.batch, n))
    if args.workload == "dubins6d":
        return ("6-D relative-Dubins pair %s fp64 (GLF, WENO5a, odeCFL3, minVOverTime), slab-decomposed along dim 0 "
                "over %d GPU(s)" % ("x".join([str(args.planes0 or n)] + [str(n)] * 5), args.gpus))
    if args.gpus == 1:
        return "air3D %d^3 fp64 (Dubins relative, GLF, WENO5a, odeCFL3, minVOverTime), 1xB200" % n
    return ("air3D %dx%dx%d fp64 slab-decomposed along dim 0 over %d GPUs (%d planes/GPU, 3-plane halo exchange "
            "per RK stage)" % (n * args.gpus, n, n, args.gpus, n))


# ----------------------------------------------------------------------------- GPU arm
def committed_traffic():
    """dram bytes per stage launch from the committed ncu --set full capture (profiles/r02_traffic.json, made from
    profiles/r02_ncu_stage_summary.txt), or None."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as fh:
                return float(json.load(fh)["avg_bytes_per_launch"])
        except Exception:
            continue
    return None


def run_ours(args):
    import torch
    import levelsetpy_b200 as lsp
    from levelsetpy_b200 import _lib as L
    from levelsetpy_b200.integration import rk3_step_resident
    from levelsetpy_b200.term import prepare_scheme

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    lib = L.load()
    n = args.n
    hbm_peak, peak_src = peaks()
    comp = L.COMP_MIN_OVER_TIME
    backend = {"auto": L.BACKEND_AUTO, "gather": L.BACKEND_GATHER, "tma": L.BACKEND_TMA}[args.backend]

    scaling = "weak"
    slab0 = data0 = None
    if world > 1 and args.workload in ("dint4d", "dubins6d"):
        # configs[2] / [3]: the FIXED product grid slab-decomposed along dim 0 (strong scaling), state made on the device
        import torch.distributed as dist
        from levelsetpy_b200.slab import SlabSolver
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        gg, system, fill = product_setup(lsp, args.workload, n, args.planes0)
        solver = SlabSolver(scheme_for(lsp, gg, args.weno, system), device=local, backend=backend, transport=args.transport)
        fill_resident(solver.eng, gg, fill, slab=(solver.lo, solver.hi))
        data0 = None
        step = lambda t: solver.step(t, 1e9, 0.8, comp)[0]
        points = float(np.prod(np.asarray(gg.N, dtype=np.float64)))
        barrier = lambda: dist.barrier()
        scaling = "strong"
    elif world > 1:
        import torch.distributed as dist
        from levelsetpy_b200.slab import SlabSolver
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # global grid: world*n planes along dim 0; each rank builds only its slab of the initial data
        gg = lsp.createGrid(np.array([-6.0, -10.0, 0.0]), np.array([20.0, 10.0, 2 * np.pi * (1 - 1 / n)]),
                            np.array([n * world, n, n]), pdDims=2, low_mem=True)
        solver = SlabSolver(scheme_for(lsp, gg, args.weno), device=local, backend=backend, transport=args.transport)
        lo, hi = solver.lo, solver.hi
        x0 = gg.vs[0].reshape(-1)[lo:hi].reshape(-1, 1, 1)
        slab0 = np.ascontiguousarray(np.broadcast_to(np.sqrt(x0 ** 2 + gg.xs[1] ** 2) - 5.0, (hi - lo, n, n)))
        solver.upload(slab0)
        step = lambda t: solver.step(t, 1e9, 0.8, comp)[0]
        points = float(n) ** 3 * world
        barrier = lambda: dist.barrier()
    elif args.workload == "flockbatch":
        sds, data = flock_batch_setup(lsp, args.batch, n)
        bsolver = lsp.BatchSolver(sds, device=local)
        bsolver.upload(np.stack(data))
        data0 = None
        step = lambda t: float(bsolver.step(1e9, 0.8, comp)[0][0])
        points = float(args.batch) * float(n) ** 3
        barrier = lambda: None
    elif args.workload != "air3d":
        g, system, fill = product_setup(lsp, args.workload, n, args.planes0)
        sd = scheme_for(lsp, g, args.weno, system)
        eng, ad = prepare_scheme(sd)
        eng.set_backend(backend)
        fill_resident(eng, g, fill)
        data0 = None
        step = lambda t: rk3_step_resident(eng, ad, g, t, 1e9, 0.8, np.finfo(np.float64).max, comp)[0]
        points = float(np.prod(np.asarray(g.N, dtype=np.float64)))
        barrier = lambda: None
    else:
        g, data0 = air3d_setup(lsp, n, n, n)
        sd = scheme_for(lsp, g, args.weno)
        eng, ad = prepare_scheme(sd)
        eng.set_backend(backend)
        eng.upload(data0)
        step = lambda t: rk3_step_resident(eng, ad, g, t, 1e9, 0.8, np.finfo(np.float64).max, comp)[0]
        points = float(n) ** 3
        barrier = lambda: None

    t = 0.0
    for _ in range(args.warmup):
        t = step(t)
    torch.cuda.synchronize()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = lib.hj_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    barrier()
    ev0.record()
    for _ in range(args.steps):
        t = step(t)
    ev1.record()
    torch.cuda.synchronize()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = lib.hj_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        import torch.distributed as dist
        tms = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    secs = ms * 1e-3
    value = points * args.steps / secs

    # ---- e2e: the reference-facing call with HOST buffers: H2D of y, one RK3 step, D2H of y, per step
    e2e = None
    if args.e2e_steps <= 0:
        pass
    elif data0 is None and slab0 is None:
        pass                                   # product workloads: resident-state numbers only (no host copy of a 38 GB field)
    elif world == 1:
        # the call a user of the reference makes (hji_solver.py:542): odeCFL3(termLaxFriedrichs, [t, tau], y, options, sd)
        # with a HOST numpy y and singleStep='on'; y comes back as a host array.  H2D + 3 stage kernels + D2H per call.
        opts = lsp.odeCFLset(lsp.Bundle(dict(factorCFL=0.8, singleStep="on")))
        y_np = np.ascontiguousarray(data0.reshape(-1, 1))
        te = 0.0
        for _ in range(3):                              # also brings up the engine's three pinned output arrays
            te, y_np, _ = lsp.odeCFL3(lsp.termLaxFriedrichs, [te, 1e9], y_np, opts, sd)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            te, y_np, _ = lsp.odeCFL3(lsp.termLaxFriedrichs, [te, 1e9], y_np, opts, sd)
        torch.cuda.synchronize()
        es = time.perf_counter() - t0
        e2e = {"value": points * args.e2e_steps / es, "unit": UNIT, "h2d_bytes_per_step": int(points * 8),
               "d2h_bytes_per_step": int(points * 8), "steps": args.e2e_steps, "ms_per_step": 1e3 * es / args.e2e_steps,
               "api": "levelsetpy_b200.odeCFL3(termLaxFriedrichs, [t, tau], y_host, odeCFLset(singleStep='on'), schemeData) "
                      "-> hj_ode_cfl3_step (host y in, host y out; no min-over-time epilogue: that is the driver's line)"}
    else:
        import torch.distributed as dist
        # slab job: every rank uploads its slab from pinned host memory, steps once, downloads it
        y_host = torch.from_numpy(slab0.reshape(-1)).pin_memory()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        te = 0.0
        for _ in range(args.e2e_steps):
            solver.upload(y_host.numpy())
            te = solver.step(te, 1e9, 0.8, comp)[0]
            solver.download(out=y_host.numpy())
        torch.cuda.synchronize()
        dist.barrier()
        es = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        dist.all_reduce(es, op=dist.ReduceOp.MAX)
        es = float(es.item())
        e2e = {"value": points * args.e2e_steps / es, "unit": UNIT, "h2d_bytes_per_step": int(points * 8),
               "d2h_bytes_per_step": int(points * 8), "steps": args.e2e_steps, "ms_per_step": 1e3 * es / args.e2e_steps,
               "api": "SlabSolver.upload/step/download (pinned host slabs)"}
        # the host-link ceiling of that pattern: every rank moves its slab up and down at once, nothing else running
        dev = torch.empty_like(y_host, device="cuda")
        back = torch.empty_like(y_host).pin_memory()
        up, down = torch.cuda.Stream(), torch.cuda.Stream()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            with torch.cuda.stream(up):
                dev.copy_(y_host, non_blocking=True)
            with torch.cuda.stream(down):
                back.copy_(dev, non_blocking=True)
        torch.cuda.synchronize()
        dist.barrier()
        cs = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        dist.all_reduce(cs, op=dist.ReduceOp.MAX)
        e2e["host_link_floor_ms_per_step"] = 1e3 * float(cs.item()) / 2
        e2e["host_link_note"] = ("H2D + D2H of one slab per rank, all ranks at once, both directions overlapped "
                                 "(cudaMemcpyAsync on two streams, pinned): the floor of any per-step host round trip")
        del dev, back

    # ---- BASELINE.json configs[2..4] in the same run (strong-scaled over the same ranks), each with a verify record
    workloads = None
    if args.blocks != "none" and args.workload == "air3d" and args.weno == "as_shipped":
        from levelsetpy_b200.engine import clear_engine_cache
        if world > 1:
            solver.close()
            del solver
        else:
            clear_engine_cache()
            eng.close()
        torch.cuda.empty_cache()
        workloads = {}
        for kind, key in (("dubins6d", "dubins6d_41^6"), ("dint4d", "dint4d_161^4")):
            if args.blocks not in ("all", kind):
                continue
            try:
                workloads[key] = product_block(args, lsp, L, kind, world, rank, local, hbm_peak, comp, backend)
            except Exception as e:                        # the headline line must survive a failing block
                workloads[key] = {"error": "%s: %s" % (type(e).__name__, e)}
                if world > 1:
                    raise
        if world == 1 and args.blocks in ("all", "flockbatch"):
            try:
                workloads["flockbatch_256x101^3"] = flock_block(args, lsp, L, local, hbm_peak, comp)
            except Exception as e:
                workloads["flockbatch_256x101^3"] = {"error": "%s: %s" % (type(e).__name__, e)}

    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    stage_launches = 3 * args.steps
    avg_launch_s = secs / stage_launches                 # the three stage kernels are the whole step on this path
    achieved = (points / world) * (ALG_BYTES_PER_POINT_STEP / 3.0) / avg_launch_s / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "weno": args.weno, "backend": args.backend,
                   "factorCFL": 0.8, "compMethod": "minVOverTime",
                   "l2": "inputs larger than L2 (3 x %.2f GB fields per GPU)" % (points / world * 8e-9),
                   "point_stage_updates_per_s": 3 * value},
        "clocks": clocks,
        "e2e": e2e,
        "workloads": workloads,
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": args.traffic if args.traffic is not None else (
                         committed_traffic() if (args.workload == "air3d" and world == 1 and n == 512) else None),
                     "peak_source": peak_src, "kernel": "k_stage_* (fused RHS + RK stage)",
                     "algorithmic_bytes_per_launch": (points / world) * ALG_BYTES_PER_POINT_STEP / 3.0,
                     "avg_launch_ms": 1e3 * avg_launch_s},
    }
    if world == 1 and not args.no_cpu:
        rate, csecs = cpu_port_step_rate(args.cpu_sample, args.cpu_steps)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port", "host_cores": os.cpu_count(),
                                "literal_reference": literal_reference_record(),
                                "sample": "air3D %d^3, %d TVD-RK3 steps of the numpy oracle port (%.1f s)" % (
                                    args.cpu_sample, args.cpu_steps, csecs)}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=None, help="nodes per dim (per GPU along dim 0); default per workload")
    ap.add_argument("--workload", default="air3d", choices=["air3d", "dint4d", "dubins6d", "flockbatch"],
                    help="air3d = configs[1] (the bench line); dint4d / dubins6d = configs[2] / [3] (1 GPU, resident state)")
    ap.add_argument("--planes0", type=int, default=None, help="dubins6d: dim-0 extent (default n)")
    ap.add_argument("--batch", type=int, default=256, help="flockbatch: number of grids")
    ap.add_argument("--weno", default="as_shipped", choices=["as_shipped", "intended"])
    ap.add_argument("--backend", default="auto", choices=["auto", "gather", "tma"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--transport", default="auto", choices=["auto", "peer", "p2p"],
                    help="slab halos: peer-memory pushes on the copy engines (default) or NCCL send/recv")
    ap.add_argument("--fused", default="auto", choices=["auto", "both", "hybrid", "off"],
                    help="product systems: halo planes stored into the neighbours from inside pass 2 (both sides), one "
                         "side that way and the other through the copy engines (hybrid), or all through the copy engines "
                         "piece by piece (off)")
    ap.add_argument("--blocks", default="all", choices=["all", "none", "dubins6d", "dint4d", "flockbatch"],
                    help="also measure + verify configs[2..4] in the default run ('workloads' in the JSON line)")
    ap.add_argument("--block-steps", type=int, default=5, help="timed steps of each workloads block (<= --steps)")
    ap.add_argument("--block-n", type=int, default=None, help="developer: nodes per dim of the product blocks")
    ap.add_argument("--write-golden", default=None, metavar="PATH",
                    help="N=1: write the per-plane checksums of the 41^6 verify run to PATH (committed as "
                         "tests/golden/bench_dubins6d_41_checksums.json)")
    ap.add_argument("--cpu-sample", type=int, default=101, help="CPU arm: air3D n^3 sample (101 = configs[0])")
    ap.add_argument("--cpu-steps", type=int, default=10)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--traffic", type=float, default=None, help="dram bytes/launch from an ncu --set full capture")
    args = ap.parse_args()
    if args.n is None:
        args.n = {"air3d": 512, "dint4d": 161, "dubins6d": 41, "flockbatch": 101}[args.workload]
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
