"""Batched small grids: many independent 3-D grids, one Flock each, advanced together.

The reference handles a flock of flocks with a Python loop -- one ``odeCFL3(termLaxFriedrichs, ...)`` call per grid
(ode_cfl_3.py:11), each on an L2-sized 101^3 field that is launch/latency bound on a GPU.  ``BatchSolver`` keeps all
grids of identical shape in ONE resident field ``[nbatch, N0, N1, N2]`` and advances the whole batch with one fused
kernel per RK stage (C-ABI: hj_create_batch / hj_step_batch).  Each grid keeps its own Flock parameter block (re-derived
on each of the three RHS evaluations like flock.py:213) and its own CFL time step, exactly as the per-grid loop would.
"""
import ctypes as C
import weakref

import numpy as np

from . import _lib as L
from .engine import current_stream, grid_signature, weno_mode_of
from .functors import resolve
from .integration import rk3_times

__all__ = ["BatchSolver", "batch_step_plan"]


def batch_step_plan(adapters, dxs, t, t_end, factorCFL, maxStep=np.finfo(np.float64).max):
    """Host side of one batched TVD-RK3 step: per grid, the three per-stage parameter blocks, the CFL time step
    (ode_cfl_3.py:142-143, with that grid's own dx) and the new time.  ``dxs``/``t``/``t_end`` are per-grid
    sequences.  Pure numpy (no device)."""
    nb = len(adapters)
    blocks = [[None] * nb for _ in range(3)]
    dts, t_new = np.empty(nb), np.empty(nb)
    for j, ad in enumerate(adapters):
        bs = [ad.block(), ad.block(), ad.block()]          # hamFunc re-runs the flock bookkeeping on every RHS
        al = ad.alphas(bs[0])
        inv = 0
        dx = dxs[j]
        for d in range(len(dx)):
            inv = inv + (al[d] / dx[d])                    # artificial_diss_glf.py:107, dims in order
        step_bound = 1 / inv
        dt = float(np.min(np.hstack((factorCFL * step_bound, t_end[j] - t[j], maxStep))))
        dts[j] = dt
        t_new[j] = rk3_times(t[j], dt)[2]
        for k in range(3):
            blocks[k][j] = bs[k]
    npar = max(b.size for stage in blocks for b in stage)
    params = np.zeros((3, nb, npar))
    for k in range(3):
        for j in range(nb):
            params[k, j, :blocks[k][j].size] = blocks[k][j]
    return params, dts, t_new


class BatchSolver:
    """``BatchSolver([schemeData_0, ..., schemeData_{B-1}])``: every schemeData is what the reference would hand to
    odeCFL3 for one grid (``grid``, ``hamFunc`` / ``partialFunc`` of that grid's Flock).  All grids must share shape,
    cell sizes and boundary kinds (flockGrid boxes shifted by a constant do)."""

    def __init__(self, scheme_datas, device=0):
        if not scheme_datas:
            raise ValueError("empty batch")
        self.lib = L.load()
        self.sds = list(scheme_datas)
        sig0 = grid_signature(self.sds[0].grid)
        D, N, dx, kinds, tz, _ = sig0
        # every grid keeps its own dx for its CFL step (host); the stencil coefficients on the device use grid 0's,
        # so the cell sizes may differ by rounding only (boxes shifted by a constant)
        self.dxs = [np.asarray(dx, dtype=np.float64)]
        for sd in self.sds[1:]:
            Dj, Nj, dxj, kj, tzj, _ = grid_signature(sd.grid)
            if (Dj, list(Nj), list(kj), list(tzj)) != (D, list(N), list(kinds), list(tz)) or \
                    not np.allclose(dxj, dx, rtol=1e-12, atol=0.0):
                raise ValueError("all grids of a batch must share shape, cell sizes and boundary kinds")
            self.dxs.append(np.asarray(dxj, dtype=np.float64))
        if weno_mode_of(self.sds[0]) != "as_shipped":
            raise NotImplementedError("batch contexts run the as-shipped scheme (the intended WENO epsilon is per grid)")
        self.adapters = [resolve(sd.hamFunc, sd.partialFunc, sd.grid) for sd in self.sds]
        if any(ad.system_id != L.SYS_FLOCK for ad in self.adapters):
            raise NotImplementedError("batched stepping is compiled for Flock / Bird systems")
        self.nb, self.D, self.N, self.dx, self.device = len(self.sds), D, list(N), dx, device
        self.shape = (self.nb,) + tuple(self.N)
        h = C.c_void_p()
        L.check(self.lib.hj_create_batch(C.byref(h), device, self.nb, D, (C.c_int64 * D)(*N), (C.c_double * D)(*dx),
                                         (C.c_int * D)(*kinds), (C.c_int * D)(*tz), L.WENO_AS_SHIPPED))
        self.h = h
        self._fin = weakref.finalize(self, self.lib.hj_destroy, h)
        vs = sig0[5]
        for d in range(D):
            L.check(self.lib.hj_set_axis(self.h, d, vs[d].ctypes.data, vs[d].size))
        self.t = np.zeros(self.nb)
        self._npar = None

    def upload(self, data, field=L.FIELD_STATE):
        """``data``: array [nbatch, N0, N1, N2] (or a list of per-grid arrays)."""
        a = np.ascontiguousarray(np.stack([np.asarray(x).reshape(self.N) for x in data]) if isinstance(data, (list, tuple))
                                 else np.asarray(data, dtype=np.float64).reshape(self.shape), dtype=np.float64)
        L.check(self.lib.hj_upload(self.h, current_stream(self.device), field, a.ctypes.data, 1))
        L.check(self.lib.hj_stream_sync(current_stream(self.device)))

    def download(self, field=L.FIELD_STATE):
        out = np.empty(self.shape, dtype=np.float64)
        L.check(self.lib.hj_download(self.h, current_stream(self.device), field, out.ctypes.data, 1))
        return out

    def step(self, t_end, factorCFL=0.8, comp=L.COMP_NONE, use_obstacle=False, maxStep=np.finfo(np.float64).max):
        """One CFL-limited TVD-RK3 step of every grid (each with its own dt).  Returns (t_new[nbatch], dt[nbatch])."""
        t_end = np.broadcast_to(np.asarray(t_end, dtype=np.float64), (self.nb,))
        params, dts, t_new = batch_step_plan(self.adapters, self.dxs, self.t, t_end, factorCFL, maxStep)
        if self._npar != params.shape[2]:
            L.check(self.lib.hj_set_system(self.h, L.SYS_FLOCK, None, int(params.shape[2])))
            self._npar = params.shape[2]
        params = np.ascontiguousarray(params)
        L.check(self.lib.hj_step_batch(self.h, current_stream(self.device), dts.ctypes.data, params.ctypes.data,
                                       int(comp), int(bool(use_obstacle))))
        self.t = t_new
        return t_new.copy(), dts

    def sync(self):
        L.check(self.lib.hj_stream_sync(current_stream(self.device)))

    def close(self):
        self._fin()
