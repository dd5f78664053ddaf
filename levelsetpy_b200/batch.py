"""Batched small grids: many independent 3-D grids, one Flock each, advanced together.

The reference handles a flock of flocks with a Python loop -- one ``odeCFL3(termLaxFriedrichs, ...)`` call per grid
(ode_cfl_3.py:11), each on an L2-sized 101^3 field that is launch/latency bound on a GPU.  ``BatchSolver`` keeps all
grids of identical shape in ONE resident field ``[nbatch, N0, N1, N2]`` and advances the whole batch with one fused
kernel per RK stage (C-ABI: hj_create_batch / hj_step_batch).  Each grid keeps its own Flock parameter block (re-derived
on each of the three RHS evaluations like flock.py:213) and its own CFL time step, exactly as the per-grid loop would.
A grid that has already reached its ``t_end`` is FINISHED: the per-grid loop would not call odeCFL3 for it any more
(ode_cfl_3.py:125), so its flock bookkeeping is not run, its dt is 0, its time does not move and the kernels skip
its CTAs (the field is left bit for bit as it is).
"""
import ctypes as C
import weakref

import numpy as np

from . import _lib as L
from .engine import current_stream, grid_signature, weno_mode_of
from .functors import resolve
from .integration import rk3_times

__all__ = ["BatchSolver", "batch_step_plan", "FlockBatchPlanner"]


def batch_step_plan(adapters, dxs, t, t_end, factorCFL, maxStep=np.finfo(np.float64).max):
    """Host side of one batched TVD-RK3 step: per grid, the three per-stage parameter blocks, the CFL time step
    (ode_cfl_3.py:142-143, with that grid's own dx) and the new time.  ``dxs``/``t``/``t_end`` are per-grid
    sequences.  Pure numpy (no device)."""
    nb = len(adapters)
    blocks = [[None] * nb for _ in range(3)]
    dts, t_new = np.empty(nb), np.empty(nb)
    for j, ad in enumerate(adapters):
        if not t_end[j] - t[j] > 0:                        # finished: no RHS evaluation, no bookkeeping, dt = 0
            dts[j], t_new[j] = 0.0, t[j]
            z = np.zeros(10)                               # header-only block; the kernels skip this element
            for k in range(3):
                blocks[k][j] = z
            continue
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


class FlockBatchPlanner:
    """The host side of a batched step for flocks that all have the same shape -- same number of birds, labels and
    neighbour radii, hence the same neighbour lists (the flock-of-flocks case of SURVEY.md 8d config 5) -- evaluated
    for the whole batch at once with numpy instead of one Python pass per flock.  Arithmetic and its ORDER are those
    of ``Flock._housekeeping`` (flock.py:147-188: the headings are updated agent by agent, each update seeing the
    already-updated headings of lower-indexed neighbours) and of the per-flock parameter block
    (functors._Flock.block), element for element, so blocks, dt and the birds' ``w_e`` are bit-identical to
    ``batch_step_plan``.  ``build`` returns None when the flocks are not uniform; the caller then takes the per-flock
    path."""

    MAX_NEIGH = 7       # np.sum adds fewer than 8 terms left to right; longer lists are summed pairwise

    def __init__(self, flocks, nbrs):
        self.flocks, self.nbrs = flocks, nbrs
        self.nb, self.N = len(flocks), flocks[0].N
        self._trig = {}

    @classmethod
    def build(cls, adapters):
        from .functors import _Flock
        if not adapters or any(not isinstance(ad, _Flock) or ad.mode != "flock" for ad in adapters):
            return None
        flocks = [ad.o for ad in adapters]
        N = flocks[0].N
        if N < 3 or 10 + 3 * (N - 1) > L.HJ_MAX_PARAMS:
            return None
        sig0 = None
        for f in flocks:
            if f.N != N or len(f.vehicles) != N:
                return None
            index = {id(b): i for i, b in enumerate(f.vehicles)}
            nbrs = []
            for i, b in enumerate(f.vehicles):
                try:
                    lst = [index[id(n)] for n in b.neighbors]
                except KeyError:
                    return None                       # a neighbour that is not a member of this flock
                if not 1 <= len(lst) <= cls.MAX_NEIGH:
                    return None
                # the neighbour lists must already be what _housekeeping would make them (it only ever adds)
                want = [j for j in list(range(i + 1, N)) + list(range(i - 1, -1, -1))
                        if np.abs(b.label - f.vehicles[j].label) < b.neigh_rad]
                if sorted(want) != sorted(lst):
                    return None
                nbrs.append(tuple(lst))
            sig = tuple(nbrs)
            if sig0 is None:
                sig0 = sig
            elif sig != sig0:
                return None
        return cls(flocks, sig0)

    def _cos_sin(self, th):
        out = np.empty(th.shape + (2,))
        cache = self._trig
        for idx, v in np.ndenumerate(th):
            v = float(v)
            cs = cache.get(v)
            if cs is None:
                cs = cache[v] = (float(np.cos(v)), float(np.sin(v)))      # the scalar calls the per-flock path makes
            out[idx] = cs
        return out[..., 0], out[..., 1]

    def plan(self, dxs, t, t_end, factorCFL, maxStep=np.finfo(np.float64).max):
        nb, N, nbrs = self.nb, self.N, self.nbrs
        fl = self.flocks
        t = np.asarray(t, dtype=np.float64)
        active = (np.asarray(t_end, dtype=np.float64) - t) > 0           # finished grids take no RHS evaluation
        W = np.array([[b.w_e for b in f.vehicles] for f in fl], dtype=np.float64).reshape(nb, N)
        cs = np.array([[np.asarray(b.cur_state, dtype=np.float64)[:3, 0] for b in f.vehicles] for f in fl]).reshape(nb, N, 3)
        ve = np.array([[b.v_e for b in f.vehicles] for f in fl], dtype=np.float64).reshape(nb, N)
        vp = np.array([[b.v_p for b in f.vehicles] for f in fl], dtype=np.float64).reshape(nb, N)
        wp0 = np.array([f.vehicles[0].w_p for f in fl], dtype=np.float64)
        X, Y, TH = cs[..., 0], cs[..., 1], cs[..., 2]
        C_, S_ = self._cos_sin(TH)
        K = N - 1
        npar = 10 + 3 * K
        params = np.zeros((3, nb, npar))
        # state-only parts of the block (the birds do not move during the solve)
        th_low = np.stack([np.min(TH[:, list(nbrs[i])], axis=1) for i in range(N)], axis=1)    # min neighbour heading
        th_up0 = np.max(TH[:, list(nbrs[0])], axis=1)
        a_abs0 = np.abs(vp * C_)                     # functors._Flock._abs_alpha
        a_abs1 = np.abs(ve * S_)
        a1 = ve[:, 0] - vp[:, 0] * C_[:, 0]          # functors._Flock._att
        a2 = vp[:, 0] * S_[:, 0]
        al_att = [np.abs(ve[:, 0] - vp[:, 0] * C_[:, 0]) + np.abs(th_up0 * Y[:, 0]),
                  np.abs(vp[:, 0] * S_[:, 0]) + np.abs(th_up0 * X[:, 0]),
                  wp0 + th_up0]
        amax = []
        for d, per_other in enumerate((a_abs0, a_abs1, th_low)):
            m = per_other[:, 1]
            for i in range(2, N):
                m = np.maximum(m, per_other[:, i])
            amax.append(np.maximum(m, al_att[d]))
        for k in range(3):
            for i in range(N):                       # Flock._update_headings, agent by agent (flock.py:170-188)
                lst = nbrs[i]
                ssum = W[:, lst[0]].copy()
                for j in lst[1:]:
                    ssum += W[:, j]
                W[:, i] = (1 / (1 + len(lst))) * (W[:, i] + ssum)
            P = params[k]
            P[:, 0] = float(K)
            P[:, 1] = 1.0
            wmax = W[:, nbrs[0][0]]
            for j in nbrs[0][1:]:
                wmax = np.maximum(wmax, W[:, j])
            P[:, 2] = wmax
            P[:, 3], P[:, 4], P[:, 5], P[:, 6] = a1, a2, X[:, 0], Y[:, 0]
            P[:, 7], P[:, 8], P[:, 9] = amax
            for i in range(1, N):
                c0 = 10 + 3 * (i - 1)
                P[:, c0], P[:, c0 + 1], P[:, c0 + 2] = -C_[:, i], -S_[:, i], -W[:, i]
            if k == 0:
                dx = np.asarray(dxs, dtype=np.float64).reshape(nb, 3)
                inv = 0
                for d in range(3):
                    inv = inv + (amax[d] / dx[:, d])                     # artificial_diss_glf.py:107, dims in order
                step_bound = 1 / inv
        for f, row, act in zip(fl, W, active):   # the flocks keep their mutated headings, like the reference's do
            if not act:
                continue                             # ... and a finished flock is not touched at all
            f.attacked_idx = 0
            for b, w in zip(f.vehicles, row):
                b.w_e = np.float64(w)
        dts = np.minimum(np.minimum(factorCFL * step_bound, np.asarray(t_end, dtype=np.float64) - t), maxStep)
        dts = np.where(active, dts, 0.0)
        return params, dts, np.where(active, rk3_times(t, dts)[2], t)


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
        self.planner = FlockBatchPlanner.build(self.adapters)     # None: flocks of different shapes -> per-flock host pass

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
        if self.planner is not None:
            params, dts, t_new = self.planner.plan(self.dxs, self.t, t_end, factorCFL, maxStep)
        else:
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
