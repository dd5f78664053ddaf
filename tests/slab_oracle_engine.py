"""TEST INFRASTRUCTURE: a CPU stand-in for the per-slab CUDA context (levelsetpy_b200.engine.Engine with
``slab=(lo, hi)``) whose stage "kernel" is the numpy oracle evaluated on the haloed slab.  It lets the world-size-2
gloo tests run the REAL SlabSolver (partitioning, halo send/recv ordering, edge-halo fill, alpha max-allreduce, dt)
on CPU and compare against the single-domain oracle.  Same layout as the device context: three RK buffers of
(3 + n0 + 3) dim-0 planes, interior at plane offset 3."""
import numpy as np
import torch

from oracle import hj_oracle as orc
from oracle import systems as osys

G = 3


class _SubGrid:
    pass


class OracleSlabEngine:
    def __init__(self, grid, weno, device, slab, backend=None):
        lo, hi = slab
        self.grid, self.weno, self.lo, self.hi = grid, weno, lo, hi
        self.D = int(grid.dim)
        Ng = [int(x) for x in np.asarray(grid.N).reshape(-1)]
        self.n0 = hi - lo
        self.N = [self.n0] + Ng[1:]
        self.dx = [float(x) for x in np.asarray(grid.dx).reshape(-1)]
        self.plane_elems = int(np.prod(Ng[1:]))
        self.field_elems = self.plane_elems * (self.n0 + 2 * G)
        self.hshape = (self.n0 + 2 * G,) + tuple(Ng[1:])
        self._buf = [np.zeros(self.field_elems) for _ in range(3)]
        self._aux = self._obs = None
        self.nparams = 0
        # haloed sub-grid: dim 0 carries n0+6 nodes whose interior coordinates are the global ones
        sg = _SubGrid()
        v0 = np.asarray(grid.vs[0], dtype=np.float64).reshape(-1)
        ext = v0[lo] + self.dx[0] * (np.arange(self.n0 + 2 * G) - G)
        ext[G:G + self.n0] = v0[lo:hi]
        sg.dim, sg.dx, sg.shape = self.D, grid.dx, self.hshape
        sg.N = np.array(self.hshape).reshape(-1, 1)
        sg.vs = [ext] + [np.asarray(grid.vs[d]).reshape(-1) for d in range(1, self.D)]
        sg.bdry = [_extrap] + list(grid.bdry[1:])
        sg.bdryData = [None] + list(grid.bdryData[1:]) if getattr(grid, "bdryData", None) is not None else None
        self.sub = sg
        gd = grid.bdryData[0] if getattr(grid, "bdryData", None) is not None else None
        self.tz0 = bool(getattr(gd, "towardZero", False)) if gd is not None else False

    # --- the interface SlabSolver uses
    def buffer_tensor(self, b):
        return torch.from_numpy(self._buf[b])

    def _interior(self, a):
        return a.reshape(self.hshape)[G:G + self.n0]

    def upload(self, a, field=0):
        a = np.asarray(a, dtype=np.float64).reshape(self.N)
        if field == 0:
            self._interior(self._buf[0])[...] = a
        elif field == 1:
            self._aux = a.copy()
        else:
            self._obs = a.copy()

    def download(self, like=None, field=0, shape=None):
        out = self._interior(self._buf[0]).copy()
        return out.reshape(shape) if shape is not None else out

    def set_system(self, system_id, params, tables=()):
        p = np.asarray(params, dtype=np.float64)
        self.nparams = p.size
        if system_id == 1:
            self.sys = osys.DubinsVehicleRel(self.sub, p[0], p[2])
        elif system_id == 2:
            self.sys = osys.DoubleIntegrator(self.sub, p[0])
        else:
            raise NotImplementedError(system_id)

    def alpha_max(self, t=0.0):
        sd = orc.OracleSchemeData(grid=self.sub, hamFunc=self.sys.hamiltonian, partialFunc=self.sys.dissipation)
        a = []
        for d in range(self.D):
            al = self.sys.dissipation(t, None, None, None, sd, d)
            if isinstance(al, np.ndarray):
                al = np.max(np.broadcast_to(al, self.hshape)[G:G + self.n0])
            a.append(float(al))
        return np.array(a), None

    def stage_io(self, stage):
        return (0, 0, 1, 2)[stage], (0, 1, 2, 0)[stage]

    def fill_edge_halo(self, b, side):
        h = self._buf[b].reshape(self.hshape)
        gh = orc.add_ghost_extrapolate(h[G:G + self.n0], 0, G, self.tz0)
        if side == 0:
            h[:G] = gh[:G]
        else:
            h[G + self.n0:] = gh[G + self.n0:]

    def stage(self, stage, t, dt, params=None, comp=0, use_obstacle=False, want_reduce=False):
        i, o = self.stage_io(stage)
        sd = orc.OracleSchemeData(grid=self.sub, hamFunc=self.sys.hamiltonian, partialFunc=self.sys.dissipation)
        h = self._buf[i].reshape(self.hshape)
        ydot, _ = orc.term_lax_friedrichs(t, h.reshape(-1, 1), sd, self.weno)
        ydot = ydot.reshape(self.hshape)[G:G + self.n0]
        yin = h[G:G + self.n0]
        y0 = self._interior(self._buf[0]).copy()
        if stage == 1:
            out = yin + dt * ydot
        elif stage == 2:
            out = 0.25 * (3 * y0 + (yin + dt * ydot))
        else:
            out = (1 / 3) * (y0 + 2 * (yin + dt * ydot))
            if comp == 1:
                out = np.minimum(out, y0)
            elif comp == 2:
                out = np.maximum(out, y0)
            elif comp == 3:
                out = np.minimum(out, self._aux)
            elif comp == 4:
                out = np.maximum(out, self._aux)
            if use_obstacle:
                out = np.maximum(out, -self._obs)
        self._interior(self._buf[o])[...] = out


class TwoPassOracleSlabEngine(OracleSlabEngine):
    """Stands in for a product-system context on the dimension-split path: a stage is two calls.  Pass 1 must not
    depend on the dim-0 halo planes (it poisons them to prove it), pass 2 evaluates the stage and therefore needs
    the halos the exchange posted BEFORE pass 1 has delivered by then."""

    def is_split(self):
        return True

    def stage(self, stage, t, dt, params=None, comp=0, use_obstacle=False, want_reduce=False, which_pass=0):
        if which_pass == 1:
            self.log.append(("pass1", stage))
            return
        if which_pass == 2:
            assert self.log and self.log[-1] == ("pass1", stage), "pass 2 without its pass 1"
        self.log.append(("pass2" if which_pass else "both", stage))
        OracleSlabEngine.stage(self, stage, t, dt, params, comp, use_obstacle, want_reduce)

    log = None

    def __init__(self, *a, **k):
        OracleSlabEngine.__init__(self, *a, **k)
        self.log = []


class RangedOracleSlabEngine(OracleSlabEngine):
    """Stands in for a whole-system plane-ring context: ``stage_range`` advances only dim-0 planes [z0, z1) of the slab.
    The stage is evaluated on a haloed sub-block holding exactly the planes that range may read (z0-3 .. z1+2), so an
    interior range evaluated before the halos arrive provably does not depend on them (they are NaN-poisoned by the
    test), and stage 3's in-place write of buffer 0 stays pointwise."""

    def supports_range(self):
        return True

    def stage_range(self, stage, z0, z1, t, dt, params=None, comp=0, use_obstacle=False, want_reduce=0):
        self.log.append((stage, z0, z1))
        i, o = self.stage_io(stage)
        src = self._buf[i].reshape(self.hshape)
        lo, hi = z0, z1 + 2 * G                     # haloed planes z0-3 .. z1+2 in buffer coordinates
        sub = _SubGrid()
        sub.__dict__.update(self.sub.__dict__)
        sub.shape = (hi - lo,) + self.hshape[1:]
        sub.N = np.array(sub.shape).reshape(-1, 1)
        sub.vs = [np.asarray(self.sub.vs[0])[lo:hi]] + list(self.sub.vs[1:])
        if hasattr(self.sys, "grid"):
            pass
        sysd = type(self.sys)(sub, *self._sys_args)
        sd = orc.OracleSchemeData(grid=sub, hamFunc=sysd.hamiltonian, partialFunc=sysd.dissipation)
        block = np.ascontiguousarray(src[lo:hi])
        ydot, _ = orc.term_lax_friedrichs(t, block.reshape(-1, 1), sd, self.weno)
        ydot = ydot.reshape(sub.shape)[G:G + (z1 - z0)]
        yin = block[G:G + (z1 - z0)]
        y0 = self._interior(self._buf[0])[z0:z1].copy()
        if stage == 1:
            out = yin + dt * ydot
        elif stage == 2:
            out = 0.25 * (3 * y0 + (yin + dt * ydot))
        else:
            out = (1 / 3) * (y0 + 2 * (yin + dt * ydot))
            if comp == 1:
                out = np.minimum(out, y0)
            elif comp == 2:
                out = np.maximum(out, y0)
        self._interior(self._buf[o])[z0:z1] = out

    def set_system(self, system_id, params, tables=()):
        OracleSlabEngine.set_system(self, system_id, params, tables)
        p = np.asarray(params, dtype=np.float64)
        self._sys_args = (p[0], p[2]) if system_id == 1 else (p[0],)

    log = None

    def __init__(self, *a, **k):
        OracleSlabEngine.__init__(self, *a, **k)
        self.log = []


def _extrap(*a, **k):
    raise RuntimeError("token only")


_extrap.__name__ = "addGhostExtrapolate"


class PeerOracleSlabEngine(TwoPassOracleSlabEngine):
    """The peer-memory halo interface of the CUDA context (hj_halo_export / attach / push / wait / set_fused / signal,
    hj_split_cols / hj_stage_pass_cols) on CPU buffers, for the protocols of LocalWorld: a descriptor is a key into a
    process-wide registry, a push copies the edge planes (optionally only some columns of the flattened trailing dims)
    into the neighbour's halo planes and bumps its arrival counter, a wait ASSERTS that the pushes it depends on have
    been made (one process, one "stream": a wait that would have to block is a protocol bug), and pass 2 with fused
    sides writes its edge planes into those neighbours itself."""

    registry = {}

    def __init__(self, *a, **k):
        TwoPassOracleSlabEngine.__init__(self, *a, **k)
        self.flags = np.zeros((2, 3), dtype=np.int64)       # [side of MY halo: 0 lower, 1 upper][buffer]
        self.pushed = np.zeros((2, 3), dtype=np.int64)      # [towards lower, upper neighbour][buffer]
        self.waited = np.zeros(3, dtype=np.int64)
        self.peers = [None, None]
        self.fused_sides = 0
        self.calls = []

    def halo_export(self):
        key = len(PeerOracleSlabEngine.registry)
        PeerOracleSlabEngine.registry[key] = self
        return key.to_bytes(8, "little") + bytes(504)

    def halo_attach(self, lower, upper):
        pick = lambda d: None if d is None else PeerOracleSlabEngine.registry[int.from_bytes(d[:8], "little")]
        self.peers = [pick(lower), pick(upper)]

    def halo_detach(self):
        self.peers = [None, None]

    def _planes(self, side):
        """(my source planes, the neighbour's destination planes) towards the neighbour on ``side`` of buffer views."""
        n0 = self.n0
        return (slice(G, 2 * G), slice(self.peers[0].n0 + G, self.peers[0].n0 + 2 * G)) if side == 0 else \
               (slice(n0, n0 + G), slice(0, G))

    def _copy(self, b, side, cols):
        src_sl, dst_sl = self._planes(side)
        peer = self.peers[side]
        src = self._buf[b].reshape(self.n0 + 2 * G, -1)[src_sl]
        dst = peer._buf[b].reshape(peer.n0 + 2 * G, -1)[dst_sl]
        if cols is None:
            dst[...] = src
        else:
            a, e, row_len = cols
            assert src.shape[1] % row_len == 0
            src.reshape(G, -1, row_len)[:, :, a:e].shape  # noqa: B018 (shape check)
            dst.reshape(G, -1, row_len)[:, :, a:e] = src.reshape(G, -1, row_len)[:, :, a:e]

    def _bump(self, b, side):
        self.pushed[side][b] += 1
        self.peers[side].flags[1 - side][b] = self.pushed[side][b]       # my upper neighbour's LOWER halo, and vice versa

    def halo_push(self, b, cols=None, sides=3):
        self.calls.append(("push", b, cols, sides))
        for side in (0, 1):
            if self.peers[side] is not None and sides & (1 << side):
                self._copy(b, side, cols)
                self._bump(b, side)

    def halo_signal(self, b, sides=3):
        self.calls.append(("signal", b, sides))
        for side in (0, 1):
            if self.peers[side] is not None and sides & (1 << side):
                self._bump(b, side)

    def halo_wait(self, b, npush=1):
        self.calls.append(("wait", b, npush))
        self.waited[b] += npush
        for side in (0, 1):
            if self.peers[side] is not None:
                assert self.flags[side][b] >= self.waited[b], \
                    "wait on buffer %d (side %d) before the push it depends on: %d < %d" % (
                        b, side, self.flags[side][b], self.waited[b])

    def halo_set_fused(self, sides=3):
        self.fused_sides = int(sides)

    def split_cols(self):
        return int(np.prod(self.N[1:])), 2

    def stage(self, stage, t, dt, params=None, comp=0, use_obstacle=False, want_reduce=False, which_pass=0):
        TwoPassOracleSlabEngine.stage(self, stage, t, dt, params, comp, use_obstacle, want_reduce, which_pass)
        if which_pass == 2 and self.fused_sides:           # the kernel's own stores into the neighbours' halo planes
            o = self.stage_io(stage)[1]
            for side in (0, 1):
                if self.peers[side] is not None and self.fused_sides & (1 << side):
                    self._copy(o, side, None)

    def stage_cols(self, stage, a, e, t, dt, params=None, comp=0, use_obstacle=False, want_reduce=0):
        """Pass 2 on columns [a, e) of the flattened trailing dims: the whole stage is evaluated from the buffer as it
        is -- the halo columns outside [a, e) may be stale -- and only the columns [a, e) of the result are kept (a
        star stencil reads the dim-0 halos at its own column only)."""
        self.log.append(("cols", stage, a, e))
        i, o = self.stage_io(stage)
        keep = self._buf[o].copy()
        OracleSlabEngine.stage(self, stage, t, dt, params, comp, use_obstacle, want_reduce)
        new = self._buf[o].reshape(self.n0 + 2 * G, -1)
        old = keep.reshape(self.n0 + 2 * G, -1)
        old[G:G + self.n0, a:e] = new[G:G + self.n0, a:e]
        self._buf[o][...] = keep
