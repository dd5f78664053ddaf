"""Multi-GPU: slab decomposition of the grid along dim 0 (the slowest-varying dim, so a slab is contiguous), one
process per GPU.

Per RK stage every rank
  1. refreshes the 3 stored halo planes of the buffer the stage reads: interior faces get the neighbour's edge
     planes (send/recv, NCCL over NVLink under ``torch.distributed``; wrap-around pair when dim 0 is periodic),
     global non-periodic faces are filled locally with the addGhostExtrapolate ghosts of the rank's own edge planes
     (add_ghost_extrapolate.py:88-110);
  2. ('intended' WENO only) max-allreduces the per-dim max(D1^2) that defines the WENO epsilon
     (upwind_first_weno5a.py:154-156);
  3. launches the fused stage kernel on its slab.
max_x alpha_d -- hence dt -- is state-only for every registered system: it is max-allreduced once and dt is then
computed identically on every rank with the reference's own host arithmetic (artificial_diss_glf.py:104-109,
ode_cfl_3.py:142-143).  No other data-path collective exists.

Transport of the halo planes: on CUDA contexts the rank PUSHES its edge planes into the neighbours' stored halos
through peer memory (C-ABI hj_halo_*: CUDA IPC handles exchanged once, then one copy-engine transfer per neighbour and
stage, signalled by a counter -- no SM copy kernel contends with the stage kernel that runs under the transfer);
``transport="p2p"`` keeps the ``torch.distributed`` send/recv path (NCCL on GPUs, gloo in the CPU tests).  The host-side
group is only used to move the descriptors and for the one-off max-allreduce.  ``LocalWorld`` runs several slabs inside
one process (the single-GPU test of the halo path) over the same two transports.
"""
import numpy as np

from . import _lib as L
from .engine import Engine, weno_mode_of
from .functors import resolve
from .integration import rk3_times
from .utilities import isfield, warn

__all__ = ["partition", "SlabSolver", "DistComm", "LocalWorld"]

GHOST = L.HJ_GHOST


def partition(n0, world):
    """Contiguous [lo, hi) plane ranges of dim 0, sizes differing by at most one (larger slabs first)."""
    base, rem = divmod(int(n0), int(world))
    if base < GHOST:
        raise ValueError("slab decomposition needs >= %d planes of dim 0 per rank (N[0]=%d over %d ranks)" % (GHOST, n0, world))
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


class DistComm:
    """torch.distributed transport (backend nccl on GPUs, gloo on CPU)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self._side = None

    def post(self, sends, recvs):
        """sends/recvs: lists of (tensor, peer, tag), posted as one batch.  Returns a handle for ``wait``: work
        launched on the current stream after ``post`` runs under the transfer until ``wait`` orders the stream behind
        it.  CUDA tensors: the batch is issued from a side stream (ordered after everything already on the current
        stream), so that whichever stream the backend runs its copy kernels on, the caller's next kernel is not
        queued behind them."""
        dist = self.dist
        ops = [dist.P2POp(dist.irecv, t, p, self.group, tag) for t, p, tag in recvs]
        ops += [dist.P2POp(dist.isend, t, p, self.group, tag) for t, p, tag in sends]
        if not ops:
            return None
        if ops[0].tensor.is_cuda:
            import torch
            if self._side is None:
                self._side = torch.cuda.Stream(device=ops[0].tensor.device)
            cur = torch.cuda.current_stream(ops[0].tensor.device)
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                for req in dist.batch_isend_irecv(ops):
                    req.wait()                      # orders the side stream (not the host) behind the transfer
            return ("stream", cur)
        return ("reqs", dist.batch_isend_irecv(ops))

    def wait(self, handle):
        if handle is None:
            return
        kind, h = handle
        if kind == "stream":
            h.wait_stream(self._side)
        else:
            for req in h:
                req.wait()

    def exchange(self, sends, recvs):
        """Posted as one batch; returns after local completion is ordered on the current stream (NCCL) /
        finished (gloo)."""
        self.wait(self.post(sends, recvs))

    def allreduce_max(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return t

    def barrier(self):
        self.dist.barrier(group=self.group)

    def allgather_bytes(self, b):
        """Every rank's ``bytes`` object, in rank order."""
        out = [None] * self.world
        self.dist.all_gather_object(out, b, group=self.group)
        return out


class _LocalComm:
    """One member of a LocalWorld: exchanges are deferred to the world, which copies between its slabs."""

    def __init__(self, world, rank):
        self.world_obj, self.rank, self.world = world, rank, world.world


class SlabSolver:
    """One rank's slab of a TVD-RK3 / WENO5 / GLF solve.

    schemeData: the reference's bundle (grid = the GLOBAL grid).  ``comm``: DistComm() by default.
    ``engine_factory(grid, weno, device, slab, backend)`` builds the per-slab context (default: the CUDA Engine)."""

    def __init__(self, schemeData, device=0, backend=None, comm=None, engine_factory=None, transport="auto", overlap=True,
                 pieces="auto", fused="auto"):
        """``transport``: "peer" (copy-engine pushes into peer memory, hj_halo_*), "p2p" (send/recv of the group) or
        "auto" (peer wherever the per-slab engine offers it).  ``overlap=False``: exchange first, then compute (the
        protocol of the gather / intended paths; attribution runs).  ``pieces``: product systems over the peer
        transport advance pass 2 in that many column pieces and push each finished piece of the edge planes while the
        next is computed ("auto": 4 when the halo is a large fraction of the slab, else 1 = whole-plane pushes).
        ``fused`` ("auto" = on for product systems over the peer transport): pass 2 itself stores its edge planes into
        the neighbours' halo planes (peer stores over NVLink inside the stage kernel, hj_halo_set_fused) -- compute and
        halo transfer are ONE kernel, no copy follows it; ``pieces`` is then irrelevant.  ``fused="hybrid"``: only the
        planes for the upper neighbour go that way, those for the lower one are pushed by the copy engines behind the
        kernel (under the next stage's pass 1), so that the two transports share the link.  "auto" picks hybrid when the
        halo is more than half of the slab (41 planes over 8 ranks: 49.2 ms per step against 53.3 ms with both
        sides stored from the kernel and 57.7 ms with copy-engine pieces), both sides otherwise."""
        sd = schemeData
        for f in ("grid", "hamFunc", "partialFunc"):
            assert isfield(sd, f), "%s not in bundle thisschemeData" % f
        self.comm = comm if comm is not None else DistComm()
        self.rank, self.world = self.comm.rank, self.comm.world
        g = sd.grid
        self.grid = g
        self.weno = weno_mode_of(sd)
        self.adapter = resolve(sd.hamFunc, sd.partialFunc, g)
        N0 = int(np.asarray(g.N).reshape(-1)[0])
        self.lo, self.hi = partition(N0, self.world)[self.rank]
        self.n0 = self.hi - self.lo
        factory = engine_factory if engine_factory is not None else Engine
        self.eng = factory(g, self.weno, device, (self.lo, self.hi), backend)
        self.periodic0 = getattr(g.bdry[0], "__name__", "") == "addGhostPeriodic"
        self.lo_peer = self.rank - 1 if self.rank > 0 else (self.world - 1 if self.periodic0 else None)
        self.hi_peer = self.rank + 1 if self.rank < self.world - 1 else (0 if self.periodic0 else None)
        self.plane = self.eng.plane_elems
        self.bufs = [self.eng.buffer_tensor(b) for b in range(3)]
        self._alpha = None
        self._overlap = None
        self._ranged = None
        self._tables = list(enumerate(self.adapter.tables(g)))
        self.shape = (self.n0,) + tuple(int(x) for x in np.asarray(g.N).reshape(-1)[1:])
        self._no_overlap = not overlap
        self.mode = "full"        # attribution runs (bench.py): "compute" = no exchange, "comm" = no stage kernels
        self._pieces_arg = pieces
        self._fused_arg = fused
        self._fused = None
        self._cols = None
        self._primed = False      # pieces protocol: the halos of the state the next step reads are on their way
        if transport not in ("auto", "peer", "p2p"):
            raise ValueError("transport must be 'auto', 'peer' or 'p2p', got %r" % (transport,))
        can_peer = hasattr(self.eng, "halo_export")
        if transport == "peer" and not can_peer:
            raise NotImplementedError("the per-slab engine has no peer-memory halo transport")
        self.peer = can_peer and transport != "p2p"
        if self.peer and hasattr(self.comm, "allgather_bytes"):
            self.attach_peers(self.comm.allgather_bytes(self.eng.halo_export()))

    def attach_peers(self, descs):
        """``descs``: every rank's ``Engine.halo_export()`` bytes in rank order (moved by the host-side group)."""
        pick = lambda r: descs[r] if (r is not None and r != self.rank) else None
        self.eng.halo_attach(pick(self.lo_peer), pick(self.hi_peer))

    def close(self):
        """Collective: unmap the neighbours' buffers on every rank BEFORE any rank frees its own (memory exported through
        CUDA IPC must outlive its importers), then destroy the context."""
        if self.peer and hasattr(self.eng, "halo_detach"):
            self.eng.halo_detach()
            if hasattr(self.comm, "barrier"):
                self.comm.barrier()
        if hasattr(self.eng, "close"):
            self.eng.close()

    # ------------------------------------------------------------------ fields
    def upload(self, slab, field=L.FIELD_STATE):
        """``slab``: this rank's planes [lo, hi) of the field, shape (hi-lo, N1, ...), numpy or torch CUDA."""
        self.eng.upload(slab, field)
        if field == L.FIELD_STATE:
            self.state_changed()

    def state_changed(self):
        """Collective.  The resident state was replaced (upload, or written through ``eng.buffer_tensor``): halo pieces
        pushed ahead for the old state are consumed and dropped."""
        if self._primed and self.mode != "compute":
            self.eng.halo_wait(0, 1 if self.fused() else len(self.pieces() or [None]))
        self._primed = False

    def set_mode(self, mode):
        """Collective: "full", "compute" (no exchange) or "comm" (no stage kernels) -- attribution runs."""
        self.state_changed()
        self.mode = mode

    def download(self, out=None):
        if out is not None:
            o = np.asarray(out)
            if o.dtype == np.float64 and o.flags.c_contiguous and hasattr(self.eng, "lib"):
                self.eng.download(shape=self.shape, out=o)      # straight into the caller's (pinned) buffer
            else:
                np.copyto(o.reshape(self.shape), self.eng.download(shape=self.shape))
            return out
        return self.eng.download(shape=self.shape)

    # ------------------------------------------------------------------ halos
    def _faces(self, b):
        """(send_up, send_down, recv_lo, recv_hi) views of buffer b: my top / bottom 3 interior planes and my
        lower / upper halo planes."""
        p, n0, t = self.plane, self.n0, self.bufs[b]
        return (t[n0 * p:(n0 + GHOST) * p], t[GHOST * p:2 * GHOST * p], t[0:GHOST * p],
                t[(n0 + GHOST) * p:(n0 + 2 * GHOST) * p])

    def halo_ops(self, b):
        """Point-to-point operations that refresh buffer b's halos: tag 0 travels up (towards higher ranks), tag 1
        down.  Receives are listed in the order the peer posts the matching sends (up first, then down), so the
        two messages of a 2-rank periodic ring pair up correctly on NCCL, which matches by order, not tag."""
        up, down, rlo, rhi = self._faces(b)
        sends, recvs = [], []
        if self.hi_peer is not None and self.hi_peer != self.rank:
            sends.append((up, self.hi_peer, 0))
        if self.lo_peer is not None and self.lo_peer != self.rank:
            sends.append((down, self.lo_peer, 1))
            recvs.append((rlo, self.lo_peer, 0))
        if self.hi_peer is not None and self.hi_peer != self.rank:
            recvs.append((rhi, self.hi_peer, 1))
        return sends, recvs

    def finish_halos(self, b):
        """The halo faces no neighbour fills: periodic self-wrap (world == 1) and extrapolated global edges."""
        up, down, rlo, rhi = self._faces(b)
        if self.lo_peer == self.rank:
            rlo.copy_(up)
            rhi.copy_(down)
        if self.lo_peer is None:
            self.eng.fill_edge_halo(b, 0)
        if self.hi_peer is None:
            self.eng.fill_edge_halo(b, 1)

    def post(self, b):
        """Start refreshing the neighbours' (peer transport) / my (p2p) halos of buffer b; work queued on the current
        stream afterwards runs under the transfer until ``wait``."""
        if self.mode == "compute":
            return ("none", b)
        if self.peer:
            self.eng.halo_push(b)
            return ("peer", b)
        return self.comm.post(*self.halo_ops(b))

    def wait(self, handle):
        if handle is not None and handle[0] == "none":
            return
        if handle is not None and handle[0] == "peer":
            self.eng.halo_wait(handle[1])
        else:
            self.comm.wait(handle)

    def exchange(self, b):
        if isinstance(self.comm, _LocalComm):
            raise RuntimeError("slabs of a LocalWorld are stepped through LocalWorld.step")
        if self.peer:
            self.wait(self.post(b))
        else:
            self.comm.exchange(*self.halo_ops(b))
        self.finish_halos(b)

    # ------------------------------------------------------------------ dt
    def local_alpha(self):
        """max over THIS slab of the state-only alpha_d (device reduction, cached by the context)."""
        self.eng.set_system(self.adapter.system_id, self.adapter.block(), self._tables)
        return np.asarray(self.eng.alpha_max()[0], dtype=np.float64)

    def step_bound(self, block):
        """stepBound = 1 / sum_d max_x alpha_d / dx_d with the maxima taken over the WHOLE grid
        (artificial_diss_glf.py:104-109, dims summed in order)."""
        al = self.adapter.alphas(block)
        if al is None:
            if self._alpha is None:
                import torch
                local, _ = self.eng.alpha_max()
                t = torch.tensor(np.asarray(local, dtype=np.float64), device=self.bufs[0].device)  # system set by begin_step
                self._alpha = [float(x) for x in self.comm.allreduce_max(t).cpu().numpy()]
            al = self._alpha
        inv = 0
        for d in range(self.eng.D):
            inv = inv + (al[d] / self.eng.dx[d])
        return 1 / inv

    # ------------------------------------------------------------------ stepping
    def begin_step(self, t, t_end, factorCFL, maxStep=np.finfo(np.float64).max):
        """Host part of one odeCFL3 step (ode_cfl_3.py:129-143): parameter blocks for the three RHS evaluations
        and the CFL-limited dt -- identical on every rank."""
        ad = self.adapter
        blocks = [ad.block()]
        self.eng.set_system(ad.system_id, blocks[0], self._tables)
        sb = self.step_bound(blocks[0])
        dt = float(np.min(np.hstack((factorCFL * sb, t_end - t, maxStep))))
        if ad.time_varying:
            safety = min(1.0, 1.2 * factorCFL)
            for name in ("Second", "Third"):
                b = ad.block()
                blocks.append(b)
                sbk = self.step_bound(b)
                if dt > safety * sbk:
                    warn("%s substep violated CFL effective number %s" % (name, dt / sbk))
        else:
            blocks = [None, None, None]
        self._step = (t, dt, blocks)
        return dt

    def run_stage(self, stage, comp=L.COMP_NONE, use_obstacle=False, which_pass=0):
        """Stage kernel(s) on this slab; the halos of the buffer it reads must already be current (pass 1 of a
        product system's stage reads none)."""
        t, dt, blocks = self._step
        if self.mode == "comm":
            return
        if self.weno == "intended" and which_pass != 2:
            eps = self.eng.eps_prepass(self.eng.stage_io(stage)[0])
            if self.world > 1:
                self.comm.allreduce_max(eps)
        kw = {"which_pass": which_pass} if which_pass else {}
        self.eng.stage(stage, t, dt, blocks[stage - 1], comp, use_obstacle, **kw)

    def overlapped(self):
        """Product systems on the dimension-split path under as_shipped WENO: the halo exchange of a stage runs
        under its first kernel.  ('intended' needs the halos for the WENO eps pre-pass, so it exchanges first.)"""
        return self.two_pass() and (self.peer or hasattr(self.comm, "post"))

    def ranged(self):
        """Whole 3-D systems on the plane-ring backend under as_shipped WENO: the planes whose dim-0 stencil stays
        inside the slab are advanced under the halo exchange, the two 3-plane edge ranges after it (hj_stage_range)."""
        if self._ranged is None:
            r = bool(self.weno != "intended" and not self._no_overlap and self.eng.D == 3 and self.n0 > 2 * GHOST
                     and not self.two_pass() and getattr(self.eng, "supports_range", lambda: False)())
            if not self._ready():
                return r
            self._ranged = r
        return self._ranged

    def fused(self):
        """Product system over the peer transport: pass 2 pushes its own edge planes (one kernel computes and moves)."""
        if self._fused is None:
            f = bool(self._fused_arg and self.peer and self.two_pass() and hasattr(self.eng, "halo_set_fused")
                     and self.world > 1)
            if not self._ready():
                return f                      # asked before the first step: the context cannot tell yet, do not cache
            self._fused = f
        return self._fused

    def hybrid(self):
        """Fused pushes towards the upper neighbour only, copy engines towards the lower one."""
        if self._fused_arg == "hybrid":
            return True
        return self._fused_arg == "auto" and 2 * GHOST * 2 > self.n0        # both halos together exceed half the slab

    def _ready(self):
        """The per-slab context knows its system and holds a state (true from the first begin_step on)."""
        return getattr(self, "_step", None) is not None

    def _step_fused(self, comp, use_obstacle):
        """Every stage: pass 1 | wait for the neighbours' planes of the buffer this stage reads (stored into my halos by
        THEIR pass 2 of the previous stage) | pass 2, whose stores of my edge planes go to the neighbours' halos of the
        buffer it writes as well | signal.  No copy is ever queued in the steady state."""
        t, dt, blocks = self._step
        talk, work = self.mode != "compute", self.mode != "comm"
        sides = (2 if self.hybrid() else 3) if self.mode == "full" else 0
        self.eng.halo_set_fused(sides)
        if not self._primed:
            if talk:
                self.eng.halo_push(0)                     # the state as uploaded: one copy-engine push of whole planes
            self._primed = True
        for stage in (1, 2, 3):
            b_in, b_out = self.eng.stage_io(stage)
            self.run_stage(stage, comp, use_obstacle, which_pass=1)
            self.finish_halos(b_in)
            if talk:
                self.eng.halo_wait(b_in, 1)
            if work:
                self.run_stage(stage, comp, use_obstacle, which_pass=2)
            if talk:
                if sides:
                    self.eng.halo_signal(b_out, sides)
                if sides != 3:                            # the other side(s): copy engines, behind the kernel (hybrid;
                    self.eng.halo_push(b_out, sides=3 & ~sides)   # attribution mode "comm": everything)

    def pieces(self):
        """Column pieces [(begin, end), ...] of pass 2 and the axis length, or None for whole-plane pushes."""
        if self._cols is None:
            k = self._pieces_arg
            if not (self.peer and self.two_pass() and hasattr(self.eng, "split_cols")) or self.fused():
                k = 1
            elif k == "auto":
                k = 4 if 2 * GHOST * 4 > self.n0 else 1           # the halo is more than a quarter of the slab
            if int(k) <= 1:
                self._cols = (0, None)
            else:
                V, q = self.eng.split_cols()
                per = -(-V // (int(k) * q)) * q
                self._cols = (V, [(a, min(V, a + per)) for a in range(0, V, per)])
        return self._cols[1]

    def _step_pieces(self, comp, use_obstacle):
        """Product system, peer transport: every stage = pass 1, then pass 2 piece by piece; a piece waits for the same
        piece of the neighbours' edge planes of the buffer it reads and, once computed, is pushed into the neighbours'
        halos of the buffer it wrote -- the transfer for the NEXT stage runs under the rest of this one."""
        V, cols = self._cols
        t, dt, blocks = self._step
        talk, work = self.mode != "compute", self.mode != "comm"
        if not self._primed:
            if talk:
                for a, e in cols:
                    self.eng.halo_push(0, (a, e, V))
            self._primed = True
        for stage in (1, 2, 3):
            b_in, b_out = self.eng.stage_io(stage)
            self.run_stage(stage, comp, use_obstacle, which_pass=1)
            self.finish_halos(b_in)
            for a, e in cols:
                if talk:
                    self.eng.halo_wait(b_in, 1)
                if work:
                    self.eng.stage_cols(stage, a, e, t, dt, blocks[stage - 1], comp if stage == 3 else L.COMP_NONE,
                                        use_obstacle and stage == 3)
                if talk:
                    self.eng.halo_push(b_out, (a, e, V))

    def two_pass(self):
        if self._overlap is None:
            o = bool(self.weno != "intended" and not self._no_overlap and getattr(self.eng, "is_split", lambda: False)())
            if not self._ready():
                return o
            self._overlap = o
        return self._overlap

    def step(self, t, t_end, factorCFL, comp=L.COMP_NONE, use_obstacle=False, maxStep=np.finfo(np.float64).max):
        """One CFL-limited TVD-RK3 step of the distributed field.  Returns (t_new, dt)."""
        dt = self.begin_step(t, t_end, factorCFL, maxStep)
        if self.overlapped() and self.fused():
            self._step_fused(comp, use_obstacle)
            return rk3_times(t, dt)[2], dt
        if self.overlapped() and self.pieces() is not None:
            self._step_pieces(comp, use_obstacle)
            return rk3_times(t, dt)[2], dt
        for stage in (1, 2, 3):
            b = self.eng.stage_io(stage)[0]
            if self.overlapped():
                reqs = self.post(b)
                self.run_stage(stage, comp, use_obstacle, which_pass=1)
                self.wait(reqs)
                self.finish_halos(b)
                self.run_stage(stage, comp, use_obstacle, which_pass=2)
            elif self.ranged() and (self.peer or hasattr(self.comm, "post")):
                t_, dt_, blocks = self._step
                reqs = self.post(b)
                if self.mode != "comm":
                    self.eng.stage_range(stage, GHOST, self.n0 - GHOST, t_, dt_, blocks[stage - 1], comp, use_obstacle)
                self.wait(reqs)
                self.finish_halos(b)
                if self.mode != "comm":
                    self.eng.stage_range(stage, 0, GHOST, t_, dt_, blocks[stage - 1], comp, use_obstacle)
                    self.eng.stage_range(stage, self.n0 - GHOST, self.n0, t_, dt_, blocks[stage - 1], comp, use_obstacle)
            else:
                if self.mode != "compute":
                    self.exchange(b)
                else:
                    self.finish_halos(b)
                self.run_stage(stage, comp, use_obstacle)
        return rk3_times(t, dt)[2], dt


class LocalWorld:
    """``world`` slabs driven in lock-step inside ONE process (all on one device): exercises exactly the halo,
    edge-fill and reduction code of the multi-process path, with tensor copies standing in for send/recv and an
    element-wise maximum over the slabs standing in for the max-allreduce."""

    poison_halos = False     # tests: NaN the halo planes a stage will receive before its pass 1 runs

    def __init__(self, schemeData, world, device=0, backend=None, engine_factory=None, transport="auto", pieces="auto",
                 fused="auto"):
        self.world = int(world)
        self.slabs = [SlabSolver(schemeData, device, backend, _LocalComm(self, r), engine_factory, transport, True, pieces,
                                 fused) for r in range(self.world)]
        self.peer = all(s.peer for s in self.slabs) and self.world > 1
        if self.peer:                       # contexts of one process attach by pointer (hj_halo_attach)
            descs = [s.eng.halo_export() for s in self.slabs]
            for s in self.slabs:
                s.attach_peers(descs)

    def upload(self, data, field=L.FIELD_STATE):
        for s in self.slabs:
            s.upload(np.ascontiguousarray(data[s.lo:s.hi]), field)       # (state_changed: pieces pushed ahead are dropped)

    def download(self):
        return np.concatenate([s.download() for s in self.slabs], axis=0)

    def _exchange(self, b):
        if self.peer:                       # every push is queued before the first wait: one stream serves all slabs
            for s in self.slabs:
                s.eng.halo_push(b)
            for s in self.slabs:
                s.eng.halo_wait(b)
            for s in self.slabs:
                s.finish_halos(b)
            return
        for s in self.slabs:
            sends, _ = s.halo_ops(b)
            for t, peer, tag in sends:
                _, _, rlo, rhi = self.slabs[peer]._faces(b)
                (rlo if tag == 0 else rhi).copy_(t)
        for s in self.slabs:
            s.finish_halos(b)

    def _step_fused(self, comp, use_obstacle):
        """SlabSolver._step_fused for every slab in lock-step on one stream (a wait only depends on signals queued in the
        previous stage, or when the state was primed)."""
        sides = 2 if any(s.hybrid() for s in self.slabs) else 3
        for s in self.slabs:
            s.eng.halo_set_fused(sides)
            if not s._primed:
                s.eng.halo_push(0)
                s._primed = True
        for stage in (1, 2, 3):
            b_in, b_out = self.slabs[0].eng.stage_io(stage)
            for s in self.slabs:
                t_, dt_, blocks = s._step
                s.eng.stage(stage, t_, dt_, blocks[stage - 1], comp, use_obstacle, which_pass=1)
                s.finish_halos(b_in)
            for s in self.slabs:
                t_, dt_, blocks = s._step
                s.eng.halo_wait(b_in, 1)
                s.eng.stage(stage, t_, dt_, blocks[stage - 1], comp, use_obstacle, which_pass=2)
                s.eng.halo_signal(b_out, sides)
                if sides != 3:
                    s.eng.halo_push(b_out, sides=3 & ~sides)

    def _step_pieces(self, comp, use_obstacle):
        """SlabSolver._step_pieces for every slab in lock-step on one stream: all pushes a wait depends on were queued in
        the previous stage (or when the state was primed), so no wait can block the stream forever."""
        V, cols = self.slabs[0]._cols
        for s in self.slabs:
            if not s._primed:
                for a, e in cols:
                    s.eng.halo_push(0, (a, e, V))
                s._primed = True
        for stage in (1, 2, 3):
            b_in, b_out = self.slabs[0].eng.stage_io(stage)
            for s in self.slabs:
                t_, dt_, blocks = s._step
                s.eng.stage(stage, t_, dt_, blocks[stage - 1], comp, use_obstacle, which_pass=1)
                s.finish_halos(b_in)
            for a, e in cols:
                for s in self.slabs:
                    t_, dt_, blocks = s._step
                    s.eng.halo_wait(b_in, 1)
                    s.eng.stage_cols(stage, a, e, t_, dt_, blocks[stage - 1], comp, use_obstacle)
                    s.eng.halo_push(b_out, (a, e, V))

    @staticmethod
    def _max_over(tensors):
        import torch
        m = torch.stack([x.clone() for x in tensors]).max(dim=0).values
        for x in tensors:
            x.copy_(m)

    def step(self, t, t_end, factorCFL, comp=L.COMP_NONE, use_obstacle=False, maxStep=np.finfo(np.float64).max):
        if not self.slabs[0].adapter.host_alpha and self.slabs[0]._alpha is None:
            amax = np.max(np.stack([s.local_alpha() for s in self.slabs]), axis=0)
            for s in self.slabs:
                s._alpha = [float(x) for x in amax]
        dts = [s.begin_step(t, t_end, factorCFL, maxStep) for s in self.slabs]
        assert all(d == dts[0] for d in dts), "dt must be identical on every rank"
        if self.peer and self.slabs[0].overlapped() and self.slabs[0].fused():
            self._step_fused(comp, use_obstacle)
            return rk3_times(t, dts[0])[2], dts[0]
        if self.peer and self.slabs[0].overlapped() and self.slabs[0].pieces() is not None:
            self._step_pieces(comp, use_obstacle)
            return rk3_times(t, dts[0])[2], dts[0]
        two = self.slabs[0].two_pass()        # product systems: pass 1 runs before the halos arrive (as under NCCL)
        ranged = all(s.ranged() for s in self.slabs)   # whole 3-D systems: interior planes before, edge planes after
        for stage in (1, 2, 3):
            b = self.slabs[0].eng.stage_io(stage)[0]
            if ranged:
                for s in self.slabs:
                    if self.poison_halos:
                        _, _, rlo, rhi = s._faces(b)
                        rlo.fill_(float("nan"))
                        rhi.fill_(float("nan"))
                    t_, dt_, blocks = s._step
                    s.eng.stage_range(stage, GHOST, s.n0 - GHOST, t_, dt_, blocks[stage - 1], comp, use_obstacle)
                self._exchange(b)
                for s in self.slabs:
                    t_, dt_, blocks = s._step
                    s.eng.stage_range(stage, 0, GHOST, t_, dt_, blocks[stage - 1], comp, use_obstacle)
                    s.eng.stage_range(stage, s.n0 - GHOST, s.n0, t_, dt_, blocks[stage - 1], comp, use_obstacle)
                continue
            if two:
                for s in self.slabs:
                    if self.poison_halos:
                        _, _, rlo, rhi = s._faces(b)
                        rlo.fill_(float("nan"))
                        rhi.fill_(float("nan"))
                    t_, dt_, blocks = s._step
                    s.eng.stage(stage, t_, dt_, blocks[stage - 1], comp, use_obstacle, which_pass=1)
            self._exchange(b)
            if self.slabs[0].weno == "intended":
                self._max_over([s.eng.eps_prepass(b) for s in self.slabs])
            for s in self.slabs:
                t_, dt_, blocks = s._step
                kw = {"which_pass": 2} if two else {}
                s.eng.stage(stage, t_, dt_, blocks[stage - 1], comp, use_obstacle, **kw)
        return rk3_times(t, dts[0])[2], dts[0]
