"""Python face of one ``hj_ctx`` (include/hjb200.h): grid + scheme + resident fields on one GPU.

Everything numerical happens in the CUDA library; this module only marshals the reference's host-side
objects (grid Bundles, numpy / torch arrays) into the C-ABI.  PyTorch is optional here: numpy arrays are
staged through the library's own device-memory helpers; torch CUDA tensors are passed by pointer.
"""
import ctypes as C
import sys
import weakref

import numpy as np

from . import _lib as L

__all__ = ["Engine", "engine_for_grid", "DeviceBuffer", "current_stream", "weno_mode_of", "clear_engine_cache"]

_WENO = {"as_shipped": L.WENO_AS_SHIPPED, "intended": L.WENO_INTENDED, "eno3a": L.SCHEME_ENO3A, "eno2": L.SCHEME_ENO2,
         L.WENO_AS_SHIPPED: L.WENO_AS_SHIPPED, L.WENO_INTENDED: L.WENO_INTENDED}
# schemeData.CoStateCalc (by name) -> derivative scheme of the context; the WENO5 names take schemeData.wenoMode
_COSTATE_SCHEME = {"upwindFirstENO3a": "eno3a", "upwindFirstENO3": "eno3a", "upwindFirstENO2": "eno2"}


def _torch():
    return sys.modules.get("torch")


def is_torch_tensor(x):
    t = _torch()
    return t is not None and isinstance(x, t.Tensor)


def current_stream(device=None):
    """The CUDA stream work is enqueued on: torch's current stream if torch is loaded, else the default stream."""
    t = _torch()
    if t is not None and t.cuda.is_available():
        return int(t.cuda.current_stream(device).cuda_stream)
    return 0


def weno_mode_of(scheme_data=None, default="as_shipped"):
    """``schemeData.wenoMode`` ('as_shipped' | 'intended'); the reference has no such switch, so the default is
    its shipped behaviour."""
    if scheme_data is not None:
        name = getattr(getattr(scheme_data, "CoStateCalc", None), "__name__", None)
        if name in _COSTATE_SCHEME:
            return _COSTATE_SCHEME[name]
    mode = getattr(scheme_data, "wenoMode", default) if scheme_data is not None else default
    if mode not in ("as_shipped", "intended"):
        raise ValueError("wenoMode must be 'as_shipped' or 'intended', got %r" % (mode,))
    return mode


class DeviceBuffer:
    """A plain cudaMalloc'ed fp64 array owned by the library (exposes ``__cuda_array_interface__``)."""

    def __init__(self, n, device=0):
        self.n, self.device = int(n), device
        p = C.c_void_p()
        L.check(L.load().hj_dev_alloc(device, self.n * 8, C.byref(p)))
        self.ptr = p.value
        self._fin = weakref.finalize(self, L.load().hj_dev_free, C.c_void_p(self.ptr))

    @property
    def __cuda_array_interface__(self):
        return {"shape": (self.n,), "typestr": "<f8", "data": (self.ptr, False), "version": 3, "strides": None}

    def from_host(self, a, stream=0):
        a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
        assert a.size == self.n
        L.check(L.load().hj_memcpy(self.ptr, a.ctypes.data, self.n * 8, 1, stream, 1))
        return self

    def to_host(self, stream=0):
        out = np.empty(self.n, dtype=np.float64)
        L.check(L.load().hj_memcpy(out.ctypes.data, self.ptr, self.n * 8, 2, stream, 1))
        return out

    def free(self):
        self._fin()


class _RawCuda:
    """``__cuda_array_interface__`` wrapper around memory the library owns (no ownership taken)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def grid_signature(grid):
    """(N, dx, bc kinds, towardZero, vs) of a reference-style grid Bundle, validated."""
    D = int(grid.dim)
    N = [int(x) for x in np.asarray(grid.N).reshape(-1)]
    dx = [float(x) for x in np.asarray(grid.dx).reshape(-1)]
    if len(N) != D or len(dx) != D:
        raise ValueError("grid.N / grid.dx do not agree with grid.dim")
    kinds, tz = [], []
    for d in range(D):
        fn = grid.bdry[d]
        name = getattr(fn, "__name__", None)
        if name == "addGhostPeriodic":
            kinds.append(L.BC_PERIODIC)
            tz.append(0)
        elif name == "addGhostExtrapolate":
            kinds.append(L.BC_EXTRAPOLATE)
            gd = grid.bdryData[d] if getattr(grid, "bdryData", None) is not None else None
            tz.append(1 if (gd is not None and getattr(gd, "towardZero", False)) else 0)
        else:
            raise NotImplementedError(
                "grid.bdry[%d] = %r has no device implementation (addGhostExtrapolate / addGhostPeriodic only; "
                "no CPU fallback)" % (d, fn))
    vs = [np.ascontiguousarray(np.asarray(grid.vs[d], dtype=np.float64).reshape(-1)) for d in range(D)]
    for d in range(D):
        if vs[d].size != N[d]:
            raise ValueError("Inconsistent grid size in dimension %d" % d)
    return D, N, dx, kinds, tz, vs


class Engine:
    """One hj_ctx.  ``slab=(lo, hi)`` restricts the context to planes [lo, hi) of dim 0 with stored halos."""

    def __init__(self, grid, weno="as_shipped", device=0, slab=None, backend=None):
        self.lib = L.load()
        D, N, dx, kinds, tz, vs = grid_signature(grid)
        self.D, self.N_global, self.dx, self.device = D, list(N), dx, device
        self.weno = weno
        self.slab = slab
        if slab is not None:
            lo, hi = slab
            N = [hi - lo] + N[1:]
            kinds = [L.BC_HALO] + kinds[1:]
            vs = [vs[0][lo:hi].copy()] + vs[1:]
            self.bc0_global = grid_signature(grid)[3][0]
            self.tz0_global = grid_signature(grid)[4][0]
        self.N, self.kinds, self.tz, self.vs = N, kinds, tz, vs
        self.shape = tuple(N)
        self.nodes = int(np.prod([float(n) for n in N]))
        h = C.c_void_p()
        L.check(self.lib.hj_create(C.byref(h), device, D, (C.c_int64 * D)(*N), (C.c_double * D)(*dx),
                                   (C.c_int * D)(*kinds), (C.c_int * D)(*tz), _WENO[weno]))
        self.h = h
        self._fin = weakref.finalize(self, self.lib.hj_destroy, h)
        for d in range(D):
            L.check(self.lib.hj_set_axis(self.h, d, vs[d].ctypes.data, vs[d].size))
        if backend is not None:
            L.check(self.lib.hj_set_backend(self.h, backend))
        self.adapter = None
        self._params = None
        self._tables_key = None
        self.nparams = 0

    # ------------------------------------------------------------------ system
    def set_system(self, system_id, params, tables=()):
        params = np.ascontiguousarray(params, dtype=np.float64).reshape(-1)
        key = tuple((slot, tab.tobytes()) for slot, tab in tables)
        if key != self._tables_key:
            for slot, tab in tables:
                tab = np.ascontiguousarray(tab, dtype=np.float64).reshape(-1)
                L.check(self.lib.hj_set_table(self.h, slot, tab.ctypes.data, tab.size))
            self._tables_key = key
        if self._params is None or self._params[0] != system_id or not np.array_equal(self._params[1], params):
            L.check(self.lib.hj_set_system(self.h, system_id, params.ctypes.data, params.size))
            self._params = (system_id, params.copy())
        self.nparams = params.size

    # ------------------------------------------------------------------ array marshalling
    def _flat(self, a, n=None):
        n = self.nodes if n is None else n
        if is_torch_tensor(a):
            t = _torch()
            if not a.is_cuda:
                raise ValueError("torch tensors must live on the GPU (pass numpy arrays for host data)")
            a = a.detach()
            if a.dtype != t.float64:
                a = a.to(t.float64)
            a = a.contiguous().reshape(-1)
            if a.numel() != n:
                raise ValueError("array has %d entries, grid has %d nodes" % (a.numel(), n))
            return a, a.data_ptr(), 0
        a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
        if a.size != n:
            raise ValueError("array has %d entries, grid has %d nodes" % (a.size, n))
        return a, a.ctypes.data, 1

    def _like(self, like, n, shape):
        """Output array of the same kind as ``like`` plus its pointer."""
        if is_torch_tensor(like):
            t = _torch()
            out = t.empty(n, dtype=t.float64, device=like.device)
            return out, out.data_ptr(), 0, lambda: out.reshape(shape)
        buf = DeviceBuffer(n, self.device)
        return buf, buf.ptr, 1, lambda: buf.to_host(self.stream()).reshape(shape)

    def _to_device(self, a, n=None):
        """Dense device copy (pointer) of a numpy array, or the tensor's own storage."""
        flat, ptr, is_host = self._flat(a, n)
        if not is_host:
            return flat, ptr
        buf = DeviceBuffer(flat.size, self.device).from_host(flat, self.stream())
        return buf, buf.ptr

    def stream(self):
        return current_stream(self.device)

    # ------------------------------------------------------------------ resident fields
    def upload(self, a, field=L.FIELD_STATE):
        flat, ptr, is_host = self._flat(a)
        L.check(self.lib.hj_upload(self.h, self.stream(), field, ptr, is_host))
        self._keep = flat

    def download(self, like=None, field=L.FIELD_STATE, shape=None, out=None):
        """The dense field; ``out``: a C-contiguous float64 numpy array (e.g. pinned) that receives it in place."""
        shape = (self.nodes, 1) if shape is None else shape
        if out is not None:
            o = np.asarray(out)
            assert o.dtype == np.float64 and o.flags.c_contiguous and o.size == self.nodes
            L.check(self.lib.hj_download(self.h, self.stream(), field, o.ctypes.data, 1))
            return o.reshape(shape)
        if is_torch_tensor(like):
            t = _torch()
            out = t.empty(self.nodes, dtype=t.float64, device=like.device)
            L.check(self.lib.hj_download(self.h, self.stream(), field, out.data_ptr(), 0))
            return out.reshape(shape)
        out = np.empty(self.nodes, dtype=np.float64)
        L.check(self.lib.hj_download(self.h, self.stream(), field, out.ctypes.data, 1))
        return out.reshape(shape)

    # ------------------------------------------------------------------ operators
    def deriv(self, data, dim):
        """upwindFirstWENO5a: (derivL, derivR), same kind/shape as ``data``."""
        keep, pin = self._to_device(data)
        shape = tuple(data.shape)
        oL, pL, _, getL = self._like(data, self.nodes, shape)
        oR, pR, _, getR = self._like(data, self.nodes, shape)
        L.check(self.lib.hj_deriv(self.h, self.stream(), pin, int(dim), pL, pR))
        return getL(), getR()

    def deriv_candidates(self, data, dim):
        """upwindFirstENO3aHelper: ([dL0, dL1, dL2], [dR0, dR1, dR2]), same kind / shape as ``data``."""
        keep, pin = self._to_device(data)
        shape = tuple(data.shape)
        o, po, _, get = self._like(data, 6 * self.nodes, (6,) + shape)
        L.check(self.lib.hj_deriv_candidates(self.h, self.stream(), pin, int(dim), po))
        a = get()
        return [a[0], a[1], a[2]], [a[3], a[4], a[5]]

    def add_ghost(self, data, dim, width):
        keep, pin = self._to_device(data)
        shape = list(self.shape)
        shape[dim] += 2 * width
        n = int(np.prod(shape))
        o, po, _, get = self._like(data, n, tuple(shape))
        L.check(self.lib.hj_add_ghost(self.h, self.stream(), pin, int(dim), int(width), po))
        return get()

    def rhs(self, t, y):
        """termLaxFriedrichs: (ydot (n,1), stepBound, reductions dict)."""
        keep, pin = self._to_device(y)
        o, po, _, get = self._like(y, self.nodes, (self.nodes, 1))
        sb = C.c_double()
        red = (C.c_double * (3 * self.D + 1))()
        L.check(self.lib.hj_rhs(self.h, self.stream(), float(t), pin, po, C.byref(sb), red))
        r = np.array(red[:])
        D = self.D
        return get(), sb.value, dict(alphaMax=r[:D], derivMin=r[D:2 * D], derivMax=r[2 * D:3 * D], nan=bool(r[3 * D]))

    def deriv_range(self, y=None, stage=1):
        """(derivMin, derivMax) of artificial_diss_glf.py:82-88 -- per dim the min / max over the grid of the upwind pair --
        of a dense array ``y`` (numpy or torch CUDA tensor), or of the resident buffer RK stage ``stage`` reads (hj_deriv_range)."""
        lo, hi = (C.c_double * self.D)(), (C.c_double * self.D)()
        keep, pin = (None, None) if y is None else self._to_device(y)
        L.check(self.lib.hj_deriv_range(self.h, self.stream(), pin, int(stage), lo, hi))
        return np.array(lo[:]), np.array(hi[:])

    def _ptr_list(self, arrays):
        """(keep-alive list, C array of D device pointers) for a list of D dense arrays."""
        if len(arrays) != self.D:
            raise ValueError("need one array per grid dimension (%d), got %d" % (self.D, len(arrays)))
        keep, ptrs = [], (C.c_void_p * self.D)()
        for d, a in enumerate(arrays):
            k, p = self._to_device(a)
            keep.append(k)
            ptrs[d] = p
        return keep, ptrs

    def ham(self, t, derivs):
        """hamFunc(t, data, derivC, schemeData) on dense arrays (hj_ham); same kind / shape as ``derivs[0]``."""
        keep, ptrs = self._ptr_list(derivs)
        o, po, _, get = self._like(derivs[0], self.nodes, tuple(derivs[0].shape))
        L.check(self.lib.hj_ham(self.h, self.stream(), float(t), ptrs, po))
        return get()

    def alpha(self, t, dim, like):
        """partialFunc(..., dim) as a dense array shaped like ``like`` (hj_alpha)."""
        o, po, _, get = self._like(like, self.nodes, tuple(like.shape))
        L.check(self.lib.hj_alpha(self.h, self.stream(), float(t), int(dim), po))
        return get()

    def diss_glf(self, t, derivL, derivR):
        """artificialDissipationGLF on dense arrays (hj_diss_glf): (diss, stepBound, reductions dict)."""
        kl, pl = self._ptr_list(derivL)
        kr, pr = self._ptr_list(derivR)
        o, po, _, get = self._like(derivL[0], self.nodes, tuple(derivL[0].shape))
        sb = C.c_double()
        red = (C.c_double * (3 * self.D + 1))()
        L.check(self.lib.hj_diss_glf(self.h, self.stream(), float(t), pl, pr, po, C.byref(sb), red))
        r = np.array(red[:])
        D = self.D
        return get(), sb.value, dict(alphaMax=r[:D], derivMin=r[D:2 * D], derivMax=r[2 * D:3 * D])

    def alpha_max(self, t=0.0):
        a = (C.c_double * self.D)()
        sb = C.c_double()
        L.check(self.lib.hj_alpha_max(self.h, self.stream(), float(t), a, C.byref(sb)))
        return np.array(a[:]), sb.value

    def step(self, t, dt, stage_params=None, comp=L.COMP_NONE, use_obstacle=False, want_reduce=False):
        p = None
        if stage_params is not None:
            sp = np.ascontiguousarray(stage_params, dtype=np.float64).reshape(-1)
            assert sp.size == 3 * self.nparams
            p = sp.ctypes.data
            self._sp_keep = sp
        L.check(self.lib.hj_step(self.h, self.stream(), float(t), float(dt), p, int(comp), int(bool(use_obstacle)),
                                 int(bool(want_reduce))))

    def stage(self, stage, t, dt, params=None, comp=L.COMP_NONE, use_obstacle=False, want_reduce=False, which_pass=0):
        """One RK stage.  which_pass = 1 / 2: only the first / second kernel of a product system's stage on the
        dimension-split path (pass 1 reads no dim-0 halo: a slab job runs it under its halo exchange)."""
        p = None
        if params is not None:
            sp = np.ascontiguousarray(params, dtype=np.float64).reshape(-1)
            p = sp.ctypes.data
            self._sp_keep = sp
        if which_pass:
            L.check(self.lib.hj_stage_pass(self.h, self.stream(), int(stage), int(which_pass), float(t), float(dt), p,
                                           int(comp), int(bool(use_obstacle)), int(bool(want_reduce))))
        else:
            L.check(self.lib.hj_stage(self.h, self.stream(), int(stage), float(t), float(dt), p, int(comp),
                                      int(bool(use_obstacle)), int(bool(want_reduce))))

    def stage_range(self, stage, z_begin, z_end, t, dt, params=None, comp=L.COMP_NONE, use_obstacle=False, want_reduce=0):
        """Stage ``stage`` on planes [z_begin, z_end) of the marched dim only (whole systems, plane-ring backend)."""
        p = None
        if params is not None:
            sp = np.ascontiguousarray(params, dtype=np.float64).reshape(-1)
            p = sp.ctypes.data
            self._sp_keep = sp
        L.check(self.lib.hj_stage_range(self.h, self.stream(), int(stage), int(z_begin), int(z_end), float(t), float(dt),
                                        p, int(comp), int(bool(use_obstacle)), int(want_reduce)))

    def split_cols(self):
        """(length of the flattened trailing-dims axis, column quantum) of a product system on the split path."""
        v, q = C.c_int64(), C.c_int()
        L.check(self.lib.hj_split_cols(self.h, C.byref(v), C.byref(q)))
        return v.value, q.value

    def stage_cols(self, stage, col_begin, col_end, t, dt, params=None, comp=L.COMP_NONE, use_obstacle=False, want_reduce=0):
        """Pass 2 of ``stage`` on columns [col_begin, col_end) of the trailing-dims axis only (hj_stage_pass_cols)."""
        p = None
        if params is not None:
            sp = np.ascontiguousarray(params, dtype=np.float64).reshape(-1)
            p = sp.ctypes.data
            self._sp_keep = sp
        L.check(self.lib.hj_stage_pass_cols(self.h, self.stream(), int(stage), int(col_begin), int(col_end), float(t),
                                            float(dt), p, int(comp), int(bool(use_obstacle)), int(want_reduce)))

    def supports_range(self):
        """True if hj_stage_range can advance this context (probed with an empty-range call that must fail as INVALID,
        not as UNSUPPORTED)."""
        rc = self.lib.hj_stage_range(self.h, self.stream(), 1, 0, 0, 0.0, 0.0, None, 0, 0, 0)
        return rc == L.HJ_ERR_INVALID

    def is_split(self):
        """True if this context advances its (product) system as two kernels per stage (system and state set)."""
        return bool(self.lib.hj_is_split(self.h))

    def snapshot(self):
        """Device copy of the resident state (the frame the next ``change()`` compares with)."""
        L.check(self.lib.hj_snapshot(self.h, self.stream()))

    def change(self):
        """(max |state - snapshot|, has_nan) from one device reduction (hji_solver.py:661-672, :544)."""
        m, n = C.c_double(), C.c_int()
        L.check(self.lib.hj_change(self.h, self.stream(), C.byref(m), C.byref(n)))
        return m.value, bool(n.value)

    def discount(self, gamma, mode=0, take_max=False, max_val=0.0):
        L.check(self.lib.hj_discount(self.h, self.stream(), float(gamma), int(mode), int(bool(take_max)), float(max_val)))

    def mask_obstacle(self):
        """y = max(y, -obstacle) as its own pass (normally fused into stage 3)."""
        self.discount(1.0, 2)

    def set_restrict(self, sign):
        """termRestrictUpdate: +1 -> ydot = max(ydot, 0); -1 -> min(ydot, 0); 0 -> off."""
        L.check(self.lib.hj_set_restrict(self.h, int(sign)))

    def set_pipeline_planes(self, planes):
        """Chunk height of the pipelined host-buffer step (hj_ode_cfl3_step); 0 = default.  Results do not depend on it."""
        L.check(self.lib.hj_set_pipeline_planes(self.h, int(planes)))

    def step_rk2(self, t, dt, stage_params=None, comp=L.COMP_NONE, use_obstacle=False, want_reduce=False):
        p = None
        if stage_params is not None:
            sp = np.ascontiguousarray(stage_params, dtype=np.float64).reshape(-1)
            assert sp.size == 2 * self.nparams
            p = sp.ctypes.data
            self._sp_keep = sp
        L.check(self.lib.hj_step_rk2(self.h, self.stream(), float(t), float(dt), p, int(comp), int(bool(use_obstacle)),
                                     int(bool(want_reduce))))

    def step_reductions(self):
        n = 3 * self.D + 1
        buf = (C.c_double * (3 * n))()
        L.check(self.lib.hj_step_reductions(self.h, self.stream(), buf))
        r = np.array(buf[:]).reshape(3, n)
        D = self.D
        return [dict(alphaMax=x[:D], derivMin=x[D:2 * D], derivMax=x[2 * D:3 * D], nan=bool(x[3 * D])) for x in r]

    def ode_cfl3_single(self, t, t_end, factor_cfl, max_step, y, comp=L.COMP_NONE, use_obstacle=False):
        """hj_ode_cfl3_single on a numpy array (in place) or a torch CUDA tensor (in place)."""
        flat, ptr, is_host = self._flat(y)
        tn, dt = C.c_double(), C.c_double()
        L.check(self.lib.hj_ode_cfl3_single(self.h, self.stream(), float(t), float(t_end), float(factor_cfl),
                                            float(max_step), ptr, is_host, int(comp), int(bool(use_obstacle)),
                                            C.byref(tn), C.byref(dt)))
        return tn.value, flat, dt.value

    def pinned_out(self):
        """Next of this engine's three pinned host arrays (``nodes`` doubles each), handed out round-robin: what the
        host-buffer odeCFL3 path returns its ``y`` in.  An array stays untouched until three further calls -- long
        enough for the driver's ``yLast`` / ``y`` pattern (hji_solver.py:538-599) -- copy it to keep it longer."""
        ring = self.__dict__.setdefault("_pin_ring", [])
        if len(ring) < 3:
            p = C.c_void_p()
            L.check(self.lib.hj_host_alloc(self.nodes * 8, C.byref(p)))
            arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(self.nodes,))
            weakref.finalize(self, self.lib.hj_host_free, p)
            ring.append(arr)
            self._pin_next = 0
            return arr
        self._pin_next = (self._pin_next + 1) % 3
        return ring[self._pin_next]

    def ode_cfl3_step(self, t, t_end, factor_cfl, max_step, y_in, y_out, comp=L.COMP_NONE, use_obstacle=False):
        """hj_ode_cfl3_step on host numpy arrays (or torch CUDA tensors): reads y_in, writes y_out.  (t_new, dt)."""
        fin, pin, host_in = self._flat(y_in)
        fout, pout, host_out = self._flat(y_out)
        assert host_in == host_out, "y_in and y_out must both be host arrays or both be CUDA tensors"
        tn, dt = C.c_double(), C.c_double()
        L.check(self.lib.hj_ode_cfl3_step(self.h, self.stream(), float(t), float(t_end), float(factor_cfl),
                                          float(max_step), pin, pout, host_in, int(comp), int(bool(use_obstacle)),
                                          C.byref(tn), C.byref(dt)))
        return tn.value, dt.value

    def sync(self):
        L.check(self.lib.hj_stream_sync(self.stream()))

    def set_backend(self, backend):
        L.check(self.lib.hj_set_backend(self.h, backend))

    def buffer_ptr(self, which):
        p = C.c_void_p()
        L.check(self.lib.hj_state_ptr(self.h, which, C.byref(p)))
        return p.value

    @property
    def plane_elems(self):
        return int(self.lib.hj_plane_elems(self.h))

    @property
    def field_elems(self):
        return int(self.lib.hj_field_elems(self.h))

    # ------------------------------------------------------------------ slab (multi-GPU) support
    def buffer_tensor(self, which):
        """Zero-copy torch view (1-D, field_elems doubles, halo planes included) of RK buffer ``which``."""
        t = _torch()
        if t is None:
            import torch as t
        view = _RawCuda(self.buffer_ptr(which), self.field_elems, "<f8")
        return t.as_tensor(view, device=t.device("cuda", self.device))

    def fill_edge_halo(self, which, side):
        L.check(self.lib.hj_fill_edge_halo(self.h, self.stream(), int(which), int(side)))

    # peer-memory halos (hj_halo_*): descriptors are plain bytes the caller moves between ranks
    def halo_export(self):
        buf = C.create_string_buffer(L.HJ_HALO_DESC_BYTES)
        L.check(self.lib.hj_halo_export(self.h, buf))
        return bytes(buf.raw)

    def halo_attach(self, lower, upper):
        """``lower`` / ``upper``: the neighbours' ``halo_export()`` bytes, or None where the slab has no neighbour."""
        keep = [C.create_string_buffer(d, L.HJ_HALO_DESC_BYTES) if d is not None else None for d in (lower, upper)]
        L.check(self.lib.hj_halo_attach(self.h, *[C.cast(k, C.c_void_p) if k is not None else None for k in keep]))

    def halo_push(self, which, cols=None, sides=3):
        """Push my edge planes of RK buffer ``which`` into the neighbours' halos (``sides``: bit 0 lower, bit 1 upper),
        ordered behind the current stream.  ``cols`` = (begin, end, row_len): only those columns of every row."""
        b, e, n = cols if cols is not None else (0, 0, 0)
        L.check(self.lib.hj_halo_push(self.h, self.stream(), int(which), int(sides), int(b), int(e), int(n)))

    def halo_set_fused(self, sides=3):
        """Pass 2 of the split path stores its edge planes into the halos of the neighbours in ``sides`` (bit 0 lower,
        bit 1 upper; 0 = off) itself (hj_halo_set_fused)."""
        L.check(self.lib.hj_halo_set_fused(self.h, int(sides)))

    def halo_signal(self, which, sides=3):
        L.check(self.lib.hj_halo_signal(self.h, self.stream(), int(which), int(sides)))

    def halo_wait(self, which, npush=1):
        L.check(self.lib.hj_halo_wait(self.h, self.stream(), int(which), int(npush)))

    def halo_detach(self):
        L.check(self.lib.hj_halo_detach(self.h))

    def eps_prepass(self, which):
        """intended WENO: per-dim raw max(D1^2) of buffer ``which`` -> torch int64 view (D entries, ordered
        encoding: larger value <=> larger int64) that a slab job max-allreduces before the stage."""
        t = _torch()
        if t is None:
            import torch as t
        p = C.c_void_p()
        L.check(self.lib.hj_eps_prepass(self.h, self.stream(), int(which), C.byref(p)))
        return t.as_tensor(_RawCuda(p.value, self.D, "<i8"), device=t.device("cuda", self.device))

    def stage_io(self, stage):
        a, b = C.c_int(), C.c_int()
        L.check(self.lib.hj_stage_io(self.h, stage, C.byref(a), C.byref(b)))
        return a.value, b.value

    def close(self):
        """Destroy the context now.  The handle is dropped, so a later call fails in ctypes (NULL ctx ->
        HJ_ERR_INVALID) instead of touching freed memory."""
        self._fin()
        self.h = None


# ---------------------------------------------------------------------- engine cache
_CACHE = {}
_CACHE_MAX = 4


def clear_engine_cache():
    """Drop the cache's references.  Contexts are destroyed when their last holder lets go (weakref.finalize on the
    Engine), never under a caller that still holds one."""
    _CACHE.clear()


def engine_for_grid(grid, weno="as_shipped", device=None):
    """Engines are cached by grid *value* (N, dx, BCs, vs), so the shallow grid copies the reference makes
    (term_lax_friedrich.py:91) land on the same context."""
    if device is None:
        t = _torch()
        device = t.cuda.current_device() if (t is not None and t.cuda.is_available()) else 0
    D, N, dx, kinds, tz, vs = grid_signature(grid)
    key = (device, weno, tuple(N), tuple(dx), tuple(kinds), tuple(tz), tuple(v.tobytes() for v in vs))
    eng = _CACHE.get(key)
    if eng is None:
        while len(_CACHE) >= _CACHE_MAX:
            _CACHE.pop(next(iter(_CACHE)))          # not closed: other holders may still use it (see clear_engine_cache)
        eng = Engine(grid, weno, device)
        _CACHE[key] = eng
    return eng
