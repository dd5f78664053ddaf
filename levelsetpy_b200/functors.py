"""schemeData.hamFunc / partialFunc  ->  registered device functor + parameter block.

``resolve(hamFunc, partialFunc, grid)`` accepts bound methods of this package's system classes *or* of the
reference's own DynamicalSystems classes (recognised by class name + attributes).  Anything else raises: there
is no CPU fallback and no generic Python-callable path on the device.

Parameter block layouts are those documented in csrc/hj_systems.cuh / INTEGRATION.md.
"""
import numpy as np

from . import _lib as L

__all__ = ["resolve", "Adapter", "generic_adapter"]


def _scalar(x, what):
    a = np.asarray(x, dtype=np.float64).reshape(-1)
    if a.size != 1:
        raise NotImplementedError("%s must be a scalar for the device functor" % what)
    return float(a[0])


def _vs(grid, d):
    return np.ascontiguousarray(np.asarray(grid.vs[d], dtype=np.float64).reshape(-1))


class Adapter:
    """system_id + per-RHS parameter blocks for one hamFunc/partialFunc pair."""

    system_id = 0
    time_varying = False       # parameter block changes from one RHS evaluation to the next (Flock)
    host_alpha = False         # alpha_d are host scalars inside the block (Flock)
    dynamic = False            # the block depends on the derivative range of the field (genericPartial)

    def tables(self, grid):
        return []

    def block(self):
        """Parameter block for the NEXT RHS evaluation (may mutate the system, like the reference's hamFunc)."""
        raise NotImplementedError

    def alphas(self, block):
        return None


class _DubinsRel(Adapter):
    system_id = L.SYS_DUBINS_REL
    ndim = 3

    def __init__(self, owner, base=0):
        self.o, self.base = owner, base

    def block(self):
        o = self.o
        return np.array([_scalar(o.v_e, "v_e"), _scalar(o.v_p, "v_p"), _scalar(o.w(1), "w(1)"),
                         _scalar(o.w_e, "w_e"), _scalar(o.w_p, "w_p")])

    def tables(self, grid):
        x3 = _vs(grid, self.base + 2)
        # host numpy trig == what the reference feeds its Hamiltonian (dubins_relative.py:81-82)
        return [np.cos(x3), np.sin(x3)]


class _DoubleInt(Adapter):
    system_id = L.SYS_DOUBLE_INT
    ndim = 2

    def __init__(self, owner, base=0):
        self.o, self.base = owner, base

    def block(self):
        return np.array([_scalar(self.o.control_law, "u_bound")])


class _Product(Adapter):
    def __init__(self, owner):
        subs = [_adapter_for_owner(s, "hamiltonian", "dissipation") for s in owner.subsystems]
        kinds = [type(s) for s in subs]
        if kinds == [_DubinsRel, _DubinsRel]:
            self.system_id = L.SYS_DUBINS_REL_PAIR
        elif kinds == [_DoubleInt, _DoubleInt]:
            self.system_id = L.SYS_DOUBLE_INT_PAIR
        else:
            raise NotImplementedError("no compiled product functor for %s" % [k.__name__ for k in kinds])
        base = 0
        for s in subs:
            s.base = base
            base += s.ndim
        self.subs = subs

    def block(self):
        return np.concatenate([s.block() for s in self.subs])

    def tables(self, grid):
        out = []
        for s in self.subs:
            out += s.tables(grid)
        return out


class _Flock(Adapter):
    """Flock (min over birds) or a lone Bird.  bird.py:266-273 (abs), :305-316 (attacked), :339-344, :367-372."""
    system_id = L.SYS_FLOCK
    time_varying = True
    host_alpha = True

    def __init__(self, owner, mode):
        self.o, self.mode = owner, mode      # mode: 'flock' | 'bird' (attacked form) | 'bird_abs'

    @staticmethod
    def _abs_coeffs(b):
        th = float(np.asarray(b.cur_state, dtype=np.float64)[2, 0])
        return [-np.cos(th), -np.sin(th), -float(b.w_e)]

    @staticmethod
    def _abs_alpha(b):
        cs = np.asarray(b.cur_state, dtype=np.float64)
        w_low = min([float(np.asarray(n.cur_state)[2, 0]) for n in b.neighbors])
        return [float(np.abs(b.v_p * np.cos(cs[2, 0]))), float(np.abs(b.v_e * np.sin(cs[2, 0]))), w_low]

    @staticmethod
    def _att(b):
        cs = np.asarray(b.cur_state, dtype=np.float64)
        W = max([n.w_e for n in b.neighbors])
        a1 = b.v_e - b.v_p * np.cos(cs[2, 0])
        a2 = b.v_p * np.sin(cs[2, 0])
        w_up = max([float(np.asarray(n.cur_state)[2, 0]) for n in b.neighbors])
        al = [float(np.abs(b.v_e - b.v_p * np.cos(cs[2, 0])) + np.abs(w_up * cs[1, 0])),
              float(np.abs(b.v_p * np.sin(cs[2, 0])) + np.abs(w_up * cs[0, 0])),
              float(b.w_p + w_up)]
        return [float(W), float(a1), float(a2), float(cs[0, 0]), float(cs[1, 0])], al

    def block(self, mutate=True):
        """``mutate=False``: the block of the flock as it stands (partialFunc alone does not re-run the bookkeeping,
        flock.py:236-258)."""
        o = self.o
        if self.mode == "flock":
            if mutate:
                o._housekeeping()                          # flock.py:213 -- mutates headings every RHS
                o.attacked_idx = 0                         # flock.py:216
            if len(o.vehicles) == 2:
                raise IndexError("list index out of range")    # shapeUnion 2-shape bug, shape_ops.py:37
            att, others = o.vehicles[0], list(o.vehicles[1:])
        elif self.mode == "bird":
            att, others = o, []
        else:
            att, others = None, [o]
        K = len(others)
        if L.HJ_MAX_PARAMS < 10 + 3 * K:
            raise NotImplementedError("flock of %d birds exceeds the device parameter block" % (K + 1))
        hdr = [float(K), 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]
        alphas = [self._abs_alpha(b) for b in others]
        if att is not None:
            a, al = self._att(att)
            hdr = [float(K), 1.0] + a
            alphas.append(al)
        amax = [max(a[d] for a in alphas) for d in range(3)]       # flock.py:257 (scalarised)
        coeffs = []
        for b in others:
            coeffs += self._abs_coeffs(b)
        return np.array(hdr + amax + coeffs, dtype=np.float64)

    def alphas(self, block):
        return [float(block[7]), float(block[8]), float(block[9])]


class _GenericDubinsCar(Adapter):
    """genericHam / genericPartial over a ``DubinsCar``-shaped dynSys (Hamiltonians/generic_ham.py, generic_partial.py).
    Block (csrc/hj_systems.cuh: GenericF<DubinsCarDyn>): uSign dSign hamSign | uU uL | dU[3] dL[3] | speed wMax dMax[3]."""
    system_id = L.SYS_GENERIC_DUBINS_CAR
    ndim = 3
    time_varying = True        # keeps the host-buffer pipeline (one fixed block per step) away
    dynamic = True

    def __init__(self, scheme):
        self.sd = scheme
        self.dyn = scheme.dynSys
        for f in ("uIn", "dIn", "deriv", "side"):
            if f in scheme.__dict__ and getattr(scheme, f) is not None:
                raise NotImplementedError("schemeData.%s (generic_ham.py:21-43) has no device path" % f)
        if "TIdim" in self.dyn.__dict__ and self.dyn.TIdim:
            raise NotImplementedError("dynSys.TIdim (generic_ham.py:49-51) has no device path")
        if "partialFunc" in self.dyn.__dict__:
            raise NotImplementedError("dynSys.partialFunc (generic_partial.py:10-12) is an arbitrary Python callable")

    def tables(self, grid):
        x2 = _vs(grid, 2)
        return [np.cos(x2), np.sin(x2)]            # host numpy trig == what dynSys.dynamics sees (np.cos(x[2]))

    def modes(self):
        sd = self.sd
        u = sd.uMode if "uMode" in sd.__dict__ else "min"          # generic_ham.py:11-12
        d = sd.dMode if "dMode" in sd.__dict__ else "max"          # :14-15 (genericHam runs first in termLaxFriedrichs)
        tm = sd.tMode if "tMode" in sd.__dict__ else "backward"    # :17-18
        for m, what in ((u, "uMode"), (d, "dMode")):
            if m not in ("min", "max"):
                raise ValueError("Unknown %s!" % what)
        return u, d, tm

    def block_for_range(self, deriv_min, deriv_max, t=0.0, d_mode=None):
        """Parameter block for one RHS evaluation.  deriv_min / deriv_max: the D grid-wide scalars of
        artificial_diss_glf.py:82-88 (None: Hamiltonian only, alpha inputs zeroed)."""
        dyn = self.dyn
        u_mode, dm, t_mode = self.modes()
        d_mode = d_mode or dm
        if deriv_min is None:
            uU = uL = 0.0
            dU = dL = [0.0, 0.0, 0.0]
        else:
            lo = [float(np.asarray(v).reshape(-1)[0]) for v in deriv_min]
            hi = [float(np.asarray(v).reshape(-1)[0]) for v in deriv_max]
            # generic_partial.py:28-40: the dynSys's own methods, on the scalar range
            uU = dyn.get_opt_u(t, hi, u_mode, None)
            uL = dyn.get_opt_u(t, lo, u_mode, None)
            dU = dyn.get_opt_v(t, hi, d_mode, None)
            dL = dyn.get_opt_v(t, lo, d_mode, None)
        return np.array([1.0 if u_mode == "max" else -1.0, 1.0 if d_mode == "max" else -1.0,
                         -1.0 if t_mode == "backward" else 1.0, _scalar(uU, "uU"), _scalar(uL, "uL")]
                        + [_scalar(v, "dU") for v in dU] + [_scalar(v, "dL") for v in dL]
                        + [_scalar(dyn.speed, "speed"), _scalar(dyn.wMax, "wMax")] + [_scalar(v, "dMax") for v in dyn.dMax])

    def block(self):
        raise NotImplementedError("a generic dynSys needs the derivative range of the field: block_for_range()")


_GENERIC = {"DubinsCar": _GenericDubinsCar}


def generic_adapter(scheme, set_defaults=None):
    """Adapter for schemeData.hamFunc = genericHam / partialFunc = genericPartial: schemeData.dynSys must be an instance of
    a dynSys class with a compiled device functor (recognised by class name + attributes: DubinsCar).
    ``set_defaults``: 'ham' / 'partial' write the reference's mode defaults into the bundle (generic_ham.py:11-18,
    generic_partial.py:16-20) as a standalone call of that function does."""
    if "dynSys" not in scheme.__dict__:
        raise AttributeError("'Bundle' object has no attribute 'dynSys'")          # generic_ham.py:8
    if set_defaults:
        if "uMode" not in scheme.__dict__:
            scheme.uMode = "min"
        if "dMode" not in scheme.__dict__:
            scheme.dMode = "max" if set_defaults == "ham" else "min"
        if set_defaults == "ham" and "tMode" not in scheme.__dict__:
            scheme.tMode = "backward"
    name = type(scheme.dynSys).__name__
    if name not in _GENERIC:
        raise NotImplementedError(
            "schemeData.dynSys is a %s: genericHam / genericPartial call its Python methods get_opt_u / get_opt_v / "
            "dynamics on full-grid arrays, which cannot run inside the fused kernel; registered device dynSys classes: %s "
            "(csrc/hj_systems.cuh shows how to add one); there is no CPU fallback" % (name, ", ".join(sorted(_GENERIC))))
    ad = scheme.__dict__.get("_hjb200_generic")
    if ad is None or ad.dyn is not scheme.dynSys:
        ad = _GENERIC[name](scheme)
        try:
            scheme._hjb200_generic = ad
        except Exception:
            pass
    ad.sd = scheme
    return ad


def _adapter_for_owner(owner, ham_name, part_name):
    name = type(owner).__name__
    pair = (ham_name, part_name)
    if name == "DubinsVehicleRel" and pair == ("hamiltonian", "dissipation"):
        return _DubinsRel(owner)
    if name == "DoubleIntegrator" and pair == ("hamiltonian", "dissipation"):
        return _DoubleInt(owner)
    if name == "ProductSystem" and pair == ("hamiltonian", "dissipation"):
        return _Product(owner)
    if name == "Flock" and pair == ("hamiltonian", "dissipation"):
        return _Flock(owner, "flock")
    if name == "Bird" and pair == ("hamiltonian", "dissipation"):
        return _Flock(owner, "bird")
    if name == "Bird" and pair == ("hamiltonian_abs", "dissipation_abs"):
        return _Flock(owner, "bird_abs")
    raise NotImplementedError(
        "hamFunc/partialFunc = %s.%s/%s has no registered device functor (registered: DubinsVehicleRel, "
        "DoubleIntegrator, Bird, Flock, ProductSystem of those); arbitrary Python callables cannot run inside the "
        "fused kernel and there is no CPU fallback" % (name, ham_name, part_name))


def resolve(ham_func, partial_func, grid=None, scheme=None):
    """Adapter for a (hamFunc, partialFunc) pair; raises NotImplementedError for unregistered callables and ValueError
    for mismatched pairs.  genericHam / genericPartial (Hamiltonians/generic_ham.py:5) resolve through
    ``scheme.dynSys`` (generic_adapter)."""
    names = (getattr(ham_func, "__name__", None), getattr(partial_func, "__name__", None))
    if "genericHam" in names or "genericPartial" in names:
        if names != ("genericHam", "genericPartial"):
            raise ValueError("genericHam and genericPartial come as a pair (hji_solver.py:413-415)")
        if scheme is None:
            raise NotImplementedError("genericHam / genericPartial need schemeData (dynSys, uMode, dMode, tMode)")
        ad = generic_adapter(scheme)
        if grid is not None and ad.ndim != grid.dim:
            raise ValueError("system is %d-D but the grid is %d-D" % (ad.ndim, grid.dim))
        return ad
    ho, po = getattr(ham_func, "__self__", None), getattr(partial_func, "__self__", None)
    if ho is None or po is None:
        raise NotImplementedError(
            "hamFunc=%r / partialFunc=%r are not bound methods of a registered DynamicalSystem; arbitrary Python "
            "callables cannot run inside the fused kernel and there is no CPU fallback" % (ham_func, partial_func))
    if ho is not po:
        raise ValueError("hamFunc and partialFunc belong to different system objects")
    ad = _adapter_for_owner(ho, ham_func.__name__, partial_func.__name__)
    if grid is not None:
        nd = {L.SYS_DUBINS_REL: 3, L.SYS_DOUBLE_INT: 2, L.SYS_FLOCK: 3, L.SYS_DUBINS_REL_PAIR: 6,
              L.SYS_DOUBLE_INT_PAIR: 4}[ad.system_id]
        if nd != grid.dim:
            raise ValueError("system is %d-D but the grid is %d-D" % (nd, grid.dim))
    return ad
