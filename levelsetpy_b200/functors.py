"""schemeData.hamFunc / partialFunc  ->  registered device functor + parameter block.

``resolve(hamFunc, partialFunc, grid)`` accepts bound methods of this package's system classes *or* of the
reference's own DynamicalSystems classes (recognised by class name + attributes).  Anything else raises: there
is no CPU fallback and no generic Python-callable path on the device.

Parameter block layouts are those documented in csrc/hj_systems.cuh / INTEGRATION.md.
"""
import numpy as np

from . import _lib as L

__all__ = ["resolve", "Adapter"]


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


def resolve(ham_func, partial_func, grid=None):
    """Adapter for a (hamFunc, partialFunc) pair; raises NotImplementedError for unregistered callables
    (e.g. genericHam/genericPartial, Hamiltonians/generic_ham.py:5) and ValueError for mismatched pairs."""
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
