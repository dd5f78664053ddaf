"""CPU ORACLE (systems) -- TEST INFRASTRUCTURE ONLY; never imported by the product path.

numpy fp64 restatements of the reference's DynamicalSystems ``hamiltonian`` / ``dissipation``
pairs (the hamFunc / partialFunc callables of term_lax_friedrich.py:111 and
artificial_diss_glf.py:98), each following the reference file:line it names, plus the
``ProductSystem`` construction SURVEY.md section 8(d) defines for the 4-D and 6-D configs
(the reference has no system above 3-D).

Unlike the reference these read coordinates from the 1-D ``grid.vs`` axes (broadcast), never
from a dense ``grid.xs`` -- same numbers, and usable at 41**6.
"""
from __future__ import annotations

import numpy as np


def _axis(grid, d, ndim=None, offset=0):
    """grid.vs[d] broadcast along axis (d) of an ``ndim``-dimensional array."""
    ndim = grid.dim if ndim is None else ndim
    v = np.asarray(grid.vs[d], dtype=np.float64).reshape(-1)
    shape = [1] * ndim
    shape[d] = v.size
    return v.reshape(shape)


class SubGrid:
    """View of dims [lo, lo+n) of a parent grid that still broadcasts against the parent's arrays."""

    def __init__(self, parent, lo, n):
        self.parent, self.lo, self.dim_local = parent, lo, n
        self.dim = parent.dim
        self.vs = parent.vs
        self.dx = parent.dx


class DubinsVehicleRel:
    """DynamicalSystems/dubins_relative.py:12-111 (scalar bounds: v_e=v_p=u_bound, w_e=w_p=w_bound, :44-61)."""

    ndim = 3

    def __init__(self, grid, u_bound=5, w_bound=5, dims=(0, 1, 2)):
        self.grid, self.dims = grid, tuple(dims)
        self.v_e = 1 * u_bound
        self.v_p = 1 * u_bound
        self.w_e = 1 * w_bound
        self.w_p = 1 * w_bound
        self.w1 = 1 * w_bound                                              # self.w(1)

    def _x(self, k):
        return _axis(self.grid, self.dims[k])

    def hamiltonian(self, t, data, derivs, sd=None):
        d0, d1, d2 = self.dims
        p1, p2, p3 = derivs[d0], derivs[d1], derivs[d2]
        x1, x2, x3 = self._x(0), self._x(1), self._x(2)
        p1_coeff = self.v_e - self.v_p * np.cos(x3)                        # :81
        p2_coeff = self.v_p * np.sin(x3)                                   # :82
        return (p1 * p1_coeff - p2 * p2_coeff
                - self.w1 * np.abs(p1 * x2 - p2 * x1 - p3) + self.w1 * np.abs(p3))   # :84-88

    def dissipation(self, t, data, dmin, dmax, sd, dim):
        k = self.dims.index(dim)
        x1, x2, x3 = self._x(0), self._x(1), self._x(2)
        shape = np.broadcast_shapes(x1.shape, x2.shape, x3.shape)
        if k == 0:
            return np.broadcast_to(np.abs(self.v_e - self.v_p * np.cos(x3)) + np.abs(self.w1 * x2), shape)   # :106-107
        if k == 1:
            return np.broadcast_to(np.abs(self.v_p * np.sin(x3)) + np.abs(self.w1 * x1), shape)              # :108-109
        return self.w_e + self.w_p                                         # :110-111


class DoubleIntegrator:
    """DynamicalSystems/double_integrator.py:9-89."""

    ndim = 2

    def __init__(self, grid, u_bound=1, dims=(0, 1)):
        self.grid, self.dims = grid, tuple(dims)
        self.control_law = u_bound

    def hamiltonian(self, t, data, derivs, sd=None):
        x2 = _axis(self.grid, self.dims[1])
        return -(derivs[self.dims[0]] * x2 - np.abs(derivs[self.dims[1]]) * self.control_law)   # :71-74

    def dissipation(self, t, data, dmin, dmax, sd, dim):
        k = self.dims.index(dim)
        if k == 0:
            x2 = _axis(self.grid, self.dims[1])
            shape = tuple(int(n) for n in np.asarray(self.grid.N).reshape(-1))
            return np.broadcast_to(np.abs(x2), shape)                      # :85
        return np.abs(self.control_law)                                    # :86


class ProductSystem:
    """H(x,p) = sum_k H_k(x_k, p_k) over sub-systems acting on disjoint dim blocks; alpha_d comes from the
    sub-system that owns d.  (not in the reference.)"""

    def __init__(self, subsystems):
        self.subsystems = list(subsystems)

    def hamiltonian(self, t, data, derivs, sd=None):
        h = self.subsystems[0].hamiltonian(t, data, derivs, sd)
        for s in self.subsystems[1:]:
            h = h + s.hamiltonian(t, data, derivs, sd)
        return h

    def dissipation(self, t, data, dmin, dmax, sd, dim):
        for s in self.subsystems:
            if dim in s.dims:
                return s.dissipation(t, data, dmin, dmax, sd, dim)
        raise ValueError("dim not owned by any subsystem")


class Bird:
    """One flock member: the scalar-coefficient Hamiltonians of DynamicalSystems/bird.py:235-372.

    Only the state the hot path reads is kept: ``cur_state`` (3x1), ``v_e, v_p, w_e, w_p``, ``label``,
    ``neigh_rad`` and the neighbour list."""

    ndim = 3
    dims = (0, 1, 2)

    def __init__(self, grid, u_bound=1.0, w_bound=1.0, init_xyw=None, label=0, neigh_rad=3):
        self.grid = grid
        self.label, self.neigh_rad = label, neigh_rad
        self.v_e = self.v_p = u_bound                                      # bird.py:60-70 (v(.) ignores its argument)
        self.w_e = self.w_p = w_bound                                      # :73-78
        self.cur_state = np.asarray(init_xyw, dtype=np.float64).reshape(3, 1)
        self.neighbors = []

    def update_neighbor(self, other):                                      # :148-159
        if other in self.neighbors or other is self:
            return
        self.neighbors.append(other)

    @property
    def valence(self):
        return len(self.neighbors)

    def hamiltonian_abs(self, t, data, derivs, sd=None):                   # :266-273
        p1, p2, p3 = derivs[0], derivs[1], derivs[2]
        p1_coeff = -np.cos(self.cur_state[2, 0])
        p2_coeff = -np.sin(self.cur_state[2, 0])
        theta_r = -self.w_e
        return p1 * p1_coeff + p2 * p2_coeff + p3 * theta_r

    def hamiltonian(self, t, data, derivs, sd=None):                       # :305-316
        p1, p2, p3 = derivs[0], derivs[1], derivs[2]
        cs = self.cur_state
        p1_coeff = self.v_e - self.v_p * np.cos(cs[2, 0])
        p2_coeff = self.v_p * np.sin(cs[2, 0])
        w_up = max([n.w_e for n in self.neighbors])
        return ((p1 * p1_coeff - p2 * p2_coeff)
                + w_up * np.abs(p2 * cs[0, 0] - p1 * cs[1, 0] + p3)
                + w_up * np.abs(p3))

    def dissipation_abs(self, t, data, dmin, dmax, sd, dim):               # :318-344
        w_low = min([float(n.cur_state[2, 0]) for n in self.neighbors])
        cs = self.cur_state
        if dim == 0:
            return float(np.abs(self.v_p * np.cos(cs[2, 0])))
        if dim == 1:
            return float(np.abs(self.v_e * np.sin(cs[2, 0])))
        return w_low

    def dissipation(self, t, data, dmin, dmax, sd, dim):                   # :346-372
        w_up = max([float(n.cur_state[2, 0]) for n in self.neighbors])
        cs = self.cur_state
        if dim == 0:
            return float(np.abs(self.v_e - self.v_p * np.cos(cs[2, 0])) + np.abs(w_up * cs[1, 0]))
        if dim == 1:
            return float(np.abs(self.v_p * np.sin(cs[2, 0])) + np.abs(w_up * cs[0, 0]))
        return float(self.w_p + w_up)


class Flock:
    """DynamicalSystems/flock.py:97-258.  ``hamiltonian`` re-runs ``_housekeeping`` (neighbour + heading
    consensus update, :147-188) on every call, takes the element-wise minimum (shapeUnion,
    InitialConditions/shape_ops.py:39) over {H_abs(j != 0), H(attacked 0)} (:212-233); ``dissipation`` is the
    maximum over vehicles of the per-vehicle scalar alphas (:248-258).

    The shipped ``np.maximum.reduce(alphas, dtype=object)`` is fed a ragged list and raises on numpy >= 1.24;
    the scalar maximum below is what it evaluated to on the numpy versions it ran on."""

    ndim = 3
    dims = (0, 1, 2)

    def __init__(self, grid, vehicles):
        self.grid, self.vehicles = grid, list(vehicles)
        self.N = len(self.vehicles)
        self.attacked_idx = 0
        self._housekeeping()                                               # :145

    def _housekeeping(self):                                               # :147-163
        for i in range(self.N):
            for j in range(i + 1, self.N):
                self._compare(self.vehicles[i], self.vehicles[j])
            for j in range(i - 1, -1, -1):
                self._compare(self.vehicles[i], self.vehicles[j])
        for agent in self.vehicles:                                        # :162-163, :170-188
            headings = [n.w_e for n in agent.neighbors]
            agent.w_e = (1 / (1 + agent.valence)) * (agent.w_e + np.sum(headings))

    @staticmethod
    def _compare(a, b):                                                    # :165-168
        if np.abs(a.label - b.label) < a.neigh_rad:
            a.update_neighbor(b)

    def hamiltonian(self, t, data, derivs, sd=None):
        self._housekeeping()                                               # :213
        self.attacked_idx = 0                                              # :216
        others = [v for v in self.vehicles if v is not self.vehicles[0]]
        hams = [v.hamiltonian_abs(t, data, derivs, sd) for v in others]    # :222-225
        hams.append(self.vehicles[0].hamiltonian(t, data, derivs, sd))     # :229
        return np.minimum.reduce(hams)                                     # :232-233 (shape_ops.py:39)

    def dissipation(self, t, data, dmin, dmax, sd, dim):
        others = [v for v in self.vehicles if v is not self.vehicles[self.attacked_idx]]
        alphas = [v.dissipation_abs(t, data, dmin, dmax, sd, dim) for v in others]   # :248-252
        alphas.append(self.vehicles[self.attacked_idx].dissipation(t, data, dmin, dmax, sd, dim))   # :254-255
        return max(float(a) for a in alphas)                               # :257 (see class docstring)
