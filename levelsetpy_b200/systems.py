"""DynamicalSystems call surface: objects whose bound methods ``hamiltonian`` / ``dissipation`` are handed to
``schemeData.hamFunc`` / ``schemeData.partialFunc``.

On the device these are compiled functors (csrc/hj_systems.cuh); the Python objects hold the parameters and the
(host-side, scalar) flock bookkeeping, and their bound methods are the tokens ``functors.resolve`` maps to a
functor id + parameter block.  The reference's own DubinsVehicleRel / DoubleIntegrator / Bird / Flock instances
are recognised too (duck-typed), so an existing script keeps constructing its systems the way it does today.
"""
import numpy as np

__all__ = ["DubinsVehicleRel", "DoubleIntegrator", "Bird", "Flock", "ProductSystem"]

def _device_op(owner, ham_name, part_name, mutate):
    """(engine with this system's functor + parameter block set, adapter) for a standalone hamFunc / partialFunc call."""
    from .engine import engine_for_grid
    from .functors import _adapter_for_owner
    ad = _adapter_for_owner(owner, ham_name, part_name)
    eng = engine_for_grid(owner.grid)
    block = ad.block(mutate) if ad.time_varying else ad.block()
    eng.set_system(ad.system_id, block, list(enumerate(ad.tables(owner.grid))))
    return eng, ad, block


class _DeviceFunctorSystem:
    """Inside odeCFL3 / termLaxFriedrichs these methods are only tokens that name the compiled functor.  Called on their
    own -- by the reference's termLaxFriedrichs / artificialDissipationGLF, or by a user -- they evaluate that functor
    on dense arrays on the device (C-ABI hj_ham / hj_alpha)."""

    def hamiltonian(self, t, data, value_derivs, finite_diff_bundle=None):
        eng, _, _ = _device_op(self, "hamiltonian", "dissipation", True)
        return eng.ham(t, list(value_derivs))

    def dissipation(self, t, data, derivMin, derivMax, schemeData, dim):
        eng, ad, block = _device_op(self, "hamiltonian", "dissipation", False)
        if ad.host_alpha:
            return ad.alphas(block)[dim]                   # scalar alphas (bird.py:339-344, flock.py:248-258)
        return eng.alpha(t, dim, data)


class DubinsVehicleRel(_DeviceFunctorSystem):
    """Two Dubins vehicles in relative coordinates -- DynamicalSystems/dubins_relative.py:12-61.

    H = p1 (v_e - v_p cos x3) - p2 v_p sin x3 - w |p1 x2 - p2 x1 - p3| + w |p3|   (:84-88)"""

    ndim = 3

    def __init__(self, grid, u_bound=5, w_bound=5, x=None):
        if not np.isscalar(u_bound) or not np.isscalar(w_bound):
            raise NotImplementedError("vector-valued u_bound / w_bound have no device functor")
        self.grid = grid
        self.v = lambda u: u * u_bound
        self.w = lambda w: w * w_bound
        self.v_e, self.v_p = self.v(1), self.v(1)          # :52-53 (scalar bound)
        self.w_e, self.w_p = self.w(1), self.w(1)          # :60-61
        self.cur_state = x if x is not None else getattr(grid, "xs", None)


class DoubleIntegrator(_DeviceFunctorSystem):
    """x'' = u, |u| <= u_bound -- DynamicalSystems/double_integrator.py:9-89.  H = -(p1 x2 - |p2| u)  (:71-74)"""

    ndim = 2

    def __init__(self, grid, u_bound=1):
        self.grid = grid
        self.control_law = u_bound


class ProductSystem(_DeviceFunctorSystem):
    """Decoupled product of sub-systems on consecutive dim blocks: H = sum_k H_k(x_k, p_k), alpha_d from the
    owner of d.  (the 4-D double-integrator pair and the 6-D relative-Dubins pair.)"""

    def __init__(self, grid, subsystems):
        self.grid = grid
        self.subsystems = list(subsystems)
        self.ndim = sum(s.ndim for s in self.subsystems)
        if self.ndim != grid.dim:
            raise ValueError("sub-system dims sum to %d but the grid is %d-D" % (self.ndim, grid.dim))


class Bird(_DeviceFunctorSystem):
    """One flock member -- DynamicalSystems/bird.py:14-98 (the state the hot path reads)."""

    ndim = 3

    def __init__(self, grid, u_bound=+1, w_bound=+np.pi / 18, init_xyw=None, rw_cov=None, axis_align=2, center=None,
                 neigh_rad=3, init_random=False, label=0, payoff_width=.3):
        assert label is not None, "label of an agent cannot be empty"
        assert isinstance(init_xyw, np.ndarray), "initial state must either be a numpy or cupy array."
        self.grid, self.label, self.neigh_rad = grid, label, neigh_rad
        self.payoff_width, self.center, self.axis_align = payoff_width, center, axis_align
        self.v_e = self.v_p = u_bound                      # bird.py:60-70: v(.) ignores its argument
        self.w_e = self.w_p = w_bound
        self.neighbors = []
        cs = np.asarray(init_xyw, dtype=np.float64)
        if cs.ndim == 1:
            cs = cs.reshape(-1, 1)
        r, c = cs.shape
        self.cur_state = cs.T.copy() if r < c else cs.copy()

    def update_neighbor(self, neigh):
        if isinstance(neigh, list):
            for n in neigh:
                self.update_neighbor(n)
            return
        if neigh in self.neighbors or neigh is self:
            return
        self.neighbors.append(neigh)

    def reset_neighbors(self):
        self.neighbors = []

    @property
    def valence(self):
        return len(self.neighbors)

    def hamiltonian_abs(self, t, data, value_derivs, finite_diff_bundle=None):
        eng, _, _ = _device_op(self, "hamiltonian_abs", "dissipation_abs", True)
        return eng.ham(t, list(value_derivs))

    def dissipation_abs(self, t, data, derivMin, derivMax, schemeData, dim):
        _, ad, block = _device_op(self, "hamiltonian_abs", "dissipation_abs", False)
        return ad.alphas(block)[dim]


class Flock(_DeviceFunctorSystem):
    """A flock of Birds on one grid -- DynamicalSystems/flock.py:97-188.  The neighbour / heading-consensus
    bookkeeping (``_housekeeping``) stays on the host: it is O(#birds^2) scalar work re-run on every hamFunc
    call (flock.py:213), after which the per-bird scalar coefficients are shipped as the kernel's parameter block."""

    ndim = 3

    def __init__(self, grids, vehicles, label=1, reach_rad=1.0, avoid_rad=1.0):
        self.grid = grids
        self.vehicles = list(vehicles)
        self.N = len(self.vehicles)
        self.label, self.reach_rad, self.avoid_rad = label, reach_rad, avoid_rad
        self.attacked_idx = 0
        self._housekeeping()                               # flock.py:145

    def _housekeeping(self):                               # flock.py:147-163
        for i in range(self.N):
            for j in range(i + 1, self.N):
                self._compare_neighbor(self.vehicles[i], self.vehicles[j])
            for j in range(i - 1, -1, -1):
                self._compare_neighbor(self.vehicles[i], self.vehicles[j])
        for idx, agent in enumerate(self.vehicles):
            self._update_headings(agent, idx)

    @staticmethod
    def _compare_neighbor(agent1, agent2):                 # flock.py:165-168
        if np.abs(agent1.label - agent2.label) < agent1.neigh_rad:
            agent1.update_neighbor(agent2)

    @staticmethod
    def _update_headings(agent, idx, t=None):              # flock.py:170-188
        neighbor_headings = [neigh.w_e for neigh in agent.neighbors]
        agent.w_e = (1 / (1 + agent.valence)) * (agent.w_e + np.sum(neighbor_headings))
