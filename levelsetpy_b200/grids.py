"""Grid structure of the call surface (host only).

``createGrid`` / ``processGrid`` build the same ``Bundle`` the reference's Grids/create_grid.py:13-69 and
Grids/process_grid.py:12-302 build -- fields ``dim, min, max, N, dx, vs, xs, bdry, bdryData, shape, axis_align`` --
with bit-identical ``dx`` (process_grid.py:185) and ``vs`` (``np.linspace(min, max, N)``, :204).  ``xs`` is a dense
``np.meshgrid(..., indexing='ij')`` for small grids (:234) and a sparse (broadcastable) one above ``DENSE_XS_LIMIT``
nodes: the device path never reads ``xs`` (coordinates are generated from ``vs`` on the fly), and a dense 41**6
meshgrid would be 6 x 38 GB.
"""
import copy

import numpy as np

from .boundary import addGhostExtrapolate, addGhostPeriodic
from .utilities import Bundle, error, isfield, warn

__all__ = ["createGrid", "processGrid", "flockGrid"]

DENSE_XS_LIMIT = 1 << 24


def _col(x, dtype=None):
    return np.asarray(x, dtype=dtype).reshape(-1, 1)


def createGrid(grid_min, grid_max, N, pdDims=None, process=True, low_mem=False):
    """g = createGrid(grid_min, grid_max, N, pdDims) -- Grids/create_grid.py:13-69.

    pdDims: periodic dimension(s) (int or list); those get ``addGhostPeriodic``, the rest ``addGhostExtrapolate``.
    As in the reference the caller shrinks ``grid_max`` of a periodic dim so that the last node is not a duplicate."""
    grid_min = _col(grid_min, np.float64)
    grid_max = _col(grid_max, np.float64)
    if np.isscalar(N) or np.size(N) == 1:
        N = int(np.asarray(N).reshape(-1)[0]) * np.ones(grid_min.shape, dtype=np.int64)
    N = _col(N, np.int64)
    if not (grid_min.size == grid_max.size == N.size):
        raise AssertionError("grid min, grid_max, and N must have the same number of elements!")
    if pdDims is None:
        pd = []
    elif np.isscalar(pdDims):
        pd = [int(pdDims)]
    else:
        pd = [int(p) for p in np.asarray(pdDims).reshape(-1)]
    dim = int(grid_min.size)
    g = Bundle(dict(dim=dim, min=grid_min, max=grid_max, N=N, bdry=[None] * dim))
    g.axis_align = pdDims
    for i in range(dim):
        g.bdry[i] = addGhostPeriodic if i in pd else addGhostExtrapolate
    if process:
        g = processGrid(g, sparse_flag=low_mem)
    return g


def processGrid(gridIn, data=None, sparse_flag=False):
    """Fill in / check the derived grid fields -- Grids/process_grid.py:100-302 (Bundle input only)."""
    if not (hasattr(gridIn, "__dict__") and isfield(gridIn, "dim")):
        error("Grid structure must contain dimension")
    g = copy.copy(gridIn)
    if g.dim > 5:
        warn("Grid dimension > 5, may be dangerously large")      # process_grid.py:133-134 (warning only)
    if g.dim <= 0:
        error("Grid dimension must be positive")
    g.min = _col(g.min, np.float64) if isfield(g, "min") else np.zeros((g.dim, 1))
    g.max = _col(g.max, np.float64) if isfield(g, "max") else np.ones((g.dim, 1))
    if g.min.size == 1 and g.dim > 1:
        g.min = g.min.item() * np.ones((g.dim, 1))
    if g.max.size == 1 and g.dim > 1:
        g.max = g.max.item() * np.ones((g.dim, 1))
    if np.any(g.max <= g.min):
        error("max bound must be strictly greater than min bound in all dimensions")
    if isfield(g, "N"):
        g.N = _col(g.N, np.int64)
        if g.N.size == 1 and g.dim > 1:
            g.N = g.N.item() * np.ones((g.dim, 1), dtype=np.int64)
        if np.any(g.N <= 0):
            error("number of grid cells must be strictly positive")
    if isfield(g, "dx"):
        g.dx = _col(g.dx, np.float64)
        if np.any(g.dx <= 0):
            error("grid cell size dx must be strictly positive")
    elif isfield(g, "N"):
        g.dx = np.divide(g.max - g.min, g.N - 1)                 # process_grid.py:185
    else:
        g.N = 101 * np.ones((g.dim, 1), dtype=np.int64)
        g.dx = np.divide(g.max - g.min, g.N - 1)
    if not isfield(g, "vs"):
        g.vs = [np.expand_dims(np.linspace(g.min[i, 0].item(), g.max[i, 0].item(), num=int(g.N[i, 0])), 1)
                for i in range(g.dim)]                           # process_grid.py:204
    for i in range(g.dim):
        if int(g.N[i, 0]) != len(g.vs[i]):
            error("Inconsistent grid size in dimension %d" % i)
    nodes = int(np.prod(g.N.astype(np.float64)))
    if not isfield(g, "xs"):
        g.xs = np.meshgrid(*g.vs, indexing="ij", sparse=bool(sparse_flag or nodes > DENSE_XS_LIMIT))
    if not isfield(g, "bdry") or g.bdry is None:
        g.bdry = [addGhostPeriodic for _ in range(g.dim)]        # process_grid.py:103 default
    elif not isinstance(g.bdry, (list, tuple, np.ndarray)):
        g.bdry = [g.bdry for _ in range(g.dim)]
    if len(g.bdry) != g.dim:
        error("bdry field is not column cell vector of length dim: %d" % g.dim)
    if not isfield(g, "bdryData") or g.bdryData is None:
        g.bdryData = [None for _ in range(g.dim)]
    elif not isinstance(g.bdryData, (list, tuple)):
        g.bdryData = [g.bdryData for _ in range(g.dim)]
    if g.dim in (2, 3):
        g.axis = np.zeros((1, 2 * g.dim))
        for i in range(g.dim):
            g.axis[0, 2 * i:2 * i + 2] = [g.min[i, 0].item(), g.max[i, 0].item()]
    else:
        g.axis = []
    nshape = tuple(int(x) for x in g.N.reshape(-1))
    g.shape = nshape + (1,) if g.dim == 1 else nshape
    if data is not None and np.shape(data) != g.shape:
        error("data parameter does not agree in array size with grid")
    return g


def flockGrid(grid_mins=None, grid_maxs=None, dx=.2, num_agents=10, N=101):
    """One grid per agent, each shifted by ``dx`` from the previous -- Grids/flock_grid.py:6-42."""
    grid_mins = [list(m) for m in (grid_mins or [[-1, -1, -np.pi]])]
    grid_maxs = [list(m) for m in (grid_maxs or [[1, 1, np.pi]])]
    grids = [createGrid(np.asarray(grid_mins[0]), np.asarray(grid_maxs[0]), N=N, pdDims=2)]
    for agent in range(1, num_agents):
        grid_mins.append([x - dx for x in grid_mins[agent - 1]])
        grid_maxs.append([x - dx for x in grid_maxs[agent - 1]])
        grids.append(createGrid(np.asarray(grid_mins[agent - 1]), np.asarray(grid_maxs[agent - 1]), N=N, pdDims=2))
    return grids
