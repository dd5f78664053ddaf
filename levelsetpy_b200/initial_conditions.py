"""Signed-distance initial conditions (host numpy, one-time): the subset of the reference's InitialConditions
package that the hot-path drivers, tests and bench use to build ``data0``.  Coordinates come from ``grid.xs``
(dense or sparse meshgrid -- both broadcast)."""
import numpy as np

__all__ = ["shapeCylinder", "shapeSphere", "shapeRectangleByCorners", "shapeUnion", "shapeIntersection",
           "shapeComplement", "shapeDifference"]


def _center(grid, center):
    if center is None or not np.any(center):
        return np.zeros((grid.dim, 1))
    center = np.asarray(center, dtype=np.float64).reshape(-1, 1)
    if center.size == 1:
        center = center.item() * np.ones((grid.dim, 1))
    return center


def shapeCylinder(grid, axis_align=None, center=None, radius=1):
    """sqrt(sum_{i not in axis_align} (x_i - c_i)^2) - radius -- InitialConditions/cylinder.py:8-60."""
    if axis_align is None:
        ignore = []
    elif np.isscalar(axis_align):
        ignore = [int(axis_align)]
    else:
        ignore = [int(a) for a in np.asarray(axis_align).reshape(-1)]
    center = _center(grid, center)
    data = np.zeros(grid.shape)
    for i in range(grid.dim):
        if i not in ignore:
            data = data + (grid.xs[i] - center[i]) ** 2
    return np.sqrt(data) - radius


def shapeSphere(grid, center=None, radius=1):
    """InitialConditions/sphere.py:9."""
    return shapeCylinder(grid, [], center, radius)


def shapeRectangleByCorners(grid, lower=None, upper=None):
    """max_i max(x_i - upper_i, lower_i - x_i) -- InitialConditions/rect_corners.py:67-70."""
    lower = np.zeros((grid.dim, 1)) if lower is None else np.asarray(lower, np.float64).reshape(-1, 1) * np.ones((grid.dim, 1))
    upper = np.ones((grid.dim, 1)) if upper is None else np.asarray(upper, np.float64).reshape(-1, 1) * np.ones((grid.dim, 1))
    data = np.maximum(grid.xs[0] - upper[0], lower[0] - grid.xs[0])
    for i in range(1, grid.dim):
        data = np.maximum(data, grid.xs[i] - upper[i])
        data = np.maximum(data, lower[i] - grid.xs[i])
    return np.broadcast_to(data, grid.shape).copy()


def shapeUnion(shapes, *more):
    """Pointwise minimum -- InitialConditions/shape_ops.py:12-47 (without its 2-shape IndexError at :37)."""
    if more:
        shapes = [shapes] + list(more)
    return np.minimum.reduce(list(shapes))


def shapeIntersection(shape1, shape2):
    return np.maximum(shape1, shape2)


def shapeComplement(shape):
    return -shape


def shapeDifference(shape1, shape2):
    return np.maximum(shape1, -shape2)
