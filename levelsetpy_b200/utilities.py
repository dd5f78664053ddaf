"""Host-side MATLAB-style helpers of the call surface (Utilities/matlab_utils.py in the reference):
the attribute-bag ``Bundle`` that carries ``schemeData`` / ``grid`` / options, and a few predicates."""
import logging
import sys
import time

import numpy as np

logger = logging.getLogger("levelsetpy_b200")

realmin = sys.float_info.min
realmax = sys.float_info.max
eps = sys.float_info.epsilon

__all__ = ["Bundle", "cell", "iscell", "isbundle", "isfield", "expand", "strcmp", "isscalar", "numel", "size",
           "info", "warn", "error", "cputime", "realmin", "realmax", "eps", "to_column_mat", "deg2rad", "rad2deg"]


class Bundle(object):
    """Struct-like attribute bag (matlab_utils.py:41-57)."""

    def __init__(self, dicko=None, **kw):
        for k, v in dict(dicko or {}, **kw).items():
            object.__setattr__(self, k, v)

    def __len__(self):
        return len(self.__dict__)

    def keys(self):
        return list(self.__dict__.keys())

    def __repr__(self):
        return "Bundle(%s)" % ", ".join(sorted(self.__dict__))


def cell(n, dim=1):
    return [np.nan for _ in range(n)]


def iscell(x):
    return isinstance(x, list)


def isbundle(x):
    return isinstance(x, Bundle) or (hasattr(x, "__dict__") and type(x).__name__ == "Bundle")


def isfield(bund, field):
    return field in getattr(bund, "__dict__", {})


def expand(x, ax):
    return np.expand_dims(x, ax)


def strcmp(a, b):
    return a == b


def numel(x):
    return int(np.size(x))


def size(x, dim=None):
    s = np.shape(x)
    return s if dim is None else s[dim]


def isscalar(x):
    if isinstance(x, np.ndarray):
        return x.size == 1
    return not isinstance(x, (list, tuple))


def to_column_mat(x):
    return np.asarray(x).reshape(-1, 1)


def deg2rad(x):
    return x * (np.pi / 180)


def rad2deg(x):
    return (x * 180) / np.pi


def cputime():
    return time.time()


def info(msg):
    logger.info(msg)


def warn(msg):
    logger.warning(msg)


def error(msg):
    """The reference's error() logs and raises ValueError (matlab_utils.py:134-147)."""
    raise ValueError(msg)
