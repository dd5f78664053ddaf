"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Runs the *literal* reference (robotsorcerer/LevelSetPy, mounted read-only at
/root/reference in the build container) on CPU by faking the three packages it
needs but that are not installed here:

  * ``cupy``       -> numpy, plus an ndarray subclass that has ``.get()`` and
                      CuPy's wrap-around semantics for out-of-range integer
                      array gathers (upwind_first_weno5a.py:143-145 indexes
                      ``N`` into a length-``N`` axis, which only works on CuPy).
  * ``matplotlib`` / ``mpl_toolkits`` / ``skimage`` -> inert stubs
                      (ValueFuncs/hji_solver.py:9,22 and Visualization/* import
                      them at module import time).

The reference's modules do ``from LevelSetPy.X import *`` so a directory holding a
``LevelSetPy -> /root/reference`` symlink is put on ``sys.path`` (created under
a temp dir, nothing is written to /root/reference).

Used by ``tests/golden/make_golden.py`` (golden-vector generation) and by
``tests/test_reference_shim.py`` (skipped when /root/reference is absent, e.g.
on the GPU box).
"""
import importlib.machinery
import os
import sys
import tempfile
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("LEVELSETPY_REFERENCE", "/root/reference")


class _CpArray(np.ndarray):
    """numpy array that behaves like a cupy.ndarray where the reference needs it."""

    def get(self):
        return np.asarray(self)

    def __getitem__(self, key):
        if isinstance(key, tuple) and key and all(isinstance(k, np.ndarray) and k.dtype.kind in "iu" for k in key):
            # CuPy does not bounds-check integer-array gathers; they wrap.
            if len(key) == self.ndim:
                key = tuple(np.mod(k, n) if n > 0 else k for k, n in zip(key, self.shape))
        return super().__getitem__(key)


def _wrap(a):
    a = np.asarray(a)
    return a.view(_CpArray)


def _make_fake_cupy():
    cp = types.ModuleType("cupy")
    cp.__spec__ = importlib.machinery.ModuleSpec("cupy", None)
    cp.ndarray = _CpArray

    def _lift(fn):
        # every cupy function returns cupy arrays (0-d arrays for reductions), whatever it was fed
        def inner(*a, **k):
            r = fn(*a, **k)
            if isinstance(r, (np.ndarray, np.generic)):
                return _wrap(r)
            if isinstance(r, tuple):
                return tuple(_wrap(x) if isinstance(x, (np.ndarray, np.generic)) else x for x in r)
            return r
        inner.__name__ = getattr(fn, "__name__", "cupy_fn")
        return inner

    class _Device:
        def __init__(self, *a, **k):
            pass

        def synchronize(self):
            pass

        def use(self):
            pass

    cuda = types.ModuleType("cupy.cuda")
    cuda.Device = _Device
    cp.cuda = cuda

    def __getattr__(name):
        obj = getattr(np, name)
        if callable(obj) and not isinstance(obj, type):
            return _lift(obj)
        return obj

    cp.__getattr__ = __getattr__
    return cp, cuda


class _Anything:
    """Absorbs any attribute access / call (for matplotlib, skimage stubs)."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


def _stub_module(name):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []
    m.__all__ = []

    def __getattr__(attr):
        if attr.startswith("__") and attr.endswith("__"):
            raise AttributeError(attr)
        return _Anything()

    m.__getattr__ = __getattr__
    return m


_STUB_ROOTS = ("matplotlib", "mpl_toolkits", "skimage", "pyvista")


class _StubFinder:
    """Meta-path finder: any (sub)module of the absent plotting packages becomes an inert stub."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _stub_module(spec.name)

    def exec_module(self, module):
        pass


_installed = None


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "SpatialDerivative"))


def install():
    """Install the fakes and put the reference on sys.path. Idempotent. Returns the LevelSetPy module."""
    global _installed
    if _installed is not None:
        return _installed
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    cp, cuda = _make_fake_cupy()
    sys.modules.setdefault("cupy", cp)
    sys.modules.setdefault("cupy.cuda", cuda)
    sys.meta_path.append(_StubFinder())  # appended: real packages win if they are installed
    tmp = tempfile.mkdtemp(prefix="lsp_ref_")
    os.symlink(REFERENCE_ROOT, os.path.join(tmp, "LevelSetPy"))
    sys.path.insert(0, tmp)
    # SpatialDerivative/__init__.py:8 does an absolute ``from SpatialDerivative.Other import *``.
    sys.path.insert(0, REFERENCE_ROOT)
    import LevelSetPy  # noqa: F401
    _installed = LevelSetPy
    return LevelSetPy
