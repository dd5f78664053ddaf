"""CPU: the C-ABI library loads and exports every symbol include/hjb200.h declares (no compute calls without a
GPU), the ctypes table covers the header, and the host-side mirror of the reference's call surface (odeCFLset,
schemeData validation, functor registration, grids) behaves like the reference -- including its error behaviour."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def header_functions():
    txt = open(os.path.join(ROOT, "include", "hjb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(hj_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from levelsetpy_b200 import _lib
    lib = ctypes.CDLL(_lib.SO_PATH)
    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
        assert n in _lib.SIGNATURES, "ctypes table lacks %s" % n
    assert set(_lib.SIGNATURES) == set(names)
    lib.hj_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.hj_version()


def test_no_cpu_fallback_without_gpu(lsp):
    """On a box without a CUDA device the product path must fail loudly, never compute on the CPU."""
    from levelsetpy_b200 import _lib
    if _lib.load().hj_device_count() > 0:
        pytest.skip("a GPU is present")
    g = lsp.createGrid(np.array([-1., -1.]), np.array([1., 1.]), np.array([21, 21]))
    s = lsp.DoubleIntegrator(g, 1)
    sd = lsp.Bundle(dict(grid=g, hamFunc=s.hamiltonian, partialFunc=s.dissipation,
                         dissFunc=lsp.artificialDissipationGLF, CoStateCalc=lsp.upwindFirstWENO5))
    with pytest.raises(Exception) as ei:
        lsp.termLaxFriedrichs(0.0, np.zeros((441, 1)), sd)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)
    # the single hooks are device operators too (hj_ham / hj_diss_glf): without a GPU they fail the same loud way
    z = [np.zeros((21, 21)), np.zeros((21, 21))]
    for call in (lambda: s.hamiltonian(0, z[0], z, None),
                 lambda: lsp.artificialDissipationGLF(0, z[0], z, z, sd)):
        with pytest.raises(Exception) as ei:
            call()
        assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "levelsetpy_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


def test_ode_cfl_set_defaults_and_errors(lsp):
    """ode_cfl_set.py:81-133."""
    o = lsp.odeCFLset(lsp.Bundle({"factorCFL": 0.8, "singleStep": "on"}))
    assert o.factorCFL == 0.8 and o.singleStep == "on" and o.stats == "off" and o.postTimeStep is None
    assert o.maxStep == np.finfo(np.float64).max and o.terminalEvent is None
    assert lsp.odeCFLset(lsp.Bundle({"stats": "on"})).factorCFL == 0.5
    with pytest.raises(ValueError):
        lsp.odeCFLset()
    with pytest.raises(ValueError):
        lsp.odeCFLset(lsp.Bundle({"factorCFL": -1.0}))
    with pytest.raises(ValueError):
        lsp.odeCFLset(lsp.Bundle({"postTimeStep": 3}))


def test_rk3_time_arithmetic_matches_reference_lines():
    from levelsetpy_b200.integration import rk3_times
    t, dt = 0.3, 0.0123
    t1 = t + dt
    t2 = t1 + dt
    th = 0.25 * (3 * t + t2)
    t32 = th + dt
    assert rk3_times(t, dt) == (t1, th, (1 / 3) * (t + 2 * t32))


def test_create_grid_matches_reference_conventions(lsp):
    """create_grid.py:13-69 / process_grid.py:185-293: column N/dx/min/max, inclusive linspace, bdry by pdDims."""
    g = lsp.createGrid(np.array([-6., -10., 0.]), np.array([20., 10., 2 * np.pi * (1 - 1 / 13)]), np.array([21, 17, 13]), pdDims=2)
    assert g.dim == 3 and tuple(g.shape) == (21, 17, 13)
    assert np.asarray(g.dx).shape == (3, 1) and np.asarray(g.N).shape == (3, 1)
    assert np.allclose(np.asarray(g.dx).reshape(-1), [26 / 20, 20 / 16, 2 * np.pi * (1 - 1 / 13) / 12])
    assert [f.__name__ for f in g.bdry] == ["addGhostExtrapolate", "addGhostExtrapolate", "addGhostPeriodic"]
    assert np.asarray(g.vs[0]).reshape(-1)[0] == -6 and np.asarray(g.vs[0]).reshape(-1)[-1] == 20
    assert g.xs[1].shape == (21, 17, 13) and g.xs[1][3, 5, 7] == np.asarray(g.vs[1]).reshape(-1)[5]
    lm = lsp.createGrid(np.array([-1., -1.]), np.array([1., 1.]), np.array([9, 7]), low_mem=True)
    assert lm.xs[0].shape == (9, 1) and lm.xs[1].shape == (1, 7)


def test_scheme_validation_and_functor_registry(lsp):
    from levelsetpy_b200 import _lib as L
    from levelsetpy_b200.functors import resolve
    g3 = lsp.createGrid(np.array([-1., -1., 0.]), np.array([1., 1., 6.]), np.array([9, 9, 9]), pdDims=2)
    g2 = lsp.createGrid(np.array([-1., -1.]), np.array([1., 1.]), np.array([9, 9]))
    d = lsp.DubinsVehicleRel(g3, 5, 1)
    ad = resolve(d.hamiltonian, d.dissipation, g3)
    assert ad.system_id == L.SYS_DUBINS_REL and list(ad.block()) == [5, 5, 1, 1, 1]
    tabs = ad.tables(g3)
    assert np.array_equal(tabs[0], np.cos(np.asarray(g3.vs[2]).reshape(-1)))
    di = lsp.DoubleIntegrator(g2, 0.7)
    assert resolve(di.hamiltonian, di.dissipation, g2).system_id == L.SYS_DOUBLE_INT
    with pytest.raises(ValueError):
        resolve(di.hamiltonian, di.dissipation, g3)                  # 2-D system on a 3-D grid
    with pytest.raises(ValueError):
        resolve(d.hamiltonian, lsp.DubinsVehicleRel(g3, 5, 1).dissipation, g3)   # two different owners
    with pytest.raises(NotImplementedError):
        resolve(lambda *a: 0, lambda *a: 0, g3)                      # arbitrary callables: no device functor
    g4 = lsp.createGrid(-np.ones(4), np.ones(4), np.array([7, 7, 7, 7]))
    p = lsp.ProductSystem(g4, [lsp.DoubleIntegrator(g4, 1.0), lsp.DoubleIntegrator(g4, 0.5)])
    ap = resolve(p.hamiltonian, p.dissipation, g4)
    assert ap.system_id == L.SYS_DOUBLE_INT_PAIR and list(ap.block()) == [1.0, 0.5]
    with pytest.raises(ValueError):
        lsp.ProductSystem(g3, [lsp.DoubleIntegrator(g3, 1.0), lsp.DoubleIntegrator(g3, 0.5)])


def test_flock_block_follows_reference_housekeeping(lsp):
    """The Flock parameter block (host scalars) against the oracle's restatement of flock.py:147-258 / bird.py."""
    from levelsetpy_b200.functors import resolve
    from oracle import systems as osys
    g = lsp.createGrid(np.array([-1., -1., -np.pi]), np.array([1., 1., np.pi * (1 - 2 / 15)]), np.array([15, 15, 15]), pdDims=2)
    wb = [0.8, 1.0, 1.3, 0.6]
    xyw = [[0.1 * j - 0.05, 0.2 * j - 0.3, 0.3 * j + 0.1] for j in range(4)]
    fl = lsp.Flock(g, [lsp.Bird(g, 1.0, wb[j], init_xyw=np.array([xyw[j]]).T.copy(), label=j) for j in range(4)])
    of = osys.Flock(g, [osys.Bird(g, 1.0, wb[j], init_xyw=np.array(xyw[j]), label=j) for j in range(4)])
    ad = resolve(fl.hamiltonian, fl.dissipation, g)
    rng = np.random.default_rng(0)
    p = [rng.standard_normal((3, 4, 5)) for _ in range(3)]
    for _ in range(3):                                  # three RHS evaluations: the headings drift every call
        b = ad.block()
        h_want = of.hamiltonian(0.0, None, p)
        K = int(b[0])
        hs = [p[0] * b[10 + 3 * j] + p[1] * b[11 + 3 * j] + p[2] * b[12 + 3 * j] for j in range(K)]
        hs.append((p[0] * b[3] - p[1] * b[4]) + b[2] * np.abs(p[1] * b[5] - p[0] * b[6] + p[2]) + b[2] * np.abs(p[2]))
        assert np.allclose(np.minimum.reduce(hs), h_want, rtol=0, atol=1e-15)
        assert ad.alphas(b) == [of.dissipation(0.0, None, None, None, None, d) for d in range(3)]


def test_hjipde_solve_argument_errors(lsp):
    g = lsp.createGrid(np.array([-1., -1.]), np.array([1., 1.]), np.array([9, 9]))
    s = lsp.DoubleIntegrator(g, 1)
    sd = lsp.Bundle(dict(grid=g, hamFunc=s.hamiltonian, partialFunc=s.dissipation))
    with pytest.raises(ValueError):
        lsp.HJIPDE_solve(np.zeros((9, 9)), [0, 0.1], sd, "noSuchMethod", lsp.Bundle(dict(quiet=True)))
    with pytest.raises(ValueError):
        lsp.HJIPDE_solve(np.zeros((9, 8)), [0, 0.1], sd, "minVOverTime", lsp.Bundle(dict(quiet=True)))
    with pytest.raises(NotImplementedError):
        lsp.HJIPDE_solve(np.zeros((9, 9)), [0, 0.1], sd, "minVOverTime", lsp.Bundle(dict(quiet=True, visualize=True)))
