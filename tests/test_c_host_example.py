"""The C-ABI from a host without Python: examples/host_air3d.c is plain C99 against include/hjb200.h, linked with the
in-tree library.  CPU: it compiles with -pedantic, links, and without a CUDA device fails loudly through the library's own
error text (no CPU fallback).  GPU: its result equals the Python mirror's on the same problem (t to rounding of the C
libm's cos / sin against numpy's, field statistics within 1e-9 of range)."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "examples", "host_air3d.c")
SO = os.path.join(ROOT, "levelsetpy_b200", "_hjb200.so")


def _build(tmp_path):
    if shutil.which("gcc") is None:
        pytest.skip("no gcc on this box")
    if not os.path.exists(SO):
        import __graft_entry__
        __graft_entry__.build()
    exe = str(tmp_path / "host_air3d")
    cmd = ["gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-pedantic", "-Werror", "-ffp-contract=off",
           "-I", os.path.join(ROOT, "include"), SRC, "-o", exe, SO, "-lm",
           "-Wl,-rpath," + os.path.dirname(SO)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_host_compiles_links_and_fails_loudly_without_a_gpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the run itself is test_c_host_matches_python_mirror")
    r = subprocess.run([exe, "21", "2"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3
    assert "no CUDA device" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_host_matches_python_mirror(lsp, tmp_path):
    from levelsetpy_b200 import _lib as L
    from levelsetpy_b200.integration import rk3_step_resident
    from levelsetpy_b200.term import prepare_scheme
    exe = _build(tmp_path)
    n, steps = 41, 4
    r = subprocess.run([exe, str(n), str(steps)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    f = r.stdout.split()
    assert int(f[0]) == n and int(f[1]) == steps
    t_c, sum_c, min_c, max_c, launches = float(f[2]), float(f[3]), float(f[4]), float(f[5]), int(f[6])
    assert launches >= 3 * steps                              # the C host really launched the stage kernels
    g = lsp.createGrid(np.array([-6.0, -10.0, 0.0]), np.array([20.0, 10.0, 2 * np.pi * (1 - 1 / n)]), np.array([n, n, n]), pdDims=2)
    d0 = lsp.shapeCylinder(g, 2, np.zeros((3, 1)), 5)
    s = lsp.DubinsVehicleRel(g, 5, 1)
    sd = lsp.Bundle(dict(grid=g, hamFunc=s.hamiltonian, partialFunc=s.dissipation, dissFunc=lsp.artificialDissipationGLF,
                         CoStateCalc=lsp.upwindFirstWENO5a))
    eng, ad = prepare_scheme(sd)
    eng.upload(d0.reshape(-1, 1))
    t = 0.0
    for _ in range(steps):
        t, _ = rk3_step_resident(eng, ad, g, t, 1.0, 0.8, sys.float_info.max, L.COMP_MIN_OVER_TIME)
    y = eng.download(shape=g.shape)
    rng = float(y.max() - y.min())
    assert abs(t_c - t) <= 1e-12 * t
    assert abs(min_c - float(y.min())) <= 1e-9 * rng and abs(max_c - float(y.max())) <= 1e-9 * rng
    assert abs(sum_c - float(y.sum())) <= 1e-9 * rng * y.size
