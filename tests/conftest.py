import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, not skip: only skip when gpu tests were not asked for.
    pass


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.fixture(scope="session")
def lsp():
    import levelsetpy_b200
    return levelsetpy_b200


def make_grid(lsp_mod, gold):
    """Rebuild the grid of a golden case with this package's createGrid."""
    pd = [int(i) for i in np.nonzero(gold["periodic"])[0]]
    g = lsp_mod.createGrid(gold["grid_min"], gold["grid_max"], gold["grid_N"], pdDims=pd if pd else None)
    return g
