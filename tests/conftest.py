import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_meshes():
    z = np.load(os.path.join(ROOT, "tests", "golden", "meshes.npz"))

    def get(name):
        return z[name + "/points"], z[name + "/elements"], z[name + "/boundary"]

    return get


@pytest.fixture(scope="session")
def fdb():
    import __graft_entry__ as g
    return g.load_package()


def entry_tolerance(outer, inner, v_ref, rel=1e-12):
    """|gpu - ref| <= rel * max(|ref|, |diagonal of the column|): relative 1e-12 with an absolute floor for the
    entries that are mathematically zero (hard part 2)."""
    n = outer.size - 1
    col = np.repeat(np.arange(n), np.diff(outer))
    diag = np.zeros(n)
    on_diag = inner == col
    diag[col[on_diag]] = np.abs(v_ref[on_diag])
    return rel * np.maximum(np.abs(v_ref), diag[col])
