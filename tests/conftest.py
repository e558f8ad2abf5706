import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    import isingmodel_jl_b200
    return isingmodel_jl_b200


@pytest.fixture(scope="session")
def synth(pkg):
    from isingmodel_jl_b200 import synth as s
    return s


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def ctx(pkg):
    """Device context of the product library; fails loudly when the CUDA extension cannot run."""
    from isingmodel_jl_b200 import _lib
    return _lib.context(0)


def load_golden(name):
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def golden_matrix(g, synth, key):
    """Big matrices are not stored in the fixtures: regenerate from the seed and check the stored sha256."""
    import hashlib
    import numpy as np
    if key in g:
        return g[key]
    raise KeyError(key)


def sha(a):
    import hashlib
    import numpy as np
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
