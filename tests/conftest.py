import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """The tests load the in-tree libising_b200.so; build it (incremental make, nvcc cross-compiles without a GPU)
    when it is missing, e.g. on a fresh checkout where build() has not run yet."""
    lib = os.path.join(ROOT, "isingmodel.jl_b200", "libising_b200.so")
    if not os.path.exists(lib):
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, "isingmodel.jl_b200", "csrc"), "-j", "8"],
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=False)


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
