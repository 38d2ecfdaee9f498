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
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return load


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference scripts, stub-imported.  Skips where /root/reference is absent
    (the GPU box) — nothing marked gpu may use this."""
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference tree not present")
    return ref_import.load()


@pytest.fixture(scope="session")
def orc():
    from oracle import pcr_oracle
    pcr_oracle.build()
    return pcr_oracle


@pytest.fixture(scope="session")
def lib():
    from pointcloud_render_b200 import build as pcr_build
    pcr_build.build()
    from pointcloud_render_b200 import _native
    return _native.load_library()
