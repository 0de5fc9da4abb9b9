import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    import youth_pkg

    mod = youth_pkg.load()
    paths = mod.lib_paths()
    if not (os.path.exists(paths["cuda"]) and os.path.exists(paths["host"])):
        mod.build_all()
    return mod


@pytest.fixture(scope="session")
def oracle():
    import oracle_py

    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def small_seq(pkg):
    """6 synthetic 640x480 frames (sequence 0, noise off) + ground truth."""
    return pkg.synth_sequence(6), pkg.synth_gt(6)
