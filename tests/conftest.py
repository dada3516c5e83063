import json
import pathlib
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import __graft_entry__ as entry  # noqa: E402

CASES = ["cent-par", "coop-par", "ncoop-par", "cent-ser", "coop-ser", "ncoop-ser"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    return entry.load_package()


@pytest.fixture(scope="session")
def setups(pkg):
    raw = json.loads((ROOT / "tests" / "golden" / "setups.json").read_text())
    return {k: pkg.setupfile.setup_from_dict(v) for k, v in raw.items()}


@pytest.fixture(scope="session")
def golden():
    return np.load(ROOT / "tests" / "golden" / "golden_traj.npz")


@pytest.fixture(scope="session")
def gpu_lib(pkg):
    """The CUDA library; GPU tests fail (not skip) when it is missing."""
    return pkg.capi.lib()
