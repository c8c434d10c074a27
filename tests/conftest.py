import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def tsukuba():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "tsukuba_orb2000.npz"))


@pytest.fixture(scope="session")
def tsukuba_golden():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "tsukuba_golden.npz"))


@pytest.fixture(scope="session")
def synthetic_golden():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "synthetic_golden.npz"))


@pytest.fixture(params=["reference", "fast"])
def solver(request):
    """Runs a geometry test once per solver (include/mvslam_b200.h MVS_SOLVER_*): sets the default of both the
    library binding and the oracle binding, so GPU and oracle are always compared in the same mode."""
    from mvslam_b200 import capi
    from oracle import cbind
    old = (capi.DEFAULT_SOLVER, cbind.DEFAULT_SOLVER)
    capi.set_default_solver(request.param); cbind.set_default_solver(request.param)
    yield request.param
    capi.set_default_solver(old[0]); cbind.set_default_solver(old[1])


@pytest.fixture(autouse=True)
def _guard_bands():
    """MVS_GUARD=1 runs (tests/test_gpu_guards.py starts one): after every test, no kernel may have written past the end of
    a workspace buffer of any open context (include/mvslam_b200.h mvs_debug_guard_check)."""
    yield
    if os.environ.get("MVS_GUARD") == "1":
        from mvslam_b200 import capi
        for c in list(capi._LIVE):
            if getattr(c, "_h", None):
                assert c.debug_guard_check() == 0, "a kernel wrote past the end of a workspace buffer"
