"""compute-sanitizer is closed on this GPU pool, so out-of-bounds writes are looked for with the library's own guard
bands: with MVS_GUARD=1 every workspace buffer is allocated at exactly the requested size followed by 4 KB of a known byte
(csrc/api.cu DevBuf::ensure), and tests/conftest.py checks after every test that no band of any open context was touched.
This test re-runs the single-GPU parity suites (ragged frames, empty and thin inputs, 8k-keypoint frames, cross-check, both
solvers, every launch shape of the tail kernels, PnP, bundle adjustment, ORB) in such a process."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUITES = ["test_gpu_parity.py", "test_gpu_reference_solver.py", "test_gpu_ba.py", "test_gpu_pnp.py", "test_gpu_orb.py"]


def test_no_kernel_writes_past_its_workspace_buffers():
    if os.environ.get("MVS_GUARD") == "1":
        pytest.skip("already inside the guarded run")
    env = dict(os.environ, MVS_GUARD="1")
    files = [os.path.join(ROOT, "tests", f) for f in SUITES if os.path.exists(os.path.join(ROOT, "tests", f))]
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider", "-k", "not two_gpu and not sharded"] + files,
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    tail = (r.stdout + r.stderr)[-2000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout, tail


def test_guard_mode_detects_an_overrun():
    """the check itself: in a guarded process, a deliberate write one byte past a buffer is reported"""
    code = (
        "import numpy as np, mvslam_b200 as mvs\n"
        "ctx = mvs.Context(0)\n"
        "q = np.random.default_rng(0).integers(0, 256, (64, 32), dtype=np.uint8)\n"
        "ctx.knn2_hamming(q, q)\n"
        "assert ctx.debug_guard_check() == 0\n"
        "assert ctx.debug_guard_poke() == 0\n"
        "print('bad', ctx.debug_guard_check())\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=dict(os.environ, MVS_GUARD="1"), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "bad 1" in r.stdout, r.stdout + r.stderr
