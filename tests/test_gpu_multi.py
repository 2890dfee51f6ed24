"""ps_create_multi: ONE process, one handle, several GPUs -- the way a single cook thread
(exec/HDK_PolyStokes.C:222) would drive them.  The handle takes the caller's full-grid host arrays; the result must equal the
single-GPU handle's: `valid` bit for bit, counts and iteration count equal, velocity within 10 * tol; a second step (from another
calling thread) reproduces the first bit for bit; new per-step parameters (ps_set_params) give the fresh handle's result.  Needs >= 2 GPUs.
Every case runs in a child process (tests/multi_worker.py) under a time-out."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ["blob64_tile16_pad2", "blob48_tile8_pad1", "box64_uniform", "s3_128"]
NZ = {"blob64_tile16_pad2": 64, "blob48_tile8_pad1": 48, "box64_uniform": 64, "s3_128": 128}


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("ndev", [2, 4, 8])
@pytest.mark.parametrize("case", CASES)
def test_multi_handle_matches_single_gpu(built, case, ndev):
    if _ngpu() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    if NZ[case] < 16 * ndev:
        pytest.skip("fewer 16-layer slabs than devices")
    env = dict(os.environ, PS_MULTI_WATCHDOG_S="45")
    try:
        r = subprocess.run([sys.executable, os.path.join(HERE, "multi_worker.py"), case, str(ndev)], capture_output=True, text=True, timeout=240, env=env)
    except subprocess.TimeoutExpired as e:
        pytest.fail(f"multi-GPU handle hung on {case} x {ndev}:\n{(e.stderr or b'').decode() if isinstance(e.stderr, bytes) else e.stderr}")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
