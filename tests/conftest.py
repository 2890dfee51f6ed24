import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # oracle/_ref (the reference's own solve stage, compiled from its sources) exists only where /root/reference does; build it
    # before collection so tests/test_ref_solve.py is not skipped here.  On the GPU box the prebuilt file travels with the snapshot.
    if os.path.isdir("/root/reference/lib/include"):
        import subprocess
        try:
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "ref"])
        except Exception as e:      # the tests that need it will say so
            print(f"conftest: could not build oracle/_ref: {e}", file=sys.stderr)


@pytest.fixture(scope="session")
def built():
    """Make sure the oracle and the host-emulation twin exist (cheap no-op when up to date)."""
    import subprocess
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "polystokes_b200", "csrc"), "-s", "emul"])
    return True
