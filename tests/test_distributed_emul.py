"""N > 1 path on CPU: world_size-2 / -3 gloo jobs drive the z-slab decomposed step of the host-emulation twin
(partition, ownership ranges, halo lists, distributed CG with two all-reduces per iteration, shared-plane merge)
and the merged result is compared with the single-process oracle.  The product library runs the same host code
with NCCL in place of the gloo callbacks (tests/test_gpu_distributed.py, -m gpu)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import parity
from oracle.oracle import Oracle

HERE = os.path.dirname(os.path.abspath(__file__))


def launch(case, world, outdir, port, gpu=False, extra_env=None):
    env = dict(os.environ, **(extra_env or {}))
    env = dict(env, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), OMP_NUM_THREADS="2")
    procs = []
    for r in range(world):
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "dist_worker.py"), case, str(outdir)] + (["gpu"] if gpu else []),
                                      env=dict(env, RANK=str(r), LOCAL_RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    fail = []
    for r, p in enumerate(procs):
        try:
            out, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise AssertionError(f"rank {r} timed out")
        if p.returncode != 0:
            fail.append(f"rank {r} exit {p.returncode}:\n{out.decode()[-3000:]}")
    assert not fail, "\n".join(fail)
    return [np.load(os.path.join(outdir, f"rank{r}.npz")) for r in range(world)]


@pytest.mark.parametrize("case,world", [("blob48_tile8", 2), ("box48_uniform", 2), ("blob_36x40x64_tile16", 3), ("blob48_bicgstab6", 2), ("blob48_tile16_pad3_layers33", 3),
                                        ("blob64_tile16", 4), ("blob_40x36x96_tile32_pad3", 2), ("blob64_tile16_then_shrunk", 2)])
def test_slab_decomposed_step_matches_oracle(built, tmp_path, case, world):
    port = 29600 + (os.getpid() + hash(case)) % 300
    ranks = launch(case, world, tmp_path, port)
    parity.check_distributed(case, ranks)
