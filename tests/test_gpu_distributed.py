"""Slab-decomposed step on real GPUs: one process per GPU, NCCL halo exchange + all-reduce inside the CUDA library,
merged result against the oracle.  Needs >= 2 GPUs (gpurun --gpus 2); skipped on a single-GPU box."""
import os

import pytest

import parity
from test_distributed_emul import launch

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("case,world", [("blob48_tile8", 2), ("box48_uniform", 2), ("blob64_tile16", 2), ("blob64_tile16", 4)])
def test_gpu_slab_decomposed_step_matches_oracle(built, tmp_path, case, world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29700 + (os.getpid() + hash(case) + world) % 200
    ranks = launch(case, world, tmp_path, port, gpu=True)
    parity.check_distributed(case, ranks)
