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


@pytest.mark.parametrize("transport", ["peer", "nccl"])
@pytest.mark.parametrize("case,world", [("blob48_tile8", 2), ("box48_uniform", 2), ("blob64_tile16", 2), ("blob64_tile16", 4), ("blob48_bicgstab6", 2),
                                        ("blob48_tile16_pad3_layers33", 3), ("s3_128", 4), ("s3_128", 8), ("blob64_tile16_then_shrunk", 2)])
def test_gpu_slab_decomposed_step_matches_oracle(built, tmp_path, case, world, transport):
    """transport "peer": halo stores + fused all-reduce over NVLink peer memory (ps_peer.hpp); "nccl": the fallback."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29700 + (os.getpid() + hash(case) + world + (7 if transport == "nccl" else 0)) % 200
    ranks = launch(case, world, tmp_path, port, gpu=True, extra_env={"PS_COMM": "nccl"} if transport == "nccl" else {"PS_COMM": ""})
    assert all(int(r["peer"]) == (1 if transport == "peer" else 0) for r in ranks), "requested transport not in use"
    parity.check_distributed(case, ranks)
