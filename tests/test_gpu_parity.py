"""GPU parity: the CUDA library, called through the C ABI, against the CPU oracle on the same seeded scenes."""
import numpy as np
import pytest

import parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", list(parity.SCENE_CASES))
def test_gpu_step_matches_oracle(built, case):
    parity.run_case(case)


def test_gpu_s1_uniform_64(built):
    """BASELINE.json configs[0]: 64^3 liquid block in a solid box, uniform solve, tol 1e-6."""
    from polystokes_b200 import PolyStokesSolver, scenes
    from oracle.oracle import Oracle
    sc = scenes.scene_s1()
    o = Oracle(sc).setup()
    s = PolyStokesSolver.from_scene(sc)
    s.setup_scene(sc)
    parity.check_classification(o, s)
    parity.check_matrices(o, s)
    parity.check_operator(o, s)
    parity.check_solve(sc, o, s, {})


def test_gpu_device_resident_io(built):
    """Inputs / outputs living in HBM (torch tensors) give the same result as host arrays."""
    import torch
    from polystokes_b200 import PolyStokesSolver, scenes
    sc = scenes.blob_scene(32, seed=4)
    s = PolyStokesSolver.from_scene(sc)
    rc, vel, valid = s.step_scene(sc)
    d = lambda a: torch.from_numpy(a).cuda()
    vel_d = [d(v.copy()) for v in sc.vel]
    valid_d = [torch.zeros_like(v) for v in vel_d]
    rc2 = s.step(d(sc.surface), d(sc.collision), d(sc.viscosity), [d(v) for v in sc.vel], [d(v) for v in sc.colvel], vel_d, valid_d)
    torch.cuda.synchronize()
    assert rc == rc2
    for a in range(3):
        assert np.array_equal(vel[a], vel_d[a].cpu().numpy())
        assert np.array_equal(valid[a], valid_d[a].cpu().numpy())


def test_gpu_repeated_steps_are_deterministic(built):
    from polystokes_b200 import PolyStokesSolver, scenes
    sc = scenes.blob_scene(40)
    s = PolyStokesSolver.from_scene(sc)
    rc1, v1, _ = s.step_scene(sc)
    it1 = s.count("iterations")
    rc2, v2, _ = s.step_scene(sc)
    assert rc1 == rc2 and it1 == s.count("iterations")
    for a in range(3):
        assert np.array_equal(v1[a], v2[a])


def test_gpu_bicgstab_fallback(built):
    """CG out of iterations -> BiCGSTAB (S.cpp:784-799), against the oracle's restatement of pcg.h:134-200."""
    parity.check_bicgstab_fallback()

