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


def test_gpu_bicgstab_fallback_is_deterministic(built):
    """BiCGSTAB amplifies every rounding difference, so any race or read of stale memory in the fallback path shows up as a
    different iterate: three fresh handles (allocations land on different recycled memory) and a repeated step must agree bit for bit."""
    from polystokes_b200 import PolyStokesSolver, scenes
    sc = scenes.blob_scene(32, seed=4, maxIterations=60)
    runs = []
    for k in range(3):
        s = PolyStokesSolver.from_scene(sc)
        for rep in range(2 if k == 0 else 1):
            rc, v, _ = s.step_scene(sc)
            assert rc == 1 and s.count("usedBiCGStab") == 1
            runs.append((s.count("iterations"), s.vector("solution"), v))
        junk = [PolyStokesSolver.from_scene(scenes.blob_scene(24 + 4 * k)).step_scene(scenes.blob_scene(24 + 4 * k))]    # churn the allocator
        s.close()
    it0, x0, v0 = runs[0]
    for it, x, v in runs[1:]:
        assert it == it0 and np.array_equal(x, x0), "BiCGSTAB fallback is not reproducible"
        for a in range(3):
            assert np.array_equal(v[a], v0[a])


def test_gpu_bicgstab_fallback(built):
    """CG out of iterations -> BiCGSTAB (S.cpp:784-799), against the oracle's restatement of pcg.h:134-200."""
    parity.check_bicgstab_fallback()



@pytest.mark.parametrize("warm", [1, 0])
def test_gpu_eigen_cg(built, warm):
    """solverType EIGEN: Jacobi-preconditioned Eigen CG from the warm-start guess (S.cpp:814-862, 521-531) against the
    oracle's run on the explicit A."""
    parity.check_eigen_cg(warm=warm)


def test_gpu_eigen_cg_s2_beam(built):
    """The same path on the 128^3-style cantilever (BASELINE.json configs[1]) at 64^3: reduced tiles 8 pad 1, C6 widening."""
    from polystokes_b200 import scenes
    parity.check_eigen_cg(warm=1, scene=scenes.scene_s2(64))


def test_gpu_host_io_pinned_and_pageable(built):
    """Host callers: the overlapped PCIe path (copy streams) gives the same fields from pageable and from pinned buffers,
    with and without the valid outputs, and equals the device-resident result."""
    import torch
    from polystokes_b200 import PolyStokesSolver, scenes
    sc = scenes.blob_scene(36, seed=2)
    s = PolyStokesSolver.from_scene(sc)
    rc, vel, valid = s.step_scene(sc)                      # pageable numpy
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    ins = [pin(sc.surface), pin(sc.collision), pin(sc.viscosity)]
    pv, pc = [pin(v) for v in sc.vel], [pin(v) for v in sc.colvel]
    out = [torch.empty_like(v).pin_memory() for v in pv]
    val = [torch.empty_like(v).pin_memory() for v in pv]
    npv = lambda ts: [t.numpy() for t in ts]
    for _ in range(2):
        rc2 = s.step(ins[0].numpy(), ins[1].numpy(), ins[2].numpy(), npv(pv), npv(pc), npv(out), npv(val))
        assert rc == rc2
        for a in range(3):
            assert np.array_equal(vel[a], out[a].numpy()) and np.array_equal(valid[a], val[a].numpy())
    out2 = [np.zeros_like(v) for v in sc.vel]
    rc3 = s.step(sc.surface, sc.collision, sc.viscosity, sc.vel, sc.colvel, out2, None)     # velocity only
    assert rc3 == rc
    for a in range(3):
        assert np.array_equal(vel[a], out2[a])


@pytest.mark.parametrize("case", list(parity.EXPLICIT_CASES))
def test_gpu_explicit_A(built, case, tmp_path):
    """A2: the explicit system matrix built on the device against the oracle's sparse triple products."""
    parity.check_explicit_A(case, tmpdir=tmp_path)
