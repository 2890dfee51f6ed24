"""ps_set_params (include/polystokes_b200.h): a DOP network cooks the node with a different dt on every substep (exec/HDK_PolyStokes.C:319);
the handle must take the new per-step parameters without being rebuilt and give exactly what a fresh handle gives.  Runs on the host-emulation
twin (the product sources compiled with -DPS_EMULATE); the CUDA library is exercised by tests/multi_worker.py and tests/test_zz_adaptor.py."""
import ctypes as C

import numpy as np
import pytest

import parity
from polystokes_b200 import PolyStokesSolver, scenes, _capi


def test_new_dt_and_tiling_on_the_same_handle_equal_a_fresh_handle(built):
    sc = scenes.blob_scene(32, seed=13, tile=16, pad=2)
    a = PolyStokesSolver.from_scene(sc, lib_path=parity.EMUL_LIB)
    rc, v, _ = a.step_scene(sc)
    assert rc == 1
    P = a._P
    P.dt = float(sc.dt) * 0.5
    P.tileSize = 8
    P.tilePadding = 1
    assert a.lib.ps_set_params(a.h, P) == 1, a.last_error()
    rc2, v2, val2 = a.step_scene(sc)
    b = PolyStokesSolver(sc.nx, sc.ny, sc.nz, sc.dx, float(sc.dt) * 0.5, sc.density, lib_path=parity.EMUL_LIB, **dict(sc.params, tileSize=8, tilePadding=1))
    rc3, v3, val3 = b.step_scene(sc)
    assert rc2 == rc3 == 1 and a.count("iterations") == b.count("iterations") and a.count("regionCount") == b.count("regionCount")
    for ax in range(3):
        assert np.array_equal(v2[ax], v3[ax]) and np.array_equal(val2[ax], val3[ax])
        assert not np.array_equal(v[ax], v2[ax]), "the new dt must change the result"
    a.close(); b.close()


def test_the_grid_is_fixed_at_create(built):
    sc = scenes.blob_scene(24, seed=3)
    a = PolyStokesSolver.from_scene(sc, lib_path=parity.EMUL_LIB)
    P = _capi.ps_params()
    C.memmove(C.byref(P), C.byref(a._P), C.sizeof(P))
    P.nx = sc.nx + 1
    assert a.lib.ps_set_params(a.h, P) == -1 and "fixed at ps_create" in a.last_error()
    P.nx = sc.nx
    P.dt = 0.0
    assert a.lib.ps_set_params(a.h, P) == -1 and "dt" in a.last_error()
    assert a.lib.ps_set_params(a.h, None) == -2 and a.lib.ps_set_params(None, P) == -2
    a.close()


def test_multi_device_handle_needs_cuda(built):
    """the emulation twin is one rank per process: ps_create_multi must fail loudly there (no silent single-device fallback)"""
    sc = scenes.blob_scene(24, seed=3)
    with pytest.raises(Exception) as e:
        PolyStokesSolver.from_scene(sc, lib_path=parity.EMUL_LIB, devices=[0, 1])
    assert "ps_create_multi" in str(e.value)


def _reuse_for_another_scene(lib_path):
    """a handle that has stepped one scene is handed a different one (other liquid shape, other system size, other region count): every grow-only
    buffer and every cached table must be rebuilt -- the result equals a fresh handle's bit for bit, in both directions (larger -> smaller too)"""
    big = scenes.blob_scene(40, seed=13, tile=8, pad=1)
    small = scenes.blob_scene(40, seed=4, tile=8, pad=1)
    small.surface = small.surface + 2.5 * small.dx          # shrink the blob: fewer unknowns, fewer regions than `big`
    kw = dict(lib_path=lib_path) if lib_path else {}
    for first, second in ((big, small), (small, big)):
        a = PolyStokesSolver.from_scene(first, **kw)
        assert a.step_scene(first)[0] == 1
        n1 = a.count("nSystemSize")
        rc, v, val = a.step_scene(second)
        b = PolyStokesSolver.from_scene(second, **kw)
        rcb, vb, valb = b.step_scene(second)
        assert rc == rcb == 1 and a.count("nSystemSize") == b.count("nSystemSize") != n1
        assert a.count("iterations") == b.count("iterations") and a.count("regionCount") == b.count("regionCount")
        for ax in range(3):
            assert np.array_equal(val[ax], valb[ax]) and np.array_equal(v[ax], vb[ax]), f"axis {ax} differs from a fresh handle"
        a.close(); b.close()


def test_handle_reused_for_another_scene_equals_a_fresh_handle(built):
    _reuse_for_another_scene(parity.EMUL_LIB)


@pytest.mark.gpu
def test_gpu_handle_reused_for_another_scene_equals_a_fresh_handle(built):
    _reuse_for_another_scene(None)
