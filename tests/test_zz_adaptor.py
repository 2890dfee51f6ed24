"""(Named test_zz_* so that it runs last.)  The drop-in boundary with the reference's own types: integration/hdk_polystokes_b200_adaptor.cpp -- the body a maintainer puts in
HDK_PolyStokes::solveGasSubclass instead of PS.C:329-584 -- compiled against the HDK stand-in and the reference's exec/HDK_PolyStokes.h
(oracle/_ref/libps_ref_adaptor.so), cooked on the SIM fields of a scene, next to the compiled reference solver (libps_ref_full.so) cooked
on the same fields: same `valid` field, same velocity within the parity gate, same result code.  CPU: the emulation twin stands behind the
C ABI; GPU: the product library."""
import ctypes as C
import os

import numpy as np
import pytest

import parity
from oracle import ref_full
from polystokes_b200 import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADAPTOR = os.path.join(ROOT, "oracle", "_ref", "libps_ref_adaptor.so")
PRODUCT = os.path.join(ROOT, "polystokes_b200", "libpolystokes_b200.so")
pytestmark = pytest.mark.skipif(not (os.path.exists(ADAPTOR) and ref_full.available()), reason="oracle/_ref not built (needs /root/reference)")


class _P(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("dx", C.c_double), ("dt", C.c_double), ("density", C.c_double), ("tolerance", C.c_double),
                ("maxIterations", C.c_int32), ("liquidLayers", C.c_int32), ("solidLayers", C.c_int32), ("doReducedRegions", C.c_int32), ("doTile", C.c_int32),
                ("tileSize", C.c_int32), ("tilePadding", C.c_int32), ("solverType", C.c_int32), ("useWarmStart", C.c_int32), ("keepNonConvergedResults", C.c_int32)]


def cook(sc, backend, steps=1, devices=1, do_solve=1):
    L = C.CDLL(ADAPTOR)
    L.refadp_run2.restype = C.c_int
    assert L.refadp_bind(backend.encode()) == 0, f"could not bind the C ABI of {backend}"      # dlopen(RTLD_LOCAL): no symbol leaks between libraries
    p = sc.params
    P = _P(sc.nx, sc.ny, sc.nz, float(sc.dx), float(sc.dt), float(sc.density), float(p["tolerance"]), int(p["maxIterations"]), int(p["liquidLayers"]), int(p["solidLayers"]),
           int(p["doReduced"]), int(p["doTile"]), int(p["tileSize"]), int(p["tilePadding"]), int(p.get("solverType", 0)), int(p.get("useWarmStart", 0)),
           int(p.get("keepNonConvergedResults", 0)))
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    keep = [f32(sc.surface), f32(sc.collision), f32(sc.viscosity)] + [f32(v) for v in sc.vel] + [f32(v) for v in sc.colvel]
    vel = [np.zeros_like(v) for v in keep[3:6]]; valid = [np.zeros_like(v) for v in keep[3:6]]
    arr = lambda xs: (C.c_void_p * 3)(*[x.ctypes.data for x in xs])
    err = C.create_string_buffer(512)
    rc = L.refadp_run2(C.byref(P), C.c_void_p(keep[0].ctypes.data), C.c_void_p(keep[1].ctypes.data), C.c_void_p(keep[2].ctypes.data), arr(keep[3:6]), arr(keep[6:9]),
                       C.c_int(steps), arr(vel), arr(valid), err, C.c_int(512), C.c_int(devices), C.c_int(do_solve))
    return rc, vel, valid, err.value.decode()


def _check(sc, backend, steps=1, devices=1):
    rc, vel, valid, err = cook(sc, backend, steps, devices)
    R = ref_full.RefFull(sc).setup()
    rr = R.solve()
    rvel, rvalid = R.writeback()
    assert rc == rr, f"node result {rc} ({err}) vs the compiled reference {rr}"
    tol = max(10 * sc.params["tolerance"], 4e-7)
    for a in range(3):
        assert np.array_equal(valid[a], rvalid[a]), f"valid field axis {a}"
        assert float(np.abs(vel[a] - rvel[a]).max()) <= tol * max(float(np.abs(rvel[a]).max()), 1e-30), f"velocity axis {a}"
        untouched = rvalid[a] == 0
        assert np.array_equal(vel[a][untouched], np.asarray(sc.vel[a], dtype=np.float32)[untouched]), "invalid faces must keep their input velocity"


@pytest.mark.parametrize("case", ["blob32_tile8", "box24_uniform", "ragged_notile"])
def test_adaptor_on_emulated_library_matches_compiled_reference(built, case):
    sc = {"blob32_tile8": lambda: scenes.blob_scene(32, seed=3), "box24_uniform": lambda: scenes.box_scene(24, doReduced=0, tolerance=1e-6),
          "ragged_notile": lambda: scenes.blob_scene((28, 36, 24), seed=7, doTile=0)}[case]()
    _check(sc, parity.EMUL_LIB, steps=2 if case == "blob32_tile8" else 1)


def test_adaptor_reports_nonconvergence_like_the_node(built):
    sc = scenes.blob_scene(24, seed=5, maxIterations=3, tolerance=1e-12)
    rc, vel, valid, err = cook(sc, parity.EMUL_LIB)
    assert rc == 0 and "did not converge" in err                     # PS.C:597-600
    for a in range(3):
        assert np.array_equal(vel[a], np.asarray(sc.vel[a], dtype=np.float32)), "a non-converged step must leave the velocity untouched (PS.C:566)"


@pytest.mark.parametrize("keep", [1, 0])
def test_adaptor_without_do_solve_matches_the_node(built, keep):
    """"Do Solve" off (PS.C:513, 565-605): no solve, no error message; with keepNonConvergedResults the zero solution is recovered and written
    back (u = Mc^-1 rhs_u on active faces, the least-squares polynomial on reduced ones), without it the velocity stays as it came."""
    sc = scenes.blob_scene(32, seed=3, keepNonConvergedResults=keep)
    rc, vel, valid, err = cook(sc, parity.EMUL_LIB, do_solve=0)
    R = ref_full.RefFull(sc).setup()
    rr = R.skip_solve()
    rvel, rvalid = R.writeback()
    assert rc == rr == -3 and err == ""
    for a in range(3):
        assert np.array_equal(valid[a], rvalid[a]), f"valid field axis {a}"
        assert float(np.abs(vel[a] - rvel[a]).max()) <= 4e-7 * max(float(np.abs(rvel[a]).max()), 1e-30), f"velocity axis {a}"


@pytest.mark.gpu
def test_adaptor_multi_gpu_handle(built):
    """the node state asks for two devices: one ps_create_multi handle behind the same single cook thread"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _check(scenes.scene_s3(64), PRODUCT, steps=2, devices=2)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [40, 64])
def test_adaptor_on_gpu_library_matches_compiled_reference(built, n):
    _check(scenes.blob_scene(n, seed=3) if n == 40 else scenes.scene_s3(64), PRODUCT, steps=2)
