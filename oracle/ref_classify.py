"""ctypes binding of oracle/_ref/libps_ref_classify.so -- the reference's OWN classifier
(exec/HDK_PolyStokesSolver_Classifier.cpp, compiled unmodified from /root/reference on oracle/hdk_shim + oracle/eigen_facade;
recipe `make -C oracle ref`, harness oracle/ref_classify.cpp).  TEST INFRASTRUCTURE: imported only by tests/."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libps_ref_classify.so")
COUNT_NAMES = ("nCenter", "nFaceX", "nFaceY", "nFaceZ", "nEdgeYZ", "nEdgeXZ", "nEdgeXY", "regionCount")


class _Params(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("dx", C.c_double), ("dt", C.c_double), ("liquidLayers", C.c_int32),
                ("solidLayers", C.c_int32), ("doReducedRegions", C.c_int32), ("doTile", C.c_int32), ("tileSize", C.c_int32), ("tilePadding", C.c_int32)]


def available():
    return os.path.exists(LIB_PATH)


def slot_shape(slot, nx, ny, nz):
    """numpy shape (z, y, x) of sample slot 0 centre, 1..3 faces x / y / z, 4..6 edges of axis 0 (YZ) / 1 (XZ) / 2 (XY)."""
    ex = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 1, 1), (1, 0, 1), (1, 1, 0)][slot]
    return (nz + ex[2], ny + ex[1], nx + ex[0])


def classify(nx, ny, nz, dx, dt, params, weight_field, want_valid=True):
    """Runs the reference classifier.  `weight_field(liquid, slot)` returns the float32 weight array of a slot (liquid = 1: liquid
    weights, 0: fluid = non-solid weights), e.g. Oracle.weight_field or PolyStokesSolver.weight_field.  `params`: the scene's
    parameter dict (liquidLayers, solidLayers, doReduced, doTile, tileSize, tilePadding).
    Returns (fields[kind][slot] int64 with kind 0 labels / 1 active indices / 2 reduced indices, counts dict, valid[3] float32 or None)."""
    L = C.CDLL(LIB_PATH)
    L.refcls_run.restype = C.c_int
    W = [np.ascontiguousarray(weight_field(liq, slot), dtype=np.float32) for liq in (1, 0) for slot in range(7)]
    for i, w in enumerate(W):
        assert w.shape == slot_shape(i % 7, nx, ny, nz), (i, w.shape)
    out = [np.zeros(slot_shape(slot, nx, ny, nz), dtype=np.int64) for kind in range(3) for slot in range(7)]
    valid = [np.zeros(slot_shape(1 + a, nx, ny, nz), dtype=np.float32) for a in range(3)] if want_valid else None
    counts = np.zeros(8, dtype=np.int64)
    p = _Params(nx, ny, nz, float(dx), float(dt), int(params["liquidLayers"]), int(params["solidLayers"]), int(params["doReduced"]), int(params["doTile"]),
                int(params["tileSize"]), int(params["tilePadding"]))
    wp = (C.c_void_p * 14)(*[w.ctypes.data for w in W])
    op = (C.c_void_p * 21)(*[a.ctypes.data for a in out])
    vp = (C.c_void_p * 3)(*[v.ctypes.data for v in valid]) if want_valid else None
    rc = L.refcls_run(C.byref(p), wp, op, counts.ctypes.data_as(C.c_void_p), vp)
    if rc != 0:
        raise RuntimeError(f"refcls_run returned {rc}")
    fields = [[out[kind * 7 + slot] for slot in range(7)] for kind in range(3)]
    return fields, dict(zip(COUNT_NAMES, (int(c) for c in counts))), valid
