"""ctypes binding of oracle/_ref/libps_ref_classify.so -- the reference's OWN classifier and matrix-block construction
(exec/HDK_PolyStokesSolver_Classifier.cpp and exec/HDK_PolyStokesSolver_ConstructMatrixBlocks.cpp, compiled unmodified from
/root/reference on oracle/hdk_shim + oracle/eigen_facade; recipe `make -C oracle ref`, harness oracle/ref_classify.cpp).
TEST INFRASTRUCTURE: imported only by tests/."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libps_ref_classify.so")
COUNT_NAMES = ("nCenter", "nFaceX", "nFaceY", "nFaceZ", "nEdgeYZ", "nEdgeXZ", "nEdgeXY", "regionCount")
COEFF_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double))


class _Params(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("dx", C.c_double), ("dt", C.c_double), ("liquidLayers", C.c_int32),
                ("solidLayers", C.c_int32), ("doReducedRegions", C.c_int32), ("doTile", C.c_int32), ("tileSize", C.c_int32), ("tilePadding", C.c_int32),
                ("density", C.c_double), ("solverType", C.c_int32)]


def available():
    return os.path.exists(LIB_PATH)


def slot_shape(slot, nx, ny, nz):
    """numpy shape (z, y, x) of sample slot 0 centre, 1..3 faces x / y / z, 4..6 edges of axis 0 (YZ) / 1 (XZ) / 2 (XY)."""
    ex = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 1, 1), (1, 0, 1), (1, 1, 0)][slot]
    return (nz + ex[2], ny + ex[1], nx + ex[0])


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.refcls_create.restype = C.c_void_p; L.refcls_create.argtypes = [C.POINTER(_Params), C.c_void_p]
        L.refcls_destroy.argtypes = [C.c_void_p]
        L.refcls_results.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.refcls_construct_blocks.restype = C.c_int; L.refcls_construct_blocks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, COEFF_FN]
        L.refcls_assemble.restype = C.c_int; L.refcls_assemble.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.refcls_csr_dims.restype = C.c_int; L.refcls_csr_dims.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.refcls_csr_copy.restype = C.c_int; L.refcls_csr_copy.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.refcls_vector.restype = C.c_int64; L.refcls_vector.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        _lib = L
    return _lib


class RefClassifier:
    """The reference's classifier run on a set of integration weights; optionally its matrix-block construction on top.

    `weight_field(liquid, slot)` returns the float32 weight array of a slot (liquid = 1: liquid weights, 0: fluid = non-solid
    weights), e.g. Oracle.weight_field or PolyStokesSolver.weight_field.  `params`: the scene's parameter dict."""

    def __init__(self, nx, ny, nz, dx, dt, params, weight_field, density=1000.0, solver_type=0):
        self.nx, self.ny, self.nz = nx, ny, nz
        W = [np.ascontiguousarray(weight_field(liq, slot), dtype=np.float32) for liq in (1, 0) for slot in range(7)]
        for i, w in enumerate(W):
            assert w.shape == slot_shape(i % 7, nx, ny, nz), (i, w.shape)
        p = _Params(nx, ny, nz, float(dx), float(dt), int(params["liquidLayers"]), int(params["solidLayers"]), int(params["doReduced"]), int(params["doTile"]),
                    int(params["tileSize"]), int(params["tilePadding"]), float(density), int(solver_type))
        wp = (C.c_void_p * 14)(*[w.ctypes.data for w in W])
        self.h = lib().refcls_create(C.byref(p), wp)
        if not self.h:
            raise RuntimeError("refcls_create failed")

    def close(self):
        if getattr(self, "h", None):
            lib().refcls_destroy(self.h)
            self.h = None

    __del__ = close

    def results(self, want_valid=True):
        """(fields[kind][slot] int64 with kind 0 labels / 1 active indices / 2 reduced indices, counts dict, valid[3] float32 or None)"""
        nx, ny, nz = self.nx, self.ny, self.nz
        out = [np.zeros(slot_shape(slot, nx, ny, nz), dtype=np.int64) for kind in range(3) for slot in range(7)]
        valid = [np.zeros(slot_shape(1 + a, nx, ny, nz), dtype=np.float32) for a in range(3)] if want_valid else None
        counts = np.zeros(8, dtype=np.int64)
        op = (C.c_void_p * 21)(*[a.ctypes.data for a in out])
        vp = (C.c_void_p * 3)(*[v.ctypes.data for v in valid]) if want_valid else None
        lib().refcls_results(self.h, op, counts.ctypes.data, vp)
        fields = [[out[kind * 7 + slot] for slot in range(7)] for kind in range(3)]
        return fields, dict(zip(COUNT_NAMES, (int(c) for c in counts))), valid

    def construct_blocks(self, vel, colvel, viscosity, com, coeff_fn):
        """constructMatrixBlocks.  vel / colvel: 3 face arrays, viscosity: centre array (float32); com: (R, 3) float64 centres of mass;
        coeff_fn(offset3, axis) -> 26 basis values (buildConversionCoefficients, S.cpp:2107-2149)."""
        keep = [np.ascontiguousarray(a, dtype=np.float32) for a in list(vel) + list(colvel) + [viscosity]]
        com = np.ascontiguousarray(com, dtype=np.float64).reshape(-1)

        if isinstance(coeff_fn, COEFF_FN):          # a C function (e.g. oracle_coeff_fn()): no Python in the loop
            self._cb = coeff_fn
        else:
            def cb(off, axis, out):
                vals = coeff_fn(np.array([off[0], off[1], off[2]]), int(axis))
                for i in range(26):
                    out[i] = vals[i]
            self._cb = COEFF_FN(cb)
        vp = (C.c_void_p * 3)(*[a.ctypes.data for a in keep[0:3]]); cp = (C.c_void_p * 3)(*[a.ctypes.data for a in keep[3:6]])
        rc = lib().refcls_construct_blocks(self.h, vp, cp, keep[6].ctypes.data, com.ctypes.data if com.size else None, self._cb)
        if rc != 0:
            raise RuntimeError(f"refcls_construct_blocks returned {rc}")

    def assemble(self, mass_dense, visc_dense, best_fit):
        """assemble() (AssembleSystem.cpp / AssembleBlocks.cpp) after construct_blocks.  mass_dense / visc_dense: (R, 26, 26) row-major,
        best_fit: (R, 26) -- the region algebra of exec/HDK_PolyStokesSolver.cpp, supplied by the caller."""
        m = np.ascontiguousarray(mass_dense, dtype=np.float64).reshape(-1); v = np.ascontiguousarray(visc_dense, dtype=np.float64).reshape(-1)
        f = np.ascontiguousarray(best_fit, dtype=np.float64).reshape(-1)
        rc = lib().refcls_assemble(self.h, m.ctypes.data if m.size else None, v.ctypes.data if v.size else None, f.ctypes.data if f.size else None)
        if rc != 0:
            raise RuntimeError(f"refcls_assemble returned {rc}")

    def csr(self, name):
        r, c, n = C.c_int64(), C.c_int64(), C.c_int64()
        if lib().refcls_csr_dims(self.h, name.encode(), C.byref(r), C.byref(c), C.byref(n)) != 0:
            raise KeyError(name)
        ptr = np.empty(r.value + 1, dtype=np.int64); idx = np.empty(n.value, dtype=np.int32); val = np.empty(n.value, dtype=np.float64)
        lib().refcls_csr_copy(self.h, name.encode(), ptr.ctypes.data, idx.ctypes.data, val.ctypes.data)
        return (r.value, c.value), ptr, idx, val

    def vector(self, name):
        n = lib().refcls_vector(self.h, name.encode(), None)
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, dtype=np.float64)
        lib().refcls_vector(self.h, name.encode(), out.ctypes.data)
        return out


def oracle_coeff_fn():
    """The oracle's restatement of buildConversionCoefficients (S.cpp:2107-2149) as a C function pointer for construct_blocks."""
    from oracle import oracle as orc
    return C.cast(orc.lib().orc_conversion_coefficients, COEFF_FN)


def classify(nx, ny, nz, dx, dt, params, weight_field, want_valid=True):
    """One-shot classification: see RefClassifier.results."""
    r = RefClassifier(nx, ny, nz, dx, dt, params, weight_field)
    try:
        return r.results(want_valid)
    finally:
        r.close()
