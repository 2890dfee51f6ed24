"""ctypes binding of oracle/_ref/libps_ref_full.so -- the reference's COMPLETE per-step solver (all six
exec/HDK_PolyStokesSolver*.cpp + lib/src/Preconditioner.cpp + lib/include, compiled unmodified from /root/reference on
oracle/hdk_shim + oracle/eigen_facade; recipe `make -C oracle ref`, harness oracle/ref_full.cpp).  The interface mirrors
oracle.oracle.Oracle so the same checks run against either.  TEST INFRASTRUCTURE: imported only by tests/ and by bench.py's
cpu_baseline / --impl reference legs."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libps_ref_full.so")


class _Params(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("dx", C.c_double), ("dt", C.c_double), ("density", C.c_double), ("tolerance", C.c_double),
                ("maxIterations", C.c_int32), ("liquidLayers", C.c_int32), ("solidLayers", C.c_int32), ("doReducedRegions", C.c_int32), ("doTile", C.c_int32),
                ("tileSize", C.c_int32), ("tilePadding", C.c_int32), ("solverType", C.c_int32), ("useWarmStart", C.c_int32), ("keepNonConvergedResults", C.c_int32)]


def available():
    return os.path.exists(LIB_PATH)


def slot_shape(slot, nx, ny, nz):
    ex = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 1, 1), (1, 0, 1), (1, 1, 0)][slot]
    return (nz + ex[2], ny + ex[1], nx + ex[0])


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.reffull_create.restype = C.c_void_p; L.reffull_create.argtypes = [C.POINTER(_Params), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.reffull_destroy.argtypes = [C.c_void_p]
        L.reffull_setup.restype = C.c_int; L.reffull_setup.argtypes = [C.c_void_p]
        L.reffull_solve.restype = C.c_int; L.reffull_solve.argtypes = [C.c_void_p]
        L.reffull_skip_solve.restype = C.c_int; L.reffull_skip_solve.argtypes = [C.c_void_p]
        L.reffull_index_field.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.reffull_weight_field.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.reffull_face_field.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.reffull_count.restype = C.c_int64; L.reffull_count.argtypes = [C.c_void_p, C.c_char_p]
        L.reffull_real.restype = C.c_double; L.reffull_real.argtypes = [C.c_void_p, C.c_char_p]
        L.reffull_csr_dims.restype = C.c_int; L.reffull_csr_dims.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.reffull_csr_copy.restype = C.c_int; L.reffull_csr_copy.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.reffull_vector.restype = C.c_int64; L.reffull_vector.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        L.reffull_export.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        _lib = L
    return _lib


class RefFull:
    """The compiled reference solver on a scene (polystokes_b200.scenes.Scene); same accessors as oracle.oracle.Oracle."""

    def __init__(self, scene, **overrides):
        p = dict(scene.params, **overrides)
        self.nx, self.ny, self.nz = scene.nx, scene.ny, scene.nz
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        self._keep = [f32(scene.surface), f32(scene.collision), f32(scene.viscosity)] + [f32(v) for v in scene.vel] + [f32(v) for v in scene.colvel]
        P = _Params(scene.nx, scene.ny, scene.nz, float(scene.dx), float(scene.dt), float(scene.density), float(p["tolerance"]), int(p["maxIterations"]),
                    int(p["liquidLayers"]), int(p["solidLayers"]), int(p["doReduced"]), int(p["doTile"]), int(p["tileSize"]), int(p["tilePadding"]),
                    int(p.get("solverType", 0)), int(p.get("useWarmStart", 0)), int(p.get("keepNonConvergedResults", 0)))
        vp = (C.c_void_p * 3)(*[a.ctypes.data for a in self._keep[3:6]]); cp = (C.c_void_p * 3)(*[a.ctypes.data for a in self._keep[6:9]])
        self.h = lib().reffull_create(C.byref(P), self._keep[0].ctypes.data, self._keep[1].ctypes.data, self._keep[2].ctypes.data, vp, cp)
        if not self.h:
            raise RuntimeError("reffull_create failed")

    def close(self):
        if getattr(self, "h", None):
            lib().reffull_destroy(self.h)
            self.h = None

    __del__ = close

    def setup(self):
        lib().reffull_setup(self.h)
        return self

    def solve(self):
        """solve + valid faces + velocity recovery and write-back; returns the SolverResult."""
        return int(lib().reffull_solve(self.h))

    def skip_solve(self):
        """the node with "Do Solve" off: valid faces + (keepNonConvergedResults) write-back of the zero solution; returns INCOMPLETE (-3)."""
        return int(lib().reffull_skip_solve(self.h))

    def writeback(self):
        vel = [np.empty(slot_shape(1 + a, self.nx, self.ny, self.nz), dtype=np.float32) for a in range(3)]
        valid = [np.empty(slot_shape(1 + a, self.nx, self.ny, self.nz), dtype=np.float32) for a in range(3)]
        for a in range(3):
            lib().reffull_face_field(self.h, 0, a, vel[a].ctypes.data); lib().reffull_face_field(self.h, 1, a, valid[a].ctypes.data)
        return vel, valid

    def count(self, name):
        v = lib().reffull_count(self.h, name.encode())
        if v == -2 ** 63:
            raise KeyError(name)
        return int(v)

    def real(self, name):
        return float(lib().reffull_real(self.h, name.encode()))

    def index_field(self, kind, slot):
        out = np.empty(slot_shape(slot, self.nx, self.ny, self.nz), dtype=np.int64)
        lib().reffull_index_field(self.h, kind, slot, out.ctypes.data)
        return out

    def weight_field(self, liquid, slot):
        out = np.empty(slot_shape(slot, self.nx, self.ny, self.nz), dtype=np.float32)
        lib().reffull_weight_field(self.h, int(liquid), slot, out.ctypes.data)
        return out

    def csr(self, name):
        r, c, n = C.c_int64(), C.c_int64(), C.c_int64()
        if lib().reffull_csr_dims(self.h, name.encode(), C.byref(r), C.byref(c), C.byref(n)) != 0:
            raise KeyError(name)
        ptr = np.empty(r.value + 1, dtype=np.int64); idx = np.empty(n.value, dtype=np.int32); val = np.empty(n.value, dtype=np.float64)
        lib().reffull_csr_copy(self.h, name.encode(), ptr.ctypes.data, idx.ctypes.data, val.ctypes.data)
        return (r.value, c.value), ptr, idx, val

    def vector(self, name):
        n = lib().reffull_vector(self.h, name.encode(), None)
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, dtype=np.float64)
        lib().reffull_vector(self.h, name.encode(), out.ctypes.data)
        return out

    def face_field(self, which, axis):
        """which 0: velocity (after solve(): the written-back field), 1: valid faces"""
        out = np.empty(slot_shape(1 + axis, self.nx, self.ny, self.nz), dtype=np.float32)
        lib().reffull_face_field(self.h, int(which), int(axis), out.ctypes.data)
        return out

    def export(self, prefix, what=7):
        lib().reffull_export(self.h, str(prefix).encode(), int(what))
