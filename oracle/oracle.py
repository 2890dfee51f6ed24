"""ctypes front-end of the CPU parity oracle.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module; the product (``polystokes_b200``) never does.

PARITY UNPINNED: see the header of ``ps_oracle.hpp``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SLOTS = ["center", "faceX", "faceY", "faceZ", "edgeYZ", "edgeXZ", "edgeXY"]


def slot_shape(slot, nx, ny, nz):
    """numpy shape (z, y, x) of the field sampled at ``slot`` (S.h:193-222 / BASELINE.md section 3)."""
    r = [nx, ny, nz]
    if 1 <= slot <= 3:
        r[slot - 1] += 1
    elif slot >= 4:
        e = slot - 4
        for a in range(3):
            if a != e:
                r[a] += 1
    return (r[2], r[1], r[0])


class _Params(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("dx", C.c_double), ("dt", C.c_double), ("density", C.c_double), ("tolerance", C.c_double),
                ("maxIterations", C.c_int), ("liquidLayers", C.c_int), ("solidLayers", C.c_int),
                ("doReduced", C.c_int), ("doTile", C.c_int), ("tileSize", C.c_int), ("tilePadding", C.c_int),
                ("threads", C.c_int)]


def build(force=False):
    so = os.path.join(_HERE, "libps_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(_Params)]
        L.orc_destroy.argtypes = [C.c_void_p]
        fp = C.POINTER(C.c_float)
        L.orc_set_inputs.argtypes = [C.c_void_p] + [fp] * 9
        L.orc_set_weights.argtypes = [C.c_void_p, C.c_int, C.c_int, fp]
        for f in ("orc_setup", "orc_build_weights", "orc_classify", "orc_assemble_explicit_A", "orc_construct_guess"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_solve.argtypes = [C.c_void_p]
        L.orc_solve.restype = C.c_int
        L.orc_solve_eigen_cg.argtypes = [C.c_void_p]
        L.orc_solve_eigen_cg.restype = C.c_int
        L.orc_apply.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_writeback.argtypes = [C.c_void_p] + [fp] * 6
        L.orc_count.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_count.restype = C.c_int64
        L.orc_real.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_real.restype = C.c_double
        L.orc_index_field.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int64)]
        L.orc_index_field.restype = C.c_int64
        L.orc_weight_field.argtypes = [C.c_void_p, C.c_int, C.c_int, fp]
        L.orc_weight_field.restype = C.c_int64
        L.orc_csr_dims.argtypes = [C.c_void_p, C.c_char_p] + [C.POINTER(C.c_int64)] * 3
        L.orc_csr_copy.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_double)]
        L.orc_csr_save.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.orc_vector.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double)]
        L.orc_vector.restype = C.c_int64
        L.orc_vector_save.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.orc_inverse_partial_piv_lu.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
        L.orc_solve_full_piv_lu.argtypes = [C.POINTER(C.c_double)] * 3 + [C.c_int]
        L.orc_solve_full_piv_lu.restype = C.c_int
        L.orc_conversion_coefficients.argtypes = [C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double)]
        L.orc_num_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Oracle:
    """One oracle instance = one ``HDK_PolyStokes::Solver`` (exec/HDK_PolyStokesSolver.h:27-43)."""

    LABELS, ACTIVE, REDUCED = 0, 1, 2

    def __init__(self, scene, threads=0, **overrides):
        p = dict(scene.params)
        p.update(overrides)
        self.scene = scene
        self.p = p
        self.nx, self.ny, self.nz = scene.nx, scene.ny, scene.nz
        P = _Params(scene.nx, scene.ny, scene.nz, scene.dx, scene.dt, scene.density, p["tolerance"],
                    p["maxIterations"], p["liquidLayers"], p["solidLayers"],
                    p["doReduced"], p["doTile"], p["tileSize"], p["tilePadding"], threads)
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_create(C.byref(P)))
        self._keep = [np.ascontiguousarray(a, dtype=np.float32) for a in
                      [scene.surface, scene.collision, scene.viscosity] + list(scene.vel) + list(scene.colvel)]
        self.L.orc_set_inputs(self.h, *[_fp(a) for a in self._keep])

    def __del__(self):
        try:
            if self.h:
                self.L.orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def setup(self):
        self.L.orc_setup(self.h)
        return self

    def build_weights(self):
        self.L.orc_build_weights(self.h)

    def classify(self):
        self.L.orc_classify(self.h)

    def assemble_explicit_A(self):
        self.L.orc_assemble_explicit_A(self.h)

    def solve(self):
        return self.L.orc_solve(self.h)

    def construct_guess(self):
        """useWarmStart branch (PS.C:465-467): fills the vector "guess"."""
        self.L.orc_construct_guess(self.h)

    def solve_eigen_cg(self):
        """solverType EIGEN (S.cpp:814-862): Jacobi-preconditioned Eigen CG on the explicit A, started from "guess"."""
        return self.L.orc_solve_eigen_cg(self.h)

    def apply(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        self.L.orc_apply(self.h, _dp(x), _dp(y))
        return y

    def writeback(self):
        vel = [np.array(v, dtype=np.float32, copy=True) for v in self.scene.vel]
        valid = [np.zeros_like(v) for v in vel]
        self.L.orc_writeback(self.h, *[_fp(a) for a in vel + valid])
        return vel, valid

    def count(self, name):
        v = self.L.orc_count(self.h, name.encode())
        if v == -2 ** 63:
            raise KeyError(name)
        return int(v)

    def real(self, name):
        return float(self.L.orc_real(self.h, name.encode()))

    def index_field(self, kind, slot):
        out = np.empty(slot_shape(slot, self.nx, self.ny, self.nz), dtype=np.int64)
        self.L.orc_index_field(self.h, kind, slot, out.ctypes.data_as(C.POINTER(C.c_int64)))
        return out

    def weight_field(self, liquid, slot):
        out = np.empty(slot_shape(slot, self.nx, self.ny, self.nz), dtype=np.float32)
        self.L.orc_weight_field(self.h, int(liquid), slot, _fp(out))
        return out

    def csr(self, name):
        r, c, n = C.c_int64(), C.c_int64(), C.c_int64()
        if self.L.orc_csr_dims(self.h, name.encode(), C.byref(r), C.byref(c), C.byref(n)) != 0:
            raise KeyError(name)
        ptr = np.empty(r.value + 1, dtype=np.int64)
        idx = np.empty(n.value, dtype=np.int32)
        val = np.empty(n.value, dtype=np.float64)
        self.L.orc_csr_copy(self.h, name.encode(), ptr.ctypes.data_as(C.POINTER(C.c_int64)),
                            idx.ctypes.data_as(C.POINTER(C.c_int32)), _dp(val))
        return (r.value, c.value), ptr, idx, val

    def scipy_csr(self, name):
        import scipy.sparse as sp
        shape, ptr, idx, val = self.csr(name)
        return sp.csr_matrix((val, idx, ptr), shape=shape)

    def vector(self, name):
        n = self.L.orc_vector(self.h, name.encode(), None)
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, dtype=np.float64)
        self.L.orc_vector(self.h, name.encode(), _dp(out))
        return out

    def save_csr(self, name, path):
        return self.L.orc_csr_save(self.h, name.encode(), path.encode())

    def save_vector(self, name, path):
        return self.L.orc_vector_save(self.h, name.encode(), path.encode())


def conversion_coefficients(offset, axis):
    off = np.ascontiguousarray(offset, dtype=np.float64)
    out = np.empty(26, dtype=np.float64)
    lib().orc_conversion_coefficients(_dp(off), axis, _dp(out))
    return out


def inverse_partial_piv_lu(A):
    A = np.ascontiguousarray(A, dtype=np.float64)
    out = np.empty_like(A)
    lib().orc_inverse_partial_piv_lu(_dp(A), _dp(out), A.shape[0])
    return out


def solve_full_piv_lu(A, rhs):
    A = np.ascontiguousarray(A, dtype=np.float64)
    rhs = np.ascontiguousarray(rhs, dtype=np.float64)
    x = np.empty_like(rhs)
    rank = lib().orc_solve_full_piv_lu(_dp(A), _dp(rhs), _dp(x), A.shape[0])
    return x, rank
