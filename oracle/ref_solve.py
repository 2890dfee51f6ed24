"""ctypes binding of oracle/_ref/libps_ref_solve.so -- the reference's OWN solve stage (lib/include/pcg.h,
lib/include/ApplyPressureStressMatrix.h, util.h, units.h) compiled unmodified from /root/reference on the Eigen facade
of oracle/eigen_facade (recipe: `make -C oracle ref`; see oracle/ref_solve.cpp).  TEST INFRASTRUCTURE: imported only by
tests/ and by bench.py's cpu_baseline / --impl reference legs.  The library is built where /root/reference exists and
travels to the GPU box as a built artefact (git-ignored, not gpurun-ignored)."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libps_ref_solve.so")


class _Csr(C.Structure):
    _fields_ = [("rows", C.c_int64), ("cols", C.c_int64), ("ptr", C.c_void_p), ("idx", C.c_void_p), ("val", C.c_void_p)]


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        P = C.POINTER(_Csr)
        L.refsolve_create.argtypes = [C.c_double, P, P, P, P, P, P, P]; L.refsolve_create.restype = C.c_void_p
        L.refsolve_destroy.argtypes = [C.c_void_p]
        L.refsolve_size.argtypes = [C.c_void_p]; L.refsolve_size.restype = C.c_int64
        L.refsolve_apply.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.refsolve_solve.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_uint, C.c_void_p, C.POINTER(C.c_double)]; L.refsolve_solve.restype = C.c_int
        L.refsolve_save_market.argtypes = [C.POINTER(_Csr), C.c_char_p]; L.refsolve_save_market.restype = C.c_int
        L.refsolve_save_market_vector.argtypes = [C.c_void_p, C.c_int64, C.c_char_p]; L.refsolve_save_market_vector.restype = C.c_int
        _lib = L
    return _lib


def save_market(csr, path):
    """Eigen::saveMarket (the checkout's unsupported/Eigen/src/SparseExtra/MarketIO.h:311-335) on ((rows, cols), ptr, idx, val)."""
    (r, c), ptr, idx, val = csr
    ptr = np.ascontiguousarray(ptr, dtype=np.int64); idx = np.ascontiguousarray(idx, dtype=np.int32); val = np.ascontiguousarray(val, dtype=np.float64)
    m = _Csr(int(r), int(c), ptr.ctypes.data, idx.ctypes.data, val.ctypes.data)
    if lib().refsolve_save_market(C.byref(m), str(path).encode()) != 1:
        raise IOError(f"saveMarket could not write {path}")


def save_market_vector(v, path):
    """Eigen::saveMarketVector (MarketIO.h:347-382)."""
    v = np.ascontiguousarray(v, dtype=np.float64)
    if lib().refsolve_save_market_vector(v.ctypes.data, v.size, str(path).encode()) != 1:
        raise IOError(f"saveMarketVector could not write {path}")


class RefSolve:
    """The reference's ApplyPressureStressMatrix + pcg.h loops on a set of component matrices.

    `csr(name)` must return ((rows, cols), rowptr int64, colidx int32, values float64) for the names McInv, BInv, uInv,
    G, JG, Dt, JDt -- both oracle.Oracle.csr and polystokes_b200.PolyStokesSolver.csr do."""

    ORDER = ("McInv", "BInv", "uInv", "G", "JG", "Dt", "JDt")      # setupMatrixVectorProducts, S.cpp:741-751

    def __init__(self, csr, dt):
        self._keep = []
        args = []
        for name in self.ORDER:
            (r, c), ptr, idx, val = csr(name)
            ptr = np.ascontiguousarray(ptr, dtype=np.int64); idx = np.ascontiguousarray(idx, dtype=np.int32); val = np.ascontiguousarray(val, dtype=np.float64)
            self._keep += [ptr, idx, val]
            args.append(_Csr(int(r), int(c), ptr.ctypes.data, idx.ctypes.data, val.ctypes.data))
        self.h = lib().refsolve_create(float(dt), *[C.byref(a) for a in args])
        if not self.h:
            raise MemoryError("refsolve_create failed")
        self.n = int(lib().refsolve_size(self.h))

    def close(self):
        if self.h:
            lib().refsolve_destroy(self.h)
            self.h = None

    __del__ = close

    def apply(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.size == self.n
        y = np.empty(self.n, dtype=np.float64)
        lib().refsolve_apply(self.h, x.ctypes.data, y.ctypes.data)
        return y

    def _solve(self, which, b, tol, max_iter):
        b = np.ascontiguousarray(b, dtype=np.float64)
        assert b.size == self.n
        x = np.empty(self.n, dtype=np.float64)
        rre = C.c_double(0.0)
        it = lib().refsolve_solve(self.h, which, b.ctypes.data, float(tol), int(max_iter), x.ctypes.data, C.byref(rre))
        return int(it), x, float(rre.value)

    def pcg(self, b, tol, max_iter):
        """pcg_external_matrix_A (pcg.h:268-340): (returned iteration index, x, rre)."""
        return self._solve(0, b, tol, max_iter)

    def bicgstab(self, b, tol, max_iter):
        """bicgstab_external_matrix_A (pcg.h:134-200)."""
        return self._solve(1, b, tol, max_iter)

    def solve_spd(self, b, tol, max_iter):
        """solveSPDwithMatrixVectorPCG (S.cpp:734-812): CG from zero, BiCGSTAB from zero if CG used up its iterations.
        Returns (result, iterations, x, error, usedBiCGStab) with result 1 = SUCCESS, 0 = NOCONVERGE (S.h:61-70)."""
        it, x, err = self.pcg(b, tol, max_iter)
        used = 0
        if it == max_iter:
            used = 1
            it, x, err = self.bicgstab(b, tol, max_iter)
        return (0 if it == max_iter else 1), it, x, err, used
