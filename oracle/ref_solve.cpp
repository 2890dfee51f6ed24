// ref_solve.cpp -- TEST INFRASTRUCTURE (parity oracle).  C entry points around the REFERENCE'S OWN solve stage:
//   lib/include/ApplyPressureStressMatrix.h  (setupMatrixVectorProducts :24-68, applyMatrixVectorProducts :102-179)
//   lib/include/pcg.h                         (pcg_external_matrix_A :268-340, bicgstab_external_matrix_A :134-200)
//   lib/include/util.h, lib/include/units.h   (concatenate_*, manualMatrixTransposeVectorDistribute2, the typedefs)
//   extern/eigen/unsupported/Eigen/src/SparseExtra/MarketIO.h   (saveMarket / saveMarketVector, the export format)
// compiled UNMODIFIED from /root/reference by `make -C oracle ref` into oracle/_ref/libps_ref_solve.so.  No reference
// source is copied into this repository: the four headers are found through -I/root/reference/lib/include.  Eigen, TBB
// and the HDK are absent offline; oracle/eigen_facade/ supplies the slice of the Eigen API these headers use
// (see its Eigen/Core for what that does and does not pin).  The call sequence below is that of
// HDK_PolyStokes::Solver::solveSPDwithMatrixVectorPCG (exec/HDK_PolyStokesSolver.cpp:734-812); the preconditioner is
// the reference's IDENTITY one (lib/src/Preconditioner.cpp:271-274 returns its argument), restated here because
// Preconditioner.cpp needs the HDK to compile.
#include <cstdint>
#include <cstring>
#include <new>

#include "pcg.h"
#include "ApplyPressureStressMatrix.h"

namespace {
struct IdentityPreconditioner { Vector solve(Vector b) { return b; } };

struct RefSolve {
    ApplyPressureStressMatrix<>* applyMatrix = nullptr;
    IdentityPreconditioner* presolver = nullptr;
    Index n = 0;
    ~RefSolve() { delete applyMatrix; delete presolver; }
};

SparseMatrix load(int64_t rows, int64_t cols, const int64_t* ptr, const int32_t* idx, const double* val) {
    SparseMatrix m;
    m.setFromCompressed((Eigen::Index)rows, (Eigen::Index)cols, ptr, idx, val);
    return m;
}
Vector to_vec(const double* p, Index n) { Vector v(n); for (Index i = 0; i < n; ++i) v(i) = p[i]; return v; }
}  // namespace

extern "C" {

struct ref_csr { int64_t rows, cols; const int64_t* ptr; const int32_t* idx; const double* val; };

// matrices in the order of setupMatrixVectorProducts: McInv, BInv (= Inv_Mr_plus_2JDtuDJ), uInv, G, JG, Dt, JDt
void* refsolve_create(double dt, const ref_csr* McInv, const ref_csr* BInv, const ref_csr* uInv, const ref_csr* G, const ref_csr* JG, const ref_csr* Dt, const ref_csr* JDt) {
    RefSolve* h = new (std::nothrow) RefSolve;
    if (!h) return nullptr;
    auto L = [](const ref_csr* c) { return load(c->rows, c->cols, c->ptr, c->idx, c->val); };
    h->applyMatrix = new ApplyPressureStressMatrix<>();
    h->applyMatrix->setupMatrixVectorProducts(dt, 1. / dt, L(McInv), L(BInv), L(uInv), L(G), L(JG), L(Dt), L(JDt));
    h->presolver = new IdentityPreconditioner;
    h->n = (Index)(G->cols + Dt->cols);
    return h;
}
void refsolve_destroy(void* hv) { delete (RefSolve*)hv; }
int64_t refsolve_size(void* hv) { return (int64_t)((RefSolve*)hv)->n; }

// y = A x through ApplyPressureStressMatrix::apply (Apply.h:182-184)
void refsolve_apply(void* hv, const double* x, double* y) {
    RefSolve* h = (RefSolve*)hv;
    Vector r = h->applyMatrix->apply(to_vec(x, h->n));
    for (Index i = 0; i < h->n; ++i) y[i] = r(i);
}

// which = 0: pcg_external_matrix_A, 1: bicgstab_external_matrix_A; zero start as S.cpp:768 / :788.  Returns the function's
// return value (the iteration index at convergence, or maxIter); *rre as the function leaves it.
int refsolve_solve(void* hv, int which, const double* b, double tol, unsigned int maxIter, double* xOut, double* rre) {
    RefSolve* h = (RefSolve*)hv;
    Vector solutionVector(h->n), bv = to_vec(b, h->n), tmp_r, tmp_z, tmp_p, tmp_Ap;
    solutionVector.setZero();
    double totalTimerAapply = 0., totalTimerOther = 0., err = 0.;
    int it;
    if (which == 0)
        it = pcg_external_matrix_A(solutionVector, h->applyMatrix, bv, tmp_r, tmp_z, tmp_p, tmp_Ap, h->presolver, totalTimerAapply, totalTimerOther, err, tol, maxIter);
    else
        it = bicgstab_external_matrix_A(solutionVector, h->applyMatrix, bv, tmp_r, tmp_z, tmp_p, tmp_Ap, h->presolver, totalTimerAapply, totalTimerOther, err, tol, maxIter);
    for (Index i = 0; i < h->n; ++i) xOut[i] = solutionVector(i);
    if (rre) *rre = err;
    return it;
}

// The checkout's own MatrixMarket writer (extern/eigen/unsupported/Eigen/src/SparseExtra/MarketIO.h), called the way
// S.cpp:533-606 calls it: Eigen::saveMarket(matrix, path) and Eigen::saveMarketVector(vector, path).  Returns 1 on success.
int refsolve_save_market(const ref_csr* m, const char* path) { return Eigen::saveMarket(load(m->rows, m->cols, m->ptr, m->idx, m->val), std::string(path)) ? 1 : 0; }
int refsolve_save_market_vector(const double* v, int64_t n, const char* path) { return Eigen::saveMarketVector(to_vec(v, (Index)n), std::string(path)) ? 1 : 0; }

}  // extern "C"
