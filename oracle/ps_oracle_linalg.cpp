// ps_oracle_linalg.cpp -- TEST INFRASTRUCTURE ONLY (see ps_oracle.hpp header).
// Sparse/dense helpers restating the Eigen semantics the reference relies on.
#include "ps_oracle.hpp"
#include <fstream>
#include <limits>
#include <numeric>

namespace orc {

// extern/eigen/Eigen/src/SparseCore/SparseMatrix.h:1024-1054 (set_from_triplets) and
// 1122-1158 (collapseDuplicates): entries are bucketed by row keeping list order, duplicates
// (same row, same col) are summed into the first occurrence in list order, explicit zeros are
// kept, and the final row-major matrix has each row sorted by column.
void Csr::setFromTriplets(const std::vector<Triplet>& t) {
    ptr.assign(rows + 1, 0);
    for (const auto& e : t) { assert(e.r >= 0 && e.r < rows && e.c >= 0 && e.c < cols); ptr[e.r + 1]++; }
    for (exint r = 0; r < rows; ++r) ptr[r + 1] += ptr[r];
    std::vector<int> ci(t.size());
    std::vector<Real> cv(t.size());
    {
        std::vector<exint> cur(ptr.begin(), ptr.end() - 1);
        for (const auto& e : t) { exint p = cur[e.r]++; ci[p] = (int)e.c; cv[p] = e.v; }
    }
    std::vector<exint> nptr(rows + 1, 0);
    idx.clear(); val.clear(); idx.reserve(t.size()); val.reserve(t.size());
    std::vector<exint> ord;
    for (exint r = 0; r < rows; ++r) {
        exint b = ptr[r], e = ptr[r + 1];
        ord.resize(e - b);
        std::iota(ord.begin(), ord.end(), b);
        std::stable_sort(ord.begin(), ord.end(), [&](exint x, exint y) { return ci[x] < ci[y]; });
        for (size_t q = 0; q < ord.size(); ++q) {
            exint p = ord[q];
            if (q > 0 && ci[p] == idx.back() && (exint)idx.size() > nptr[r]) val.back() += cv[p];
            else { idx.push_back(ci[p]); val.push_back(cv[p]); }
        }
        nptr[r + 1] = (exint)idx.size();
    }
    ptr.swap(nptr);
}

Csr Csr::transpose() const {
    Csr T; T.rows = cols; T.cols = rows; T.ptr.assign(cols + 1, 0);
    for (int c : idx) T.ptr[c + 1]++;
    for (exint c = 0; c < cols; ++c) T.ptr[c + 1] += T.ptr[c];
    T.idx.resize(idx.size()); T.val.resize(val.size());
    std::vector<exint> cur(T.ptr.begin(), T.ptr.end() - 1);
    for (exint r = 0; r < rows; ++r)
        for (exint p = ptr[r]; p < ptr[r + 1]; ++p) { exint q = cur[idx[p]]++; T.idx[q] = (int)r; T.val[q] = val[p]; }
    return T;
}

// extern/eigen/Eigen/src/SparseCore/SparseDenseProduct.h:28-70 (row-major, OpenMP over rows)
void Csr::mulVec(const Real* x, Real* y) const {
#pragma omp parallel for schedule(static)
    for (exint r = 0; r < rows; ++r) {
        Real s = 0;
        for (exint p = ptr[r]; p < ptr[r + 1]; ++p) s += val[p] * x[idx[p]];
        y[r] = s;
    }
}
void Csr::mulVecAdd(const Real* x, Real* y, Real sc) const {
#pragma omp parallel for schedule(static)
    for (exint r = 0; r < rows; ++r) {
        Real s = 0;
        for (exint p = ptr[r]; p < ptr[r + 1]; ++p) s += val[p] * x[idx[p]];
        y[r] += sc * s;
    }
}

Csr spgemm(const Csr& A, const Csr& B) {
    assert(A.cols == B.rows);
    Csr C; C.rows = A.rows; C.cols = B.cols; C.ptr.assign(A.rows + 1, 0);
    std::vector<char> mask(B.cols, 0);
    std::vector<Real> acc(B.cols, 0.0);
    std::vector<int> touched;
    for (exint r = 0; r < A.rows; ++r) {
        touched.clear();
        for (exint p = A.ptr[r]; p < A.ptr[r + 1]; ++p) {
            const int k = A.idx[p]; const Real a = A.val[p];
            for (exint q = B.ptr[k]; q < B.ptr[k + 1]; ++q) {
                const int c = B.idx[q];
                if (!mask[c]) { mask[c] = 1; acc[c] = a * B.val[q]; touched.push_back(c); }
                else acc[c] += a * B.val[q];
            }
        }
        std::sort(touched.begin(), touched.end());
        for (int c : touched) { C.idx.push_back(c); C.val.push_back(acc[c]); mask[c] = 0; }
        C.ptr[r + 1] = (exint)C.idx.size();
    }
    return C;
}

Csr scaleRows(const std::vector<Real>& diag, const Csr& B) {
    Csr C = B;
    for (exint r = 0; r < C.rows; ++r)
        for (exint p = C.ptr[r]; p < C.ptr[r + 1]; ++p) C.val[p] = diag[r] * B.val[p];
    return C;
}

Csr addScaled(Real sa, const Csr& A, Real sb, const Csr& B) {
    assert(A.rows == B.rows && A.cols == B.cols);
    Csr C; C.rows = A.rows; C.cols = A.cols; C.ptr.assign(A.rows + 1, 0);
    for (exint r = 0; r < A.rows; ++r) {
        exint p = A.ptr[r], pe = A.ptr[r + 1], q = B.ptr[r], qe = B.ptr[r + 1];
        while (p < pe || q < qe) {
            if (q >= qe || (p < pe && A.idx[p] < B.idx[q])) { C.idx.push_back(A.idx[p]); C.val.push_back(sa * A.val[p]); ++p; }
            else if (p >= pe || B.idx[q] < A.idx[p]) { C.idx.push_back(B.idx[q]); C.val.push_back(sb * B.val[q]); ++q; }
            else { C.idx.push_back(A.idx[p]); C.val.push_back(sa * A.val[p] + sb * B.val[q]); ++p; ++q; }
        }
        C.ptr[r + 1] = (exint)C.idx.size();
    }
    return C;
}

// S_AB:209 `localReducedMatrix.inverse()`: for fixed sizes > 4 Eigen dispatches to
// PartialPivLU (extern/eigen/Eigen/src/LU/InverseImpl.h:25-31) and solves against the identity.
void inversePartialPivLU(const Real* Ain, Real* Ainv, int n) {
    std::vector<Real> lu(Ain, Ain + (size_t)n * n);
    std::vector<int> perm(n);
    std::iota(perm.begin(), perm.end(), 0);
    for (int k = 0; k < n; ++k) {
        int piv = k; Real best = std::fabs(lu[(size_t)k * n + k]);
        for (int i = k + 1; i < n; ++i) { Real a = std::fabs(lu[(size_t)i * n + k]); if (a > best) { best = a; piv = i; } }
        if (piv != k) { for (int j = 0; j < n; ++j) std::swap(lu[(size_t)k * n + j], lu[(size_t)piv * n + j]); std::swap(perm[k], perm[piv]); }
        const Real d = lu[(size_t)k * n + k];
        if (d != 0.0)
            for (int i = k + 1; i < n; ++i) lu[(size_t)i * n + k] /= d;
        for (int i = k + 1; i < n; ++i) {
            const Real l = lu[(size_t)i * n + k];
            for (int j = k + 1; j < n; ++j) lu[(size_t)i * n + j] -= l * lu[(size_t)k * n + j];
        }
    }
    std::vector<Real> col(n);
    for (int c = 0; c < n; ++c) {
        for (int i = 0; i < n; ++i) col[i] = (perm[i] == c) ? 1.0 : 0.0;
        for (int i = 0; i < n; ++i) { Real s = col[i]; for (int j = 0; j < i; ++j) s -= lu[(size_t)i * n + j] * col[j]; col[i] = s; }
        for (int i = n - 1; i >= 0; --i) { Real s = col[i]; for (int j = i + 1; j < n; ++j) s -= lu[(size_t)i * n + j] * col[j]; col[i] = s / lu[(size_t)i * n + i]; }
        for (int i = 0; i < n; ++i) Ainv[(size_t)i * n + c] = col[i];
    }
}

// S.cpp:415 `fullPivLu().solve(rhs)`: complete pivoting, rank decided by
// |pivot| > maxPivot * (epsilon * n) (Eigen FullPivLU default threshold), free variables zero.
void solveFullPivLU(const Real* Ain, const Real* rhs, Real* x, int n, int* rankOut) {
    std::vector<Real> lu(Ain, Ain + (size_t)n * n);
    std::vector<int> rowT(n), colT(n);
    int nonzeroPivots = n; Real maxPivot = 0;
    for (int k = 0; k < n; ++k) {
        int pr = k, pc = k; Real best = 0;
        for (int i = k; i < n; ++i) for (int j = k; j < n; ++j) { Real a = std::fabs(lu[(size_t)i * n + j]); if (a > best) { best = a; pr = i; pc = j; } }
        if (best == 0.0) { nonzeroPivots = k; for (int i = k; i < n; ++i) { rowT[i] = i; colT[i] = i; } break; }
        if (best > maxPivot) maxPivot = best;
        rowT[k] = pr; colT[k] = pc;
        if (pr != k) for (int j = 0; j < n; ++j) std::swap(lu[(size_t)k * n + j], lu[(size_t)pr * n + j]);
        if (pc != k) for (int i = 0; i < n; ++i) std::swap(lu[(size_t)i * n + k], lu[(size_t)i * n + pc]);
        const Real d = lu[(size_t)k * n + k];
        for (int i = k + 1; i < n; ++i) lu[(size_t)i * n + k] /= d;
        for (int i = k + 1; i < n; ++i) { const Real l = lu[(size_t)i * n + k]; for (int j = k + 1; j < n; ++j) lu[(size_t)i * n + j] -= l * lu[(size_t)k * n + j]; }
    }
    const Real thr = maxPivot * (std::numeric_limits<Real>::epsilon() * n);
    int rank = 0;
    for (int i = 0; i < nonzeroPivots; ++i) if (std::fabs(lu[(size_t)i * n + i]) > thr) ++rank;
    if (rankOut) *rankOut = rank;
    std::vector<Real> c(rhs, rhs + n);
    for (int k = 0; k < n; ++k) if (rowT[k] != k) std::swap(c[k], c[rowT[k]]);   // c = P rhs
    for (int i = 0; i < n; ++i) { Real s = c[i]; for (int j = 0; j < i; ++j) s -= lu[(size_t)i * n + j] * c[j]; c[i] = s; }  // unit-lower solve
    for (int i = rank - 1; i >= 0; --i) { Real s = c[i]; for (int j = i + 1; j < rank; ++j) s -= lu[(size_t)i * n + j] * c[j]; c[i] = s / lu[(size_t)i * n + i]; }
    for (int i = rank; i < n; ++i) c[i] = 0.0;
    for (int k = n - 1; k >= 0; --k) if (colT[k] != k) std::swap(c[k], c[colT[k]]);  // x = Q c
    for (int i = 0; i < n; ++i) x[i] = c[i];
}

bool saveMarket(const Csr& m, const std::string& path) {
    std::ofstream out(path.c_str(), std::ios::out);
    if (!out) return false;
    out.flags(std::ios_base::scientific);
    out.precision(std::numeric_limits<Real>::digits10 + 2);
    out << "%%MatrixMarket matrix coordinate  real general" << std::endl;
    out << m.rows << " " << m.cols << " " << m.nnz() << "\n";
    for (exint r = 0; r < m.rows; ++r)
        for (exint p = m.ptr[r]; p < m.ptr[r + 1]; ++p) out << (r + 1) << " " << (m.idx[p] + 1) << " " << m.val[p] << "\n";
    return true;
}

bool saveMarketVector(const std::vector<Real>& v, const std::string& path) {
    std::ofstream out(path.c_str(), std::ios::out);
    if (!out) return false;
    out.flags(std::ios_base::scientific);
    out.precision(std::numeric_limits<Real>::digits10 + 2);
    out << "%%MatrixMarket matrix array real general\n";
    out << v.size() << " " << 1 << "\n";
    for (Real x : v) out << x << "\n";
    return true;
}

}  // namespace orc
