// ps_oracle_solve.cpp -- TEST INFRASTRUCTURE ONLY (see ps_oracle.hpp header).
// Factored operator apply, the Krylov loops, velocity recovery and write-back, restated.
#include "ps_oracle.hpp"
#include <chrono>
#include <limits>
#include <algorithm>
#include <cstdio>

namespace orc {

static Real dot(const std::vector<Real>& a, const std::vector<Real>& b) {
    Real s = 0; const exint n = (exint)a.size();
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (exint i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

// Apply.h:24-68.  The reference recomputes McInv*G and McInv*Dt inside every apply
// (Apply.h:126,156); the products are loop invariants, so they are formed once here -- the
// per-apply arithmetic below is otherwise term-for-term that of Apply.h:102-179.
void Oracle::setupMatrixVectorProducts() {
    McInvG = scaleRows(mcInvDiag, G);
    McInvDt = scaleRows(mcInvDiag, Dt);
}

// Apply.h:102-179
void Oracle::applyMatrixVectorProducts(const Real* x, Real* y) {
    const exint np = nPressures, nt = nStresses, nu = nActiveVs, nr = nReducedVs;
    const Real* xp = x; const Real* xt = x + np;
    std::vector<Real> gx(nu), dxv(nu), A11_1(np), A21_1(nt), A12_1(np), A22_1(nt);
    // section 1 (Apply.h:124-134)
    McInvG.mulVec(xp, gx.data());
    Gt.mulVec(gx.data(), A11_1.data()); D.mulVec(gx.data(), A21_1.data());
    // section 3 (Apply.h:154-162)
    McInvDt.mulVec(xt, dxv.data());
    Gt.mulVec(dxv.data(), A12_1.data()); D.mulVec(dxv.data(), A22_1.data());
    // section 2 (Apply.h:136-152; util.h:203-230 is a transposed product)
    std::vector<Real> t1(nr), t2(nr), r1(nr), r2(nr), A11_2(np, 0.), A12_2(np, 0.), A21_2(nt, 0.), A22_2(nt, 0.);
    if (nr > 0) {
        JDt.mulVec(xt, t2.data()); BinvMat.mulVec(t2.data(), r2.data());
        JG.mulVec(xp, t1.data()); BinvMat.mulVec(t1.data(), r1.data());
        JGt.mulVec(r1.data(), A11_2.data()); JGt.mulVec(r2.data(), A12_2.data());
        DJt.mulVec(r1.data(), A21_2.data()); DJt.mulVec(r2.data(), A22_2.data());
    }
#pragma omp parallel for schedule(static)
    for (exint i = 0; i < np; ++i) {
        const Real A11 = -dt * A11_1[i] + (-A11_2[i]);
        const Real A12 = -dt * A12_1[i] + (-A12_2[i]);
        y[i] = A11 + A12;
    }
#pragma omp parallel for schedule(static)
    for (exint i = 0; i < nt; ++i) {
        const Real A21 = -dt * A21_1[i] + (-A21_2[i]);
        const Real A22_3 = -0.5 * uInvDiag[i] * xt[i];
        const Real A22 = -dt * A22_1[i] + (-A22_2[i]) + A22_3;
        y[np + i] = A21 + A22;
    }
}

// S.cpp:734-812 driving pcg.h:268-340 (pcg_external_matrix_A, identity preconditioner
// Preconditioner.cpp:271-274, zero start S.cpp:768) and the BiCGSTAB fallback pcg.h:134-200.
int Oracle::solveSPDwithMatrixVectorPCG() {
    auto t0 = std::chrono::steady_clock::now();
    setupMatrixVectorProducts();
    const exint n = nSystemSize;
    const Real tol = P.tolerance; const int maxIt = P.maxIterations;
    std::vector<Real>& x = solution;
    x.assign(n, 0.);
    std::vector<Real> r(n), z(n), p(n), Ap(n);
    usedBiCGStab = 0;
    auto pcg = [&]() -> int {
        applyMatrixVectorProducts(x.data(), Ap.data());
        for (exint i = 0; i < n; ++i) r[i] = b[i] - Ap[i];
        z = r; p = z;
        Real rsold = dot(r, z), rsnew = 0., alpha = 0., beta = 0., xmag = 0.;
        Real rre = 0.;
        for (int it = 0; it < maxIt; ++it) {
            applyMatrixVectorProducts(p.data(), Ap.data());
            alpha = rsold / dot(p, Ap);
#pragma omp parallel for schedule(static)
            for (exint i = 0; i < n; ++i) { x[i] = x[i] + alpha * p[i]; r[i] = r[i] - alpha * Ap[i]; }
            rsnew = dot(r, r);
            xmag = dot(x, x);
            rre = rsnew;
            if (rsnew / xmag < rre) rre = rsnew / xmag;
            if (rre < tol * tol) { solveError = std::sqrt(rre); return it; }
            z = r;
            rsnew = dot(r, z);
            beta = rsnew / rsold;
#pragma omp parallel for schedule(static)
            for (exint i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
            rsold = rsnew;
        }
        solveError = std::sqrt(rre);
        return maxIt;
    };
    auto bicgstab = [&]() -> int {
        applyMatrixVectorProducts(x.data(), Ap.data());
        for (exint i = 0; i < n; ++i) r[i] = b[i] - Ap[i];
        std::vector<Real> rhat = r, v(n, 0.), h(n, 0.), s(n, 0.), t(n, 0.), err(n, 0.);
        Real rhoCurr = 1., rhoOld = 1., alpha = 1., beta = 0., omega = 1., xmag = 0., rsnew = 0.;
        std::fill(p.begin(), p.end(), 0.);
        Real rre = 0.;
        for (int it = 0; it < maxIt; ++it) {
            rhoOld = rhoCurr; rhoCurr = dot(rhat, r);
            beta = (rhoCurr / rhoOld) * (alpha / omega);
            for (exint i = 0; i < n; ++i) p[i] = r[i] + beta * (p[i] - omega * v[i]);
            applyMatrixVectorProducts(p.data(), v.data());
            alpha = rhoCurr / dot(rhat, v);
            for (exint i = 0; i < n; ++i) { h[i] = x[i] + alpha * p[i]; s[i] = r[i] - alpha * v[i]; }
            applyMatrixVectorProducts(s.data(), t.data());
            omega = dot(t, s) / dot(t, t);
            for (exint i = 0; i < n; ++i) x[i] = h[i] + omega * s[i];
            xmag = std::sqrt(dot(x, x));
            applyMatrixVectorProducts(x.data(), Ap.data());
            for (exint i = 0; i < n; ++i) err[i] = b[i] - Ap[i];
            rsnew = dot(err, err);
            rre = rsnew;
            if (std::sqrt(rsnew) / xmag < rre) rre = std::sqrt(rsnew) / xmag;
            if (rre < tol) { solveError = rre; return it; }
            for (exint i = 0; i < n; ++i) r[i] = s[i] - omega * t[i];
        }
        solveError = rre;
        return maxIt;
    };
    solveIterations = pcg();
    if (solveIterations == maxIt) {
        usedBiCGStab = 1;
        std::fill(x.begin(), x.end(), 0.);
        solveIterations = bicgstab();
    }
    solveMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    solverResult = (solveIterations == maxIt) ? NOCONVERGE : SUCCESS;
    return solverResult;
}

// S.cpp:521-531: pressureGuess = -G^T u_old - JG^T v*;  stressGuess = -2 uInv (-D u_old - DJ^T v*)
// (the reference multiplies by uInv_Matrix here, not u_Matrix -- restated as written), then S_AS:413-419.
void Oracle::constructGuessVectors() {
    const exint np = nPressures, nt = nStresses, nr = nReducedVs;
    std::vector<Real> red(nr, 0.);
    for (exint r = 0; r < regionCount; ++r) for (int j = 0; j < RDOF; ++j) red[r * RDOF + j] = bestFit[r][j];
    std::vector<Real> g1(np, 0.), g2(np, 0.), d1(nt, 0.), d2(nt, 0.);
    Gt.mulVec(oldActiveVs.data(), g1.data()); D.mulVec(oldActiveVs.data(), d1.data());
    if (nr > 0) { JGt.mulVec(red.data(), g2.data()); DJt.mulVec(red.data(), d2.data()); }
    guess.assign(nSystemSize, 0.);
    for (exint i = 0; i < np; ++i) guess[i] = -g1[i] - g2[i];
    for (exint i = 0; i < nt; ++i) guess[np + i] = (-2. * uInvDiag[i]) * (-d1[i] - d2[i]);
}

// S.cpp:814-862.  Eigen::ConjugateGradient<SparseMatrix, Lower|Upper> with the default DiagonalPreconditioner:
// extern/eigen/Eigen/src/IterativeLinearSolvers/ConjugateGradient.h:28-93 (the loop restated statement by statement),
// BasicPreconditioners.h:66-94 (invdiag = 1/A_jj, 1 where the diagonal entry is absent or zero),
// solveWithGuess(b, guessVector).  Needs assembleExplicitA().
int Oracle::solveEigenCG() {
    auto t0 = std::chrono::steady_clock::now();
    const exint n = nSystemSize;
    std::vector<Real> invdiag(n, 1.);
    for (exint j = 0; j < n; ++j)
        for (exint q = A.ptr[j]; q < A.ptr[j + 1]; ++q)
            if (A.idx[q] == j) { if (A.val[q] != 0.) invdiag[j] = 1. / A.val[q]; break; }
    std::vector<Real>& x = solution;
    x = guess; x.resize(n, 0.);
    std::vector<Real> residual(n), p(n), z(n), tmp(n);
    const Real tol = P.tolerance; const int maxIters = P.maxIterations;
    int iters = maxIters; Real tol_error = tol;
    auto finish = [&]() {
        solveIterations = iters; solveError = tol_error;
        solveMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        solverResult = (tol_error <= tol) ? SUCCESS : NOCONVERGE;     // m_info, S.cpp:853-859
        return solverResult;
    };
    A.mulVec(x.data(), tmp.data());
    for (exint i = 0; i < n; ++i) residual[i] = b[i] - tmp[i];
    const Real rhsNorm2 = dot(b, b);
    if (rhsNorm2 == 0) { std::fill(x.begin(), x.end(), 0.); iters = 0; tol_error = 0; return finish(); }
    const Real considerAsZero = std::numeric_limits<Real>::min();
    const Real threshold = std::max(tol * tol * rhsNorm2, considerAsZero);
    Real residualNorm2 = dot(residual, residual);
    if (residualNorm2 < threshold) { iters = 0; tol_error = std::sqrt(residualNorm2 / rhsNorm2); return finish(); }
    for (exint i = 0; i < n; ++i) p[i] = invdiag[i] * residual[i];
    Real absNew = dot(residual, p);
    int i = 0;
    while (i < maxIters) {
        A.mulVec(p.data(), tmp.data());
        const Real alpha = absNew / dot(p, tmp);
#pragma omp parallel for schedule(static)
        for (exint k = 0; k < n; ++k) { x[k] += alpha * p[k]; residual[k] -= alpha * tmp[k]; }
        residualNorm2 = dot(residual, residual);
        if (residualNorm2 < threshold) break;
        for (exint k = 0; k < n; ++k) z[k] = invdiag[k] * residual[k];
        const Real absOld = absNew;
        absNew = dot(residual, z);
        const Real beta = absNew / absOld;
#pragma omp parallel for schedule(static)
        for (exint k = 0; k < n; ++k) p[k] = z[k] + beta * p[k];
        i++;
    }
    tol_error = std::sqrt(residualNorm2 / rhsNorm2);
    iters = i;
    return finish();
}

// S.cpp:492-510
void Oracle::recoverVelocityFromPressureStress() {
    const exint np = nPressures, nu = nActiveVs, nr = nReducedVs;
    const Real* ps = solution.data(); const Real* ts = solution.data() + np;
    velSolution.assign(nu + nr, 0.);
    std::vector<Real> gp(nu), dtau(nu);
    G.mulVec(ps, gp.data()); Dt.mulVec(ts, dtau.data());
    for (exint i = 0; i < nu; ++i) velSolution[i] = dt * (mcInvDiag[i] * (invDt * activeRHS[i] - gp[i] - dtau[i]));
    if (nr > 0) {
        std::vector<Real> jp(nr), jt(nr), w(nr), out(nr);
        JG.mulVec(ps, jp.data()); JDt.mulVec(ts, jt.data());
        for (exint i = 0; i < nr; ++i) w[i] = invDt * reducedRHS[i] - jp[i] - jt[i];
        BinvMat.mulVec(w.data(), out.data());
        for (exint i = 0; i < nr; ++i) velSolution[nu + i] = out[i];
    }
}

// S.cpp:937-1028
void Oracle::applySolutionToVelocity(float* velOut, const float* valid, int axis) {
    const Field<exint>& FL = labels[S_FACEX + axis];
    const Field<exint>& FA = activeIdx[S_FACEX + axis];
    const Field<exint>& FR = reducedIdx[S_FACEX + axis];
    const exint reducedVsOffset = nActiveVs;
    for (int k = 0; k < FL.r[2]; ++k) for (int j = 0; j < FL.r[1]; ++j) for (int i = 0; i < FL.r[0]; ++i) {
        const size_t q = FL.lin(i, j, k);
        if (valid[q] == 0.f) continue;
        const exint faceLabel = FL.d[q];
        const exint localActiveFaceIndex = FA.d[q];
        const exint reducedFaceIndex = FR.d[q];
        Real localVelocity = 0.;
        if (reducedFaceIndex >= 0) {
            Real off[3] = {(Real)i, (Real)j, (Real)k};
            off[axis] -= 0.5;
            for (int a = 0; a < 3; ++a) { off[a] *= dx; off[a] -= com[reducedFaceIndex][a]; }
            Real Cx[RDOF]; buildConversionCoefficients(off, axis, Cx);
            Real s = 0;
            for (int n = 0; n < RDOF; ++n) s += velSolution[reducedVsOffset + RDOF * reducedFaceIndex + n] * Cx[n];
            localVelocity = s;
        } else if (localActiveFaceIndex >= 0) {
            localVelocity = velSolution[faceVelocityDOF(localActiveFaceIndex, axis)];
        } else if (faceLabel == SOLID) {
            localVelocity = (Real)colVel[axis].d[q];
        }
        velOut[q] = (float)localVelocity;
    }
}

}  // namespace orc
