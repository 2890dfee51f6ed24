// ps_pcg.cu -- the per-iteration hot path: factored operator apply + CG vector kernels
// (K13-K15).
//
// Reference: y = A x as a product of stored factors (lib/include/ApplyPressureStressMatrix.h:102-179)
//     y = -dt K^T Mc^-1 K x  -  J^T B^-1 J x  -  1/2 [0; mu^-1 x_tau],   K = [G D^T], J = [JG JD^T]
// and the CG loop lib/include/pcg.h:268-340.  Here J is never stored: J = C K_red (see ps_assemble.cu), so
//   pass 1 : w = K_ext x             (one thread per face row, 8-wide slot-major ELL, coalesced 8B/4B
//                                     streams, x gathered through L1/L2); active rows scaled by dt Mc^-1
//   moments: t_r = sum_f c_f w_f,  s_r = B_r^-1 t_r   (one CTA per chunk of a region's rows, then per region)
//   expand : w_f = c_f . s_r          on the reduced rows
//   pass 2 : y = -K_ext^T w - 1/2 mu^-1 x_tau         (one thread per DOF row, ELL widths 6/2/4), fused with
//            the p.Ap dot product (warp shuffle -> CTA partial -> last CTA finishes in fixed order)
// All CG scalars live in device memory; the host only polls a convergence flag every few iterations.
#include "ps_solver.hpp"

namespace ps {

// ---- row functors shared by the CUDA kernels and the emulation twins ----
PS_D double k_row(const OpArgs& A, int64_t r, const double* __restrict__ x) {
    double s = 0.;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += A.kval[(int64_t)k * A.nRowsExt + r] * x[A.kcol[(int64_t)k * A.nRowsExt + r]];
    return s;
}
PS_D double kt_row(const OpArgs& A, int64_t j, const double* __restrict__ w) {
    double s = 0.;
    if (j < A.nP) {
#pragma unroll
        for (int k = 0; k < 6; ++k) s += A.ktpVal[(int64_t)k * A.nP + j] * w[A.ktpCol[(int64_t)k * A.nP + j]];
    } else if (j < A.nP + 3 * A.nC) {
        const int64_t jj = j - A.nP, n = 3 * A.nC;
#pragma unroll
        for (int k = 0; k < 2; ++k) s += A.ktcVal[(int64_t)k * n + jj] * w[A.ktcCol[(int64_t)k * n + jj]];
    } else {
        const int64_t jj = j - A.nP - 3 * A.nC;
#pragma unroll
        for (int k = 0; k < 4; ++k) s += A.kteVal[(int64_t)k * A.nE + jj] * w[A.kteCol[(int64_t)k * A.nE + jj]];
    }
    return s;
}

// fixed-order finish of a dot product: called by the last CTA (or the emulation) over the CTA partials
PS_D double sum_partials(const double* p, int n) { double s = 0.; for (int i = 0; i < n; ++i) s += p[i]; return s; }

// CG bookkeeping after p.Ap is known (pcg.h:313)
PS_D void finish_pAp(PcgScalars* S, double pAp) { S->pAp = pAp; S->alpha = S->rsold / pAp; }
// ... and after r.r, x.x are known (pcg.h:316-336): stop test min(rr, rr/xx) < tol^2, else beta / rsold / iter
PS_D void finish_xr(PcgScalars* S, double rr, double xx) {
    S->rsnew = rr; S->xmag = xx;
    double rre = rr;
    if (rr / xx < rre) rre = rr / xx;
    S->rre = rre;
    if (rre < S->tol2) { S->done = 1; return; }
    S->beta = rr / S->rsold;
    S->rsold = rr;
    S->iter += 1;
    if (S->iter >= S->maxIter) S->done = 2;
}

#ifndef PS_EMULATE
constexpr int HOT_THREADS = 256;
constexpr int HOT_MAX_BLOCKS = 148 * 8;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
// CTA reduce; result valid in thread 0
__device__ __forceinline__ double block_sum(double v) {
    __shared__ double ws[HOT_THREADS / 32];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.;
    if (threadIdx.x < 32) { s = threadIdx.x < HOT_THREADS / 32 ? ws[threadIdx.x] : 0.; s = warp_sum(s); }
    __syncthreads();
    return s;
}
// fixed-order sum of the CTA partials by the last CTA (thread i takes partials i, i+256, ...)
__device__ __forceinline__ double block_sum_partials(const double* p, int n) {
    double s = 0.;
    for (int i = threadIdx.x; i < n; i += HOT_THREADS) s += __ldcg(p + i);
    return block_sum(s);
}
// returns true (in every thread) for the last CTA to arrive: all partials are visible to it
__device__ __forceinline__ bool last_block(unsigned int* ticket) {
    __shared__ bool last;
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned t = atomicInc(ticket, gridDim.x - 1);
        last = (t == gridDim.x - 1);
        if (last) __threadfence();
    }
    __syncthreads();
    return last;
}

__global__ void __launch_bounds__(HOT_THREADS) pass1_kernel(OpArgs A, const double* __restrict__ x, double* __restrict__ w, double activeScale, const PcgScalars* S) {
    if (S && S->done) return;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < A.nRowsExt; r += (int64_t)gridDim.x * blockDim.x) {
        const double s = k_row(A, r, x);
        w[r] = r < A.nActiveVs ? activeScale * A.mcInv[r] * s : s;
    }
}
// mode bit 0: accumulate dot(x, y) and finish p.Ap / alpha in the last CTA
__global__ void __launch_bounds__(HOT_THREADS) pass2_kernel(OpArgs A, const double* __restrict__ w, const double* __restrict__ x, double* __restrict__ y,
                                                           double muScale, const double* __restrict__ add, double* dotPartial, PcgScalars* S, int mode) {
    if (S && S->done) return;
    const int64_t n = A.nP + A.nT;
    double acc = 0.;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        double v = -kt_row(A, j, w);
        const double xj = (mode & 1) || j >= A.nP ? (x ? x[j] : 0.) : 0.;
        if (j >= A.nP && muScale != 0.) v -= muScale * A.uInv[j - A.nP] * xj;
        if (add) v += add[j];
        y[j] = v;
        acc += xj * v;
    }
    if (mode & 1) {
        const double bs = block_sum(acc);
        if (threadIdx.x == 0) dotPartial[blockIdx.x] = bs;
        if (last_block(&S->ticket[0])) { const double t = block_sum_partials(dotPartial, gridDim.x); if (threadIdx.x == 0) finish_pAp(S, t); }
    }
}
__global__ void __launch_bounds__(HOT_THREADS) cg_update_xr_kernel(int64_t n, double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p, const double* __restrict__ Ap,
                                                                  double* dotPartial, PcgScalars* S) {
    if (S->done) return;
    const double alpha = S->alpha;
    double rr = 0., xx = 0.;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double xi = x[i] + alpha * p[i], ri = r[i] - alpha * Ap[i];
        x[i] = xi; r[i] = ri;
        rr += ri * ri; xx += xi * xi;
    }
    const double brr = block_sum(rr), bxx = block_sum(xx);
    if (threadIdx.x == 0) { dotPartial[blockIdx.x] = brr; dotPartial[gridDim.x + blockIdx.x] = bxx; }
    if (last_block(&S->ticket[1])) {
        const double trr = block_sum_partials(dotPartial, gridDim.x), txx = block_sum_partials(dotPartial + gridDim.x, gridDim.x);
        if (threadIdx.x == 0) finish_xr(S, trr, txx);
    }
}
__global__ void __launch_bounds__(HOT_THREADS) cg_update_p_kernel(int64_t n, double* __restrict__ p, const double* __restrict__ r, const PcgScalars* S) {
    if (S->done) return;
    const double beta = S->beta;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = r[i] + beta * p[i];
}
__global__ void __launch_bounds__(HOT_THREADS) cg_init_kernel(int64_t n, const double* __restrict__ b, double* x, double* r, double* p, double* dotPartial, PcgScalars* S, double tol, int maxIter) {
    double rr = 0.;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double bi = b[i];
        x[i] = 0.; r[i] = bi; p[i] = bi; rr += bi * bi;
    }
    const double brr = block_sum(rr);
    if (threadIdx.x == 0) dotPartial[blockIdx.x] = brr;
    if (last_block(&S->ticket[2])) {
        const double rs = block_sum_partials(dotPartial, gridDim.x);
        if (threadIdx.x == 0) {
            S->rsold = rs; S->pAp = 0.; S->alpha = 0.; S->beta = 0.; S->rsnew = 0.; S->xmag = 0.; S->rre = 0.;
            S->iter = 0; S->done = (rs == 0.) ? 1 : 0; S->maxIter = maxIter; S->tol2 = tol * tol;
        }
    }
}
static inline int hot_blocks(int64_t n) { int64_t b = (n + HOT_THREADS - 1) / HOT_THREADS; if (b > HOT_MAX_BLOCKS) b = HOT_MAX_BLOCKS; if (b < 1) b = 1; return (int)b; }

void k_pass1(cudaStream_t st, const OpArgs& A, const double* x, double* w, double activeScale, const PcgScalars* S) {
    if (A.nRowsExt <= 0) return;
    pass1_kernel<<<hot_blocks(A.nRowsExt), HOT_THREADS, 0, st>>>(A, x, w, activeScale, S);
    PS_COUNT_LAUNCH(1);
    PS_CUDA(cudaGetLastError());
}
void k_pass2(cudaStream_t st, const OpArgs& A, const double* w, const double* x, double* y, double muScale, const double* add, double* dotPartial, int, PcgScalars* scal, int mode) {
    pass2_kernel<<<hot_blocks(A.nP + A.nT), HOT_THREADS, 0, st>>>(A, w, x, y, muScale, add, dotPartial, scal, mode);
    PS_COUNT_LAUNCH(1);
    PS_CUDA(cudaGetLastError());
}
void k_cg_update_xr(cudaStream_t st, int64_t n, double* x, double* r, const double* p, const double* Ap, double* dotPartial, int, PcgScalars* scal) {
    cg_update_xr_kernel<<<hot_blocks(n), HOT_THREADS, 0, st>>>(n, x, r, p, Ap, dotPartial, scal);
    PS_COUNT_LAUNCH(1);
    PS_CUDA(cudaGetLastError());
}
void k_cg_update_p(cudaStream_t st, int64_t n, double* p, const double* r, const PcgScalars* scal) {
    cg_update_p_kernel<<<hot_blocks(n), HOT_THREADS, 0, st>>>(n, p, r, scal);
    PS_COUNT_LAUNCH(1);
    PS_CUDA(cudaGetLastError());
}
void k_cg_init(cudaStream_t st, int64_t n, const double* b, double* x, double* r, double* p, double* dotPartial, int, PcgScalars* scal, double tol, int maxIter) {
    cg_init_kernel<<<hot_blocks(n), HOT_THREADS, 0, st>>>(n, b, x, r, p, dotPartial, scal, tol, maxIter);
    PS_COUNT_LAUNCH(1);
    PS_CUDA(cudaGetLastError());
}
#else  // ---- serial twins ----
void k_pass1(cudaStream_t, const OpArgs& A, const double* x, double* w, double activeScale, const PcgScalars* S) {
    if (S && S->done) return;
    for (int64_t r = 0; r < A.nRowsExt; ++r) { const double s = k_row(A, r, x); w[r] = r < A.nActiveVs ? activeScale * A.mcInv[r] * s : s; }
}
void k_pass2(cudaStream_t, const OpArgs& A, const double* w, const double* x, double* y, double muScale, const double* add, double*, int, PcgScalars* S, int mode) {
    if (S && S->done) return;
    const int64_t n = A.nP + A.nT;
    double acc = 0.;
    for (int64_t j = 0; j < n; ++j) {
        double v = -kt_row(A, j, w);
        const double xj = (mode & 1) || j >= A.nP ? (x ? x[j] : 0.) : 0.;
        if (j >= A.nP && muScale != 0.) v -= muScale * A.uInv[j - A.nP] * xj;
        if (add) v += add[j];
        y[j] = v; acc += xj * v;
    }
    if (mode & 1) finish_pAp(S, acc);
}
void k_cg_update_xr(cudaStream_t, int64_t n, double* x, double* r, const double* p, const double* Ap, double*, int, PcgScalars* S) {
    if (S->done) return;
    double rr = 0., xx = 0.;
    for (int64_t i = 0; i < n; ++i) { x[i] += S->alpha * p[i]; r[i] -= S->alpha * Ap[i]; rr += r[i] * r[i]; xx += x[i] * x[i]; }
    finish_xr(S, rr, xx);
}
void k_cg_update_p(cudaStream_t, int64_t n, double* p, const double* r, const PcgScalars* S) {
    if (S->done) return;
    for (int64_t i = 0; i < n; ++i) p[i] = r[i] + S->beta * p[i];
}
void k_cg_init(cudaStream_t, int64_t n, const double* b, double* x, double* r, double* p, double*, int, PcgScalars* S, double tol, int maxIter) {
    double rr = 0.;
    for (int64_t i = 0; i < n; ++i) { x[i] = 0.; r[i] = b[i]; p[i] = b[i]; rr += b[i] * b[i]; }
    S->rsold = rr; S->pAp = S->alpha = S->beta = S->rsnew = S->xmag = S->rre = 0.; S->iter = 0; S->done = rr == 0. ? 1 : 0; S->maxIter = maxIter; S->tol2 = tol * tol;
}
#endif

// ---- reduced regions: t_r = sum_f c_f w_f over the region's coupled rows; s_r = B_r^-1 (extraScale*extra_r + tScale*t_r) ----
#ifdef PS_EMULATE
PS_D void row_basis(const Geom& g, const RegionData& RG, const double* com, int64_t row, int region, double* c) {
    const int32_t packed = RG.rowFace.p[row];
    const int axis = (packed >> 29) & 3;
    const I3 f = delin(g, SL_FACE + axis, (int64_t)(packed & 0x1fffffff));
    double ox, oy, oz;
    face_offset(g, f, axis, com + 3 * region, ox, oy, oz);
    conversion_coefficients(ox, oy, oz, axis, c);
}
#endif

#ifndef PS_EMULATE
constexpr int MOM_THREADS = 128;
__global__ void __launch_bounds__(MOM_THREADS) moments_partial_kernel(Geom g, const int32_t* __restrict__ rowFace, const int32_t* __restrict__ chunk, const double* __restrict__ com,
                                                                     const double* __restrict__ wRows, double* __restrict__ partial) {
    __shared__ double red[MOM_THREADS / 32][RDOF];
    const int region = chunk[3 * blockIdx.x], begin = chunk[3 * blockIdx.x + 1], end = chunk[3 * blockIdx.x + 2];
    double acc[RDOF];
#pragma unroll
    for (int n = 0; n < RDOF; ++n) acc[n] = 0.;
    for (int row = begin + threadIdx.x; row < end; row += MOM_THREADS) {
        const int32_t packed = rowFace[row];
        const int axis = (packed >> 29) & 3;
        const I3 f = delin(g, SL_FACE + axis, (int64_t)(packed & 0x1fffffff));
        double ox, oy, oz, c[RDOF];
        face_offset(g, f, axis, com + 3 * region, ox, oy, oz);
        conversion_coefficients(ox, oy, oz, axis, c);
        const double wv = wRows[row];
#pragma unroll
        for (int n = 0; n < RDOF; ++n) acc[n] += c[n] * wv;
    }
#pragma unroll
    for (int n = 0; n < RDOF; ++n) {
        double v = acc[n];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][n] = v;
    }
    __syncthreads();
    if (threadIdx.x < RDOF) {
        double s = 0.;
        for (int wI = 0; wI < MOM_THREADS / 32; ++wI) s += red[wI][threadIdx.x];
        partial[(size_t)blockIdx.x * RDOF + threadIdx.x] = s;
    }
}
// one warp per region: sum chunk partials in order, then the 26x26 B^-1 GEMV
__global__ void __launch_bounds__(32) moments_finish_kernel(int R, const int32_t* __restrict__ chunkStart, const double* __restrict__ partial, const double* __restrict__ Binv,
                                                           const double* __restrict__ extra, double extraScale, double tScale, double* __restrict__ tOut, double* __restrict__ sOut) {
    __shared__ double t[RDOF];
    const int r = blockIdx.x;
    if (threadIdx.x < RDOF) {
        double s = 0.;
        for (int ch = chunkStart[r]; ch < chunkStart[r + 1]; ++ch) s += partial[(size_t)ch * RDOF + threadIdx.x];
        double v = tScale * s;
        if (extra) v += extraScale * extra[(size_t)r * RDOF + threadIdx.x];
        t[threadIdx.x] = v;
        tOut[(size_t)r * RDOF + threadIdx.x] = v;
    }
    __syncwarp();
    if (threadIdx.x < RDOF) {
        const double* B = Binv + (size_t)r * RDOF * RDOF + threadIdx.x * RDOF;
        double s = 0.;
#pragma unroll
        for (int j = 0; j < RDOF; ++j) s += B[j] * t[j];
        sOut[(size_t)r * RDOF + threadIdx.x] = s;
    }
}
void reduced_moments(cudaStream_t st, const Geom& g, const RegionData& RG, const double* wRows, int64_t, const double* extraRhs, double extraScale, double tScale) {
    if (RG.count <= 0) return;
    if (RG.nRowChunks > 0) moments_partial_kernel<<<RG.nRowChunks, MOM_THREADS, 0, st>>>(g, RG.rowFace.p, RG.rowChunk.p, RG.com.p, wRows, RG.partial.p);
    PS_COUNT_LAUNCH(1);
    moments_finish_kernel<<<RG.count, 32, 0, st>>>(RG.count, RG.rowChunkStart.p, RG.partial.p, RG.Binv.p, extraRhs, extraScale, tScale, RG.t.p, RG.s.p);
    PS_COUNT_LAUNCH(1);
    PS_CUDA(cudaGetLastError());
}
#else
void reduced_moments(cudaStream_t, const Geom& g, const RegionData& RG, const double* wRows, int64_t, const double* extraRhs, double extraScale, double tScale) {
    for (int r = 0; r < RG.count; ++r) {
        double t[RDOF];
        for (int n = 0; n < RDOF; ++n) t[n] = 0.;
        for (int ch = RG.rowChunkStart.p[r]; ch < RG.rowChunkStart.p[r + 1]; ++ch)
            for (int row = RG.rowChunk.p[3 * ch + 1]; row < RG.rowChunk.p[3 * ch + 2]; ++row) {
                double c[RDOF]; row_basis(g, RG, RG.com.p, row, r, c);
                for (int n = 0; n < RDOF; ++n) t[n] += c[n] * wRows[row];
            }
        for (int n = 0; n < RDOF; ++n) { t[n] = tScale * t[n] + (extraRhs ? extraScale * extraRhs[r * RDOF + n] : 0.); RG.t.p[r * RDOF + n] = t[n]; }
        for (int i = 0; i < RDOF; ++i) { double s = 0.; for (int j = 0; j < RDOF; ++j) s += RG.Binv.p[(size_t)r * RDOF * RDOF + i * RDOF + j] * t[j]; RG.s.p[r * RDOF + i] = s; }
    }
}
#endif

// w_f = scale * c_f . s_r on the coupled reduced rows (wRows points at row nActiveVs of w)
void k_reduced_expand(cudaStream_t st, const Geom& g, const RegionData& RG, double* wRows, int64_t, double scale) {
    if (RG.nRows <= 0) return;
    const int32_t* rowFace = RG.rowFace.p; const int32_t* rowRegion = RG.rowRegion.p; const double* com = RG.com.p; const double* s = RG.s.p;
    ps_for(st, RG.nRows, PS_LAMBDA(int64_t row) {
        const int region = rowRegion[row];
        const int32_t packed = rowFace[row];
        const int axis = (packed >> 29) & 3;
        const I3 f = delin(g, SL_FACE + axis, (int64_t)(packed & 0x1fffffff));
        double ox, oy, oz, c[RDOF];
        face_offset(g, f, axis, com + 3 * region, ox, oy, oz);
        conversion_coefficients(ox, oy, oz, axis, c);
        double v = 0.;
        for (int n = 0; n < RDOF; ++n) v += c[n] * s[(int64_t)region * RDOF + n];
        wRows[row] = scale * v;
    });
}

// W1 recoverVelocityFromPressureStress, active part (S.cpp:507): u = dt Mc^-1 (rhs_u/dt - G p - D^T tau)
// (wAct already holds dt Mc^-1 K x from pass 1)
void k_recover_active(cudaStream_t st, const Geom& g, int64_t nActiveVs, const double* wAct, const double* mcInv, const double* rhsU, double* velSol) {
    const double dt = g.dt, invDt = g.invDt;
    ps_for(st, nActiveVs, PS_LAMBDA(int64_t i) { velSol[i] = dt * (mcInv[i] * (invDt * rhsU[i])) - wAct[i]; });
}

// W2 applySolutionToVelocity (S.cpp:937-1028) fused with buildValidFaces (S_Cls:4-54)
void k_writeback_velocity(cudaStream_t st, const Geom& g, const Fields& F, const Counts& C, const RegionData& RG, const double* velSol, int axis, float* velOut, bool writeValid, float* validOut) {
    const int8_t* FL = F.label[SL_FACE + axis]; const int32_t* FA = F.aidx[SL_FACE + axis]; const int32_t* FR = F.ridx[SL_FACE + axis];
    const float* cvel = F.colvel[axis];
    const double* com = RG.com.p;
    const int64_t faceOff = C.faceOff[axis], nAct = C.nActiveVs;
    const bool haveReduced = RG.count > 0;
    ps_for(st, g.n[SL_FACE + axis], PS_LAMBDA(int64_t q) {
        const int lab = FL[q];
        const bool valid = !(lab == L_UNSOLVED || lab == L_UNASSIGNED);
        if (writeValid) validOut[q] = valid ? 1.f : 0.f;
        if (!valid || !velOut) return;
        double v = 0.;
        const int ri = haveReduced ? FR[q] : -1;
        const int ai = FA[q];
        if (ri >= 0) {
            const I3 f = delin(g, SL_FACE + axis, q);
            double ox, oy, oz, c[RDOF];
            face_offset(g, f, axis, com + 3 * ri, ox, oy, oz);
            conversion_coefficients(ox, oy, oz, axis, c);
            for (int n = 0; n < RDOF; ++n) v += velSol[nAct + (int64_t)RDOF * ri + n] * c[n];
        } else if (ai >= 0) v = velSol[faceOff + ai];
        else if (lab == L_SOLID) v = (double)cvel[q];
        velOut[q] = (float)v;
    });
}

}  // namespace ps
