// ps_reduced.cu -- reduced-region dense algebra.
//
// Per region r the reference accumulates three 26x26 matrices and one 26-vector from rank-1 updates
// over the region's faces (exec/HDK_PolyStokesSolver.cpp:1330-1399 least squares, 1405-1482 mass,
// 1484-1694 viscosity), then solves/inverts them with Eigen (S.cpp:415, S_AB:195-244).  GPU layout:
//   * REDUCED cells are listed sorted by (region, voxel order); the list is cut into fixed-size chunks
//     that never straddle a region, one CTA per chunk;
//   * a CTA stages, per batch of faces, the basis rows c_f, the viscosity rows d_f and the weights in
//     shared memory and forms the partial Grams  sum c c^T (x2)  and  sum c d^T  (Phi^T W Phi);
//   * partials are summed per region in chunk order (deterministic), then one thread per region does
//     the 26x26 complete-pivoting solve and the partial-pivoting inverse.
#include "ps_solver.hpp"

namespace ps {

// D2 computeCenterOfMasses (S.cpp:328-372, 1274-1324): integer coordinate sums are exact, so the
// atomics below give the reference's value bit for bit: COM = sum * (dx / count).
void k_region_com(cudaStream_t st, const Geom& g, const Fields& F, int32_t R, unsigned long long* sums, double* com) {
    const int8_t* L = F.label[SL_CENTER]; const int32_t* Rg = F.ridx[SL_CENTER];
    dev_memset(sums, 0, (size_t)R * 4 * sizeof(unsigned long long), st);
    ps_for(st, g.n[SL_CENTER], PS_LAMBDA(int64_t q) {
        if (L[q] != L_REDUCED) return;
        const int r = Rg[q];
        const I3 c = delin(g, SL_CENTER, q);
        atomic_add(&sums[4 * r + 0], (unsigned long long)c.x); atomic_add(&sums[4 * r + 1], (unsigned long long)c.y);
        atomic_add(&sums[4 * r + 2], (unsigned long long)c.z); atomic_add(&sums[4 * r + 3], 1ull);
    });
    const double dx = g.dx;
    ps_for(st, R, PS_LAMBDA(int64_t r) {
        const double s = dx / (double)sums[4 * r + 3];
        for (int a = 0; a < 3; ++a) com[3 * r + a] = mul_rn((double)sums[4 * r + a], s);
    });
}

// scatter (region, voxel) pairs of flagged voxels to their voxel-order rank (input of the stable sort)
void k_collect_region_keys(cudaStream_t st, const Geom& g, const int32_t* rank, const uint8_t* flag, const int32_t* region, int64_t n, int32_t tag, int32_t rankOffset, int32_t* keys, int32_t* vals) {
    ps_for(st, n, PS_LAMBDA(int64_t q) {
        if (!flag[q]) return;
        const int64_t pos = (int64_t)rankOffset + rank[q];
        keys[pos] = region[q]; vals[pos] = (int32_t)q | tag;
    });
}

// SIM_RawField::getValue(indexToPos(sample)) shim: trilinear viscosity at a sample, fractions 0 or 1/2,
// lerp x, y, z as a + t*(b-a) in fp32
PS_D float local_viscosity(const Geom& g, const float* visc, int slot, const I3& idx) {
    int o[3] = {1, 1, 1};
    if (slot >= SL_FACE && slot < SL_EDGE) o[slot - SL_FACE] = 0;
    else if (slot >= SL_EDGE) { for (int a = 0; a < 3; ++a) if (a != slot - SL_EDGE) o[a] = 0; }
    const int b0 = o[0] ? idx.x : idx.x - 1, b1 = o[1] ? idx.y : idx.y - 1, b2 = o[2] ? idx.z : idx.z - 1;
    const float t0 = o[0] ? 0.f : 0.5f, t1 = o[1] ? 0.f : 0.5f, t2 = o[2] ? 0.f : 0.5f;
    float cz[2];
    for (int dz = 0; dz < 2; ++dz) {
        float cy[2];
        for (int dy = 0; dy < 2; ++dy) {
            const float a = float_at(g, visc, SL_CENTER, I3{b0, b1 + dy, b2 + dz});
            const float b = float_at(g, visc, SL_CENTER, I3{b0 + 1, b1 + dy, b2 + dz});
            cy[dy] = fadd_rn(a, fmul_rn(t0, fsub_rn(b, a)));
        }
        cz[dz] = fadd_rn(cy[0], fmul_rn(t1, fsub_rn(cy[1], cy[0])));
    }
    return fadd_rn(cz[0], fmul_rn(t2, fsub_rn(cz[1], cz[0])));
}

struct FaceTerms { double wM, wN, u; bool visc; };

// Everything the three per-region sums need from one (reduced cell, axis, dir) face:
//   c   = basis row of the face              (buildConversionCoefficients, S.cpp:2107-2149)
//   wM  = rho if the face is counted by the mass matrix   (S.cpp:1442-1472)
//   wN  = 1  if the opposite cell isActive (least-squares surface face, S.cpp:1369-1390), u = u*_face
//   d   = sum of contribution * basis row over the viscous stencil partners of the face
//         (centre terms S.cpp:1537-1603, edge terms S.cpp:1605-1683); valid when .visc is set
PS_D FaceTerms face_terms(const Geom& g, const Fields& F, const double* com, const I3& cell, int axis, int dir, int region, double* c, double* d) {
    FaceTerms t;
    const I3 face = dir ? shifted(cell, axis, 1) : cell;
    const I3 nbr = shifted(cell, axis, dir ? 1 : -1);
    const int nbrLabel = label_at(g, F.label[SL_CENTER], SL_CENTER, nbr);
    const bool nbrActive = is_active(nbrLabel);
    double ox, oy, oz;
    face_offset(g, face, axis, com + 3 * region, ox, oy, oz);
    conversion_coefficients(ox, oy, oz, axis, c);
    t.wM = (dir == 0 || nbrActive) ? g.density : 0.;
    t.wN = nbrActive ? 1. : 0.;
    t.u = (double)F.vel[axis][lin(g, SL_FACE + axis, face)];
    // the face belongs to this cell's region (and is visited exactly once) iff it is the cell's low
    // face, or its high face with a non-REDUCED cell behind it (findFaceReducedIndexFromCenter, S_Cls:1498-1528)
    t.visc = (dir == 0) || (nbrLabel != L_REDUCED);
    for (int n = 0; n < RDOF; ++n) d[n] = 0.;
    if (!t.visc) return t;
    const double dx2 = g.dx * g.dx;
    double row[RDOF];
    // cell-centred stress terms
    for (int divDir = 0; divDir < 2; ++divDir) {
        const I3 cc = divDir ? face : shifted(face, axis, -1);
        if (!is_reduced(label_at(g, F.label[SL_CENTER], SL_CENTER, cc))) continue;
        const double divSign = divDir ? 1. : -1.;
        const double visc = (double)F.viscosity[lin(g, SL_CENTER, cc)];
        for (int gradDir = 0; gradDir < 2; ++gradDir) {
            const I3 adjFace = gradDir ? shifted(cc, axis, 1) : cc;
            const double gradSign = gradDir ? 1. : -1.;
            const double contribution = -1. * divSign * gradSign * visc / dx2;
            const int adjR = F.ridx[SL_FACE + axis][lin(g, SL_FACE + axis, adjFace)];
            if (adjR < 0) continue;
            face_offset(g, adjFace, axis, com + 3 * adjR, ox, oy, oz);
            conversion_coefficients(ox, oy, oz, axis, row);
            for (int n = 0; n < RDOF; ++n) d[n] += contribution * row[n];
        }
    }
    // edge-centred stress terms
    for (int e = 0; e < 3; ++e) {
        if (e == axis) continue;
        for (int divDir = 0; divDir < 2; ++divDir) {
            const double divSign = divDir ? 1. : -1.;
            const I3 edge = divDir ? shifted(face, 3 - axis - e, 1) : face;
            if (F.label[SL_EDGE + e][lin(g, SL_EDGE + e, edge)] != L_REDUCED) continue;   // isReducedButNotBoundary
            const float visc = local_viscosity(g, F.viscosity, SL_EDGE + e, edge);
            for (int gradAxis = 0; gradAxis < 3; ++gradAxis) {
                if (gradAxis == e) continue;
                const int adjAxis = 3 - gradAxis - e;
                for (int gradDir = 0; gradDir < 2; ++gradDir) {
                    const I3 adjFace = gradDir ? edge : shifted(edge, gradAxis, -1);   // edgeToFaceMap: -1 on axis 3-adjAxis-e = gradAxis
                    const double gradSign = gradDir ? 1. : -1.;
                    const double contribution = -0.5 * divSign * gradSign * (double)visc / dx2;
                    const int adjR = index_at(g, F.ridx[SL_FACE + adjAxis], SL_FACE + adjAxis, adjFace);
                    if (adjR < 0) continue;
                    face_offset(g, adjFace, adjAxis, com + 3 * adjR, ox, oy, oz);
                    conversion_coefficients(ox, oy, oz, adjAxis, row);
                    for (int n = 0; n < RDOF; ++n) d[n] += contribution * row[n];
                }
            }
        }
    }
    return t;
}

// partial layout per chunk: [M 676][N 676][V 676][rhs 26]
constexpr int GRAM_STRIDE = 3 * RDOF * RDOF + RDOF;

#ifndef PS_EMULATE
constexpr int GRAM_THREADS = 256;
constexpr int GRAM_BATCH = 128;   // faces staged per pass (128 * (26+26+3) doubles = 56 KB smem)

__global__ void __launch_bounds__(GRAM_THREADS) gram_partial_kernel(Geom g, Fields F, const double* __restrict__ com, const int32_t* __restrict__ cellList,
                                                                   const int32_t* __restrict__ chunk, double* __restrict__ partial, int chunk0) {
    extern __shared__ double sm[];
    double* sc = sm;                              // [GRAM_BATCH][26]
    double* sd = sc + GRAM_BATCH * RDOF;          // [GRAM_BATCH][26]
    double* sw = sd + GRAM_BATCH * RDOF;          // [GRAM_BATCH][3]  wM, wN, u*wN
    const int ch = chunk0 + blockIdx.x;
    const int region = chunk[3 * ch + 0], begin = chunk[3 * ch + 1], end = chunk[3 * ch + 2];
    const int nFaces = (end - begin) * 6;
    // each thread owns up to 3 of the 676 (i,j) entries
    double accM[3] = {0, 0, 0}, accN[3] = {0, 0, 0}, accV[3] = {0, 0, 0};
    double accR = 0.;   // threads 0..25: rhs entry
    for (int base = 0; base < nFaces; base += GRAM_BATCH) {
        const int nb = min(GRAM_BATCH, nFaces - base);
        __syncthreads();
        if (threadIdx.x < nb) {
            const int item = base + threadIdx.x;
            const int cellQ = cellList[begin + item / 6];
            const int fa = (item % 6) >> 1, dir = item & 1;
            double c[RDOF], d[RDOF];
            const FaceTerms t = face_terms(g, F, com, delin(g, SL_CENTER, cellQ), fa, dir, region, c, d);
            for (int n = 0; n < RDOF; ++n) { sc[threadIdx.x * RDOF + n] = c[n]; sd[threadIdx.x * RDOF + n] = t.visc ? d[n] : 0.; }
            sw[threadIdx.x * 3 + 0] = t.wM; sw[threadIdx.x * 3 + 1] = t.wN; sw[threadIdx.x * 3 + 2] = t.wN * t.u;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int e = threadIdx.x + k * GRAM_THREADS;
            if (e < RDOF * RDOF) {
                const int i = e / RDOF, j = e % RDOF;
                double m = accM[k], nn = accN[k], v = accV[k];
                for (int f = 0; f < nb; ++f) {
                    const double ci = sc[f * RDOF + i], cj = sc[f * RDOF + j];
                    m += (sw[f * 3 + 0] * ci) * cj;
                    nn += (sw[f * 3 + 1] * ci) * cj;
                    v += ci * sd[f * RDOF + j];
                }
                accM[k] = m; accN[k] = nn; accV[k] = v;
            }
        }
        if (threadIdx.x < RDOF) {
            double rr = accR;
            for (int f = 0; f < nb; ++f) rr += sw[f * 3 + 2] * sc[f * RDOF + threadIdx.x];
            accR = rr;
        }
    }
    double* out = partial + (size_t)ch * GRAM_STRIDE;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int e = threadIdx.x + k * GRAM_THREADS;
        if (e < RDOF * RDOF) { out[e] = accM[k]; out[RDOF * RDOF + e] = accN[k]; out[2 * RDOF * RDOF + e] = accV[k]; }
    }
    if (threadIdx.x < RDOF) out[3 * RDOF * RDOF + threadIdx.x] = accR;
}

void region_gram_partials(cudaStream_t st, const Geom& g, const Fields& F, const RegionData& RG, double* partial) {
    if (RG.cellChunkHi <= RG.cellChunkLo) return;
    const size_t smem = (size_t)GRAM_BATCH * (2 * RDOF + 3) * sizeof(double);
    static bool attr = false;
    if (!attr) { PS_CUDA(cudaFuncSetAttribute(gram_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
    gram_partial_kernel<<<RG.cellChunkHi - RG.cellChunkLo, GRAM_THREADS, smem, st>>>(g, F, RG.com.p, RG.cellList.p, RG.cellChunk.p, partial, RG.cellChunkLo);
    PS_COUNT_LAUNCH(1);
    PS_CUDA(cudaGetLastError());
}
#else
void region_gram_partials(cudaStream_t, const Geom& g, const Fields& F, const RegionData& RG, double* partial) {
    for (int ch = RG.cellChunkLo; ch < RG.cellChunkHi; ++ch) {
        const int region = RG.cellChunk.p[3 * ch], begin = RG.cellChunk.p[3 * ch + 1], end = RG.cellChunk.p[3 * ch + 2];
        double* out = partial + (size_t)ch * GRAM_STRIDE;
        for (int e = 0; e < GRAM_STRIDE; ++e) out[e] = 0.;
        for (int ci = begin; ci < end; ++ci)
            for (int item = 0; item < 6; ++item) {
                double c[RDOF], d[RDOF];
                const FaceTerms t = face_terms(g, F, RG.com.p, delin(g, SL_CENTER, RG.cellList.p[ci]), item >> 1, item & 1, region, c, d);
                for (int i = 0; i < RDOF; ++i) {
                    for (int j = 0; j < RDOF; ++j) {
                        out[i * RDOF + j] += (t.wM * c[i]) * c[j];
                        out[RDOF * RDOF + i * RDOF + j] += (t.wN * c[i]) * c[j];
                        if (t.visc) out[2 * RDOF * RDOF + i * RDOF + j] += c[i] * d[j];
                    }
                    out[3 * RDOF * RDOF + i] += (t.wN * t.u) * c[i];
                }
            }
    }
}
#endif

// dense 26x26 routines, one thread per region, matrices in a private global scratch slab.
// S_AB:209 .inverse() -> PartialPivLU (extern/eigen/Eigen/src/LU/InverseImpl.h:25-31)
PS_D void inverse_partial_piv(double* lu, double* inv, int* perm) {
    const int n = RDOF;
    for (int i = 0; i < n; ++i) perm[i] = i;
    for (int k = 0; k < n; ++k) {
        int piv = k; double best = fabs(lu[k * n + k]);
        for (int i = k + 1; i < n; ++i) { const double a = fabs(lu[i * n + k]); if (a > best) { best = a; piv = i; } }
        if (piv != k) { for (int j = 0; j < n; ++j) { const double t = lu[k * n + j]; lu[k * n + j] = lu[piv * n + j]; lu[piv * n + j] = t; } const int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t; }
        const double dg = lu[k * n + k];
        if (dg != 0.) for (int i = k + 1; i < n; ++i) lu[i * n + k] /= dg;
        for (int i = k + 1; i < n; ++i) { const double l = lu[i * n + k]; for (int j = k + 1; j < n; ++j) lu[i * n + j] -= l * lu[k * n + j]; }
    }
    double col[RDOF];
    for (int c = 0; c < n; ++c) {
        for (int i = 0; i < n; ++i) col[i] = (perm[i] == c) ? 1. : 0.;
        for (int i = 0; i < n; ++i) { double s = col[i]; for (int j = 0; j < i; ++j) s -= lu[i * n + j] * col[j]; col[i] = s; }
        for (int i = n - 1; i >= 0; --i) { double s = col[i]; for (int j = i + 1; j < n; ++j) s -= lu[i * n + j] * col[j]; col[i] = s / lu[i * n + i]; }
        for (int i = 0; i < n; ++i) inv[i * n + c] = col[i];
    }
}
// S.cpp:415 fullPivLu().solve(): complete pivoting, rank threshold maxPivot * eps * n, free variables 0
PS_D void solve_full_piv(double* lu, const double* rhs, double* x, int* rowT, int* colT) {
    const int n = RDOF;
    int nonzero = n; double maxPivot = 0.;
    for (int k = 0; k < n; ++k) {
        int pr = k, pc = k; double best = 0.;
        for (int i = k; i < n; ++i) for (int j = k; j < n; ++j) { const double a = fabs(lu[i * n + j]); if (a > best) { best = a; pr = i; pc = j; } }
        if (best == 0.) { nonzero = k; for (int i = k; i < n; ++i) { rowT[i] = i; colT[i] = i; } break; }
        if (best > maxPivot) maxPivot = best;
        rowT[k] = pr; colT[k] = pc;
        if (pr != k) for (int j = 0; j < n; ++j) { const double t = lu[k * n + j]; lu[k * n + j] = lu[pr * n + j]; lu[pr * n + j] = t; }
        if (pc != k) for (int i = 0; i < n; ++i) { const double t = lu[i * n + k]; lu[i * n + k] = lu[i * n + pc]; lu[i * n + pc] = t; }
        const double dg = lu[k * n + k];
        for (int i = k + 1; i < n; ++i) lu[i * n + k] /= dg;
        for (int i = k + 1; i < n; ++i) { const double l = lu[i * n + k]; for (int j = k + 1; j < n; ++j) lu[i * n + j] -= l * lu[k * n + j]; }
    }
    const double thr = maxPivot * (2.220446049250313e-16 * n);
    int rank = 0;
    for (int i = 0; i < nonzero; ++i) if (fabs(lu[i * n + i]) > thr) ++rank;
    double c[RDOF];
    for (int i = 0; i < n; ++i) c[i] = rhs[i];
    for (int k = 0; k < n; ++k) if (rowT[k] != k) { const double t = c[k]; c[k] = c[rowT[k]]; c[rowT[k]] = t; }
    for (int i = 0; i < n; ++i) { double s = c[i]; for (int j = 0; j < i; ++j) s -= lu[i * n + j] * c[j]; c[i] = s; }
    for (int i = rank - 1; i >= 0; --i) { double s = c[i]; for (int j = i + 1; j < rank; ++j) s -= lu[i * n + j] * c[j]; c[i] = s / lu[i * n + i]; }
    for (int i = rank; i < n; ++i) c[i] = 0.;
    for (int k = n - 1; k >= 0; --k) if (colT[k] != k) { const double t = c[k]; c[k] = c[colT[k]]; c[colT[k]] = t; }
    for (int i = 0; i < n; ++i) x[i] = c[i];
}

// sums the chunk partials of every region in chunk order, then per region:
//   v* = fullPivLu(N).solve(rhs)                        (computeLeastSquaresFits, S.cpp:412-416)
//   B  = M/dt + 2 V ; B^-1 = B.inverse()                (assembleReducedInvertedBlock, S_AB:195-244)
//   rhs_r = M v*                                        (assembleReducedRHSVector, S_AB:356-367)
void region_gram_finish(cudaStream_t st, const Geom& g, RegionData& RG, int nChunks) {
    // only the regions this rank owns [regLo, regHi) (all of them on one GPU)
    const int R0 = RG.regLo, R = RG.regHi - RG.regLo;
    if (R <= 0) return;
    const double* partial = RG.partial.p; const int32_t* chunkStart = RG.cellChunkStart.p;
    double* Mr = RG.Mr.p; double* Vi = RG.Visc.p; double* Nm = RG.N.p; double* Binv = RG.Binv.p;
    double* lsq = RG.lsqRhs.p; double* fit = RG.bestFit.p; double* rhsR = RG.rhsR.p;
    const int NN = RDOF * RDOF;
    ps_for(st, (int64_t)R * NN, PS_LAMBDA(int64_t ql) {
        const int64_t q = ql + (int64_t)R0 * NN;
        const int r = (int)(q / NN), e = (int)(q % NN);
        double m = 0., n = 0., v = 0.;
        for (int ch = chunkStart[r]; ch < chunkStart[r + 1]; ++ch) {
            const double* p = partial + (size_t)ch * GRAM_STRIDE;
            m += p[e]; n += p[NN + e]; v += p[2 * NN + e];
        }
        Mr[q] = m; Nm[q] = n; Vi[q] = v;
    });
    ps_for(st, (int64_t)R * RDOF, PS_LAMBDA(int64_t ql) {
        const int64_t q = ql + (int64_t)R0 * RDOF;
        const int r = (int)(q / RDOF), e = (int)(q % RDOF);
        double s = 0.;
        for (int ch = chunkStart[r]; ch < chunkStart[r + 1]; ++ch) s += partial[(size_t)ch * GRAM_STRIDE + 3 * NN + e];
        lsq[q] = s;
    });
    // scratch: reuse the partial buffer's head is unsafe (still read above on the same stream is fine, but keep it simple)
    static thread_local DBuf<double> luA, luB;
    static thread_local DBuf<int> piv;
    luA.alloc((size_t)R * NN); luB.alloc((size_t)R * NN); piv.alloc((size_t)R * 3 * RDOF);
    double* A = luA.p; double* B = luB.p; int* pv = piv.p;
    const double invDt = g.invDt;
    ps_for(st, R, PS_LAMBDA(int64_t rl) {
        const int64_t r = rl + R0;
        double* a = A + rl * NN; double* b = B + rl * NN; int* p3 = pv + rl * 3 * RDOF;
        for (int e = 0; e < NN; ++e) { a[e] = Nm[r * NN + e]; b[e] = invDt * Mr[r * NN + e] + 2. * Vi[r * NN + e]; }
        solve_full_piv(a, lsq + r * RDOF, fit + r * RDOF, p3, p3 + RDOF);
        inverse_partial_piv(b, Binv + r * NN, p3 + 2 * RDOF);
        for (int i = 0; i < RDOF; ++i) { double s = 0.; for (int j = 0; j < RDOF; ++j) s += Mr[r * NN + i * RDOF + j] * fit[r * RDOF + j]; rhsR[r * RDOF + i] = s; }
    });
    (void)nChunks;
}

// a REDUCED face is a row of K_ext iff it has at least one entry of G / D^T (S_CMB:393-639):
// an adjacent cell with a pressure index and positive coefficient, or an adjacent isActive edge
void k_flag_coupled_faces(cudaStream_t st, const Geom& g, const Fields& F, int axis, uint8_t* flag) {
    const int8_t* FL = F.label[SL_FACE + axis]; const uint8_t* ffw = F.fluW[SL_FACE + axis];
    const int32_t* CA = F.aidx[SL_CENTER]; const uint8_t* clw = F.liqW[SL_CENTER];
    const int e1 = axis == 0 ? 1 : 0, e2 = axis == 2 ? 1 : 2;
    const int8_t* EL1 = F.label[SL_EDGE + e1]; const int8_t* EL2 = F.label[SL_EDGE + e2];
    const uint8_t* ew1 = F.liqW[SL_EDGE + e1]; const uint8_t* ew2 = F.liqW[SL_EDGE + e2];
    ps_for(st, g.n[SL_FACE + axis], PS_LAMBDA(int64_t q) {
        uint8_t f = 0;
        if (FL[q] == L_REDUCED && ffw[q] > 0) {
            const I3 fc = delin(g, SL_FACE + axis, q);
            for (int dir = 0; dir < 2; ++dir) {
                const I3 cell = dir ? fc : shifted(fc, axis, -1);
                if (!in_bounds(g, SL_CENTER, cell)) continue;
                const int64_t qc = lin(g, SL_CENTER, cell);
                if (CA[qc] >= 0 && clw[qc] > 0) f = 1;
            }
            for (int dir = 0; dir < 2; ++dir) {
                const I3 ed1 = dir ? shifted(fc, 3 - axis - e1, 1) : fc;
                const int64_t q1 = lin(g, SL_EDGE + e1, ed1);
                if (is_active(EL1[q1]) && ew1[q1] > 0) f = 1;
                const I3 ed2 = dir ? shifted(fc, 3 - axis - e2, 1) : fc;
                const int64_t q2 = lin(g, SL_EDGE + e2, ed2);
                if (is_active(EL2[q2]) && ew2[q2] > 0) f = 1;
            }
        }
        flag[q] = f;
    });
}

}  // namespace ps
