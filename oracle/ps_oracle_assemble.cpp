// ps_oracle_assemble.cpp -- TEST INFRASTRUCTURE ONLY (see ps_oracle.hpp header).
// Reduced-region dense algebra, ConstructMatrixBlocks, AssembleBlocks/AssembleSystem, restated.
#include "ps_oracle.hpp"
#include <chrono>

namespace orc {

// S.cpp:2107-2149 (QUADRATIC_REGIONS)
void Oracle::buildConversionCoefficients(const Real o[3], int axis, Real v[RDOF]) const {
    for (int n = 0; n < RDOF; ++n) v[n] = 0.;
    switch (axis) {
    case 0:
        v[0] = 1.;
        v[3] = o[0]; v[4] = o[1]; v[5] = o[2];
        v[6] = o[0] * o[0]; v[7] = o[0] * o[1]; v[8] = o[0] * o[2];
        v[9] = o[1] * o[1]; v[10] = o[1] * o[2]; v[11] = o[2] * o[2];
        break;
    case 1:
        v[1] = 1.;
        v[12] = o[0]; v[13] = o[1]; v[14] = o[2];
        v[15] = o[0] * o[0]; v[16] = o[0] * o[1]; v[17] = o[0] * o[2];
        v[18] = o[1] * o[1]; v[19] = o[1] * o[2]; v[20] = o[2] * o[2];
        break;
    case 2:
        v[2] = 1.;
        v[3] = -o[2];
        v[6] = -2. * o[0] * o[2]; v[7] = -1. * o[1] * o[2]; v[8] = -0.5 * o[2] * o[2];
        v[13] = -o[2];
        v[16] = -1. * o[0] * o[2];
        v[18] = -2. * o[1] * o[2]; v[19] = -0.5 * o[2] * o[2];
        v[21] = o[0]; v[22] = o[1]; v[23] = o[0] * o[0];
        v[24] = o[0] * o[1]; v[25] = o[1] * o[1];
        break;
    }
}

// SIM_RawField::getValue(indexToPos(sample)) -- SHIM: trilinear
// interpolation of the centre-sampled fp32 viscosity at the sample position; the index-space
// fractions are exactly 0 or 1/2; lerp x, then y, then z as a + t*(b-a) in fp32, clamp-to-edge.
float Oracle::getLocalViscosity(int samp, const I3& idx) const {
    int off2[3] = {1, 1, 1};
    if (samp >= S_FACEX && samp <= S_FACEZ) off2[samp - S_FACEX] = 0;
    else if (samp >= S_EDGEYZ) { int e = samp - S_EDGEYZ; for (int a = 0; a < 3; ++a) if (a != e) off2[a] = 0; }
    int base[3]; float t[3];
    for (int a = 0; a < 3; ++a) {
        if (off2[a] == 1) { base[a] = idx[a]; t[a] = 0.f; }
        else { base[a] = idx[a] - 1; t[a] = 0.5f; }
    }
    float c[2][2][2];
    for (int dz = 0; dz < 2; ++dz) for (int dy = 0; dy < 2; ++dy) for (int dxx = 0; dxx < 2; ++dxx)
        c[dz][dy][dxx] = viscosity.get(base[0] + dxx, base[1] + dy, base[2] + dz);
    float cy[2][2];
    for (int dz = 0; dz < 2; ++dz) for (int dy = 0; dy < 2; ++dy) cy[dz][dy] = c[dz][dy][0] + t[0] * (c[dz][dy][1] - c[dz][dy][0]);
    float cz[2];
    for (int dz = 0; dz < 2; ++dz) cz[dz] = cy[dz][0] + t[1] * (cy[dz][1] - cy[dz][0]);
    return cz[0] + t[2] * (cz[1] - cz[0]);
}

// S.cpp:328-372 + 1274-1324
void Oracle::computeCenterOfMasses() {
    const Field<exint>& L = labels[S_CENTER];
    const Field<exint>& R = reducedIdx[S_CENTER];
    com.assign(regionCount, {0., 0., 0.});
    std::vector<Real> count(regionCount, 0.);
    const int res[3] = {nx, ny, nz};
    forEachTileOrder(res, [&](int i, int j, int k) {
        if (!isReduced(L.at(i, j, k))) return;
        const exint r = R.at(i, j, k);
        count[r] += 1.; com[r][0] += (Real)i; com[r][1] += (Real)j; com[r][2] += (Real)k;
    });
    for (exint r = 0; r < regionCount; ++r) { const Real s = dx / count[r]; for (int a = 0; a < 3; ++a) com[r][a] *= s; }
}

static inline void addOuter(Real* M, Real s, const Real* c, const Real* rvec) {
    for (int i = 0; i < RDOF; ++i) { const Real ci = s * c[i]; if (ci == 0.) continue; for (int j = 0; j < RDOF; ++j) M[i * RDOF + j] += ci * rvec[j]; }
}

// S.cpp:374-417 + 1330-1399
void Oracle::computeLeastSquaresFits() {
    const Field<exint>& L = labels[S_CENTER];
    const Field<exint>& R = reducedIdx[S_CENTER];
    std::vector<std::array<Real, RDOF * RDOF>> N(regionCount);
    std::vector<std::array<Real, RDOF>> rhs(regionCount);
    for (auto& m : N) m.fill(0.);
    for (auto& v : rhs) v.fill(0.);
    const int res[3] = {nx, ny, nz};
    forEachTileOrder(res, [&](int i, int j, int k) {
        const exint r = R.at(i, j, k);
        if (r < 0) return;
        I3 cell{{i, j, k}};
        for (int axis = 0; axis < 3; ++axis) for (int dir = 0; dir < 2; ++dir) {
            I3 adj = cellToCellMap(cell, axis, dir);
            if (!isActive(L.get(adj))) continue;
            Real off[3] = {(Real)i, (Real)j, (Real)k};
            off[axis] += (dir == 0) ? -.5 : .5;
            for (int a = 0; a < 3; ++a) { off[a] *= dx; off[a] -= com[r][a]; }
            Real c[RDOF]; buildConversionCoefficients(off, axis, c);
            addOuter(N[r].data(), 1., c, c);
            const Real u = (Real)vel[axis].get(cellToFaceMap(cell, axis, dir));
            for (int n = 0; n < RDOF; ++n) rhs[r][n] += u * c[n];
        }
    });
    bestFit.assign(regionCount, {});
    for (exint r = 0; r < regionCount; ++r) { int rank; solveFullPivLU(N[r].data(), rhs[r].data(), bestFit[r].data(), RDOF, &rank); }
}

// S.cpp:419-441 + 1405-1482
void Oracle::computeReducedMassMatrices() {
    const Field<exint>& L = labels[S_CENTER];
    const Field<exint>& R = reducedIdx[S_CENTER];
    Mr.assign(regionCount, {});
    for (auto& m : Mr) m.fill(0.);
    const int res[3] = {nx, ny, nz};
    forEachTileOrder(res, [&](int i, int j, int k) {
        const exint r = R.at(i, j, k);
        if (r < 0) return;
        I3 cell{{i, j, k}};
        for (int axis = 0; axis < 3; ++axis) for (int dir = 0; dir < 2; ++dir) {
            bool doApplyFace = (dir == 0);
            if (!doApplyFace) { I3 adj = cellToCellMap(cell, axis, dir); if (isActive(L.get(adj))) doApplyFace = true; }
            if (!doApplyFace) continue;
            Real off[3] = {(Real)i, (Real)j, (Real)k};
            off[axis] += (dir == 0) ? -.5 : .5;
            for (int a = 0; a < 3; ++a) { off[a] *= dx; off[a] -= com[r][a]; }
            Real c[RDOF]; buildConversionCoefficients(off, axis, c);
            // myConstantDensity * columnVector * columnVector.transpose(): (rho*c) c^T
            addOuter(Mr[r].data(), P.density, c, c);
        }
    });
}

// S.cpp:468-490 + 1484-1694
void Oracle::computeReducedViscosityMatricesInteriorOnly() {
    const Field<exint>& CL = labels[S_CENTER];
    Visc.assign(regionCount, {});
    for (auto& m : Visc) m.fill(0.);
    for (int faceAxis = 0; faceAxis < 3; ++faceAxis) {
        const Field<exint>& FR = reducedIdx[S_FACEX + faceAxis];
        forEachTileOrder(FR.r, [&](int i, int j, int k) {
            const exint self = FR.at(i, j, k);
            if (self < 0) return;
            I3 face{{i, j, k}};
            Real selfOff[3] = {(Real)i, (Real)j, (Real)k};
            selfOff[faceAxis] -= 0.5;
            for (int a = 0; a < 3; ++a) { selfOff[a] *= dx; selfOff[a] -= com[self][a]; }
            Real colVec[RDOF]; buildConversionCoefficients(selfOff, faceAxis, colVec);
            // cell-centred stress terms
            for (int divDir = 0; divDir < 2; ++divDir) {
                I3 cell = faceToCellMap(face, faceAxis, divDir);
                if (!isReduced(CL.get(cell))) continue;
                if (cell[faceAxis] < 0 || cell[faceAxis] >= FR.r[faceAxis]) continue;
                const Real divSign = (divDir == 0) ? -1 : 1;
                const Real visc = (Real)getLocalViscosity(S_CENTER, cell);
                for (int gradDir = 0; gradDir < 2; ++gradDir) {
                    I3 adjFace = cellToFaceMap(cell, faceAxis, gradDir);
                    const Real gradSign = (gradDir == 0) ? -1. : 1.;
                    const Real contribution = -1. * divSign * gradSign * visc / (dx * dx);
                    const exint adjR = FR.get(adjFace);
                    if (adjR < 0) continue;
                    Real adjOff[3] = {(Real)adjFace[0], (Real)adjFace[1], (Real)adjFace[2]};
                    adjOff[faceAxis] -= 0.5;
                    for (int a = 0; a < 3; ++a) { adjOff[a] *= dx; adjOff[a] -= com[adjR][a]; }
                    Real rowVec[RDOF]; buildConversionCoefficients(adjOff, faceAxis, rowVec);
                    addOuter(Visc[self].data(), contribution, colVec, rowVec);
                }
            }
            // edge-centred stress terms
            for (int edgeAxis = 0; edgeAxis < 3; ++edgeAxis) {
                if (edgeAxis == faceAxis) continue;
                for (int divDir = 0; divDir < 2; ++divDir) {
                    const Real divSign = (divDir == 0) ? -1 : 1;
                    I3 edge = faceToEdgeMap(face, faceAxis, edgeAxis, divDir);
                    const float visc = getLocalViscosity(S_EDGEYZ + edgeAxis, edge);
                    if (labels[S_EDGEYZ + edgeAxis].get(edge) != REDUCED) continue;   // isReducedButNotBoundary
                    for (int gradAxis = 0; gradAxis < 3; ++gradAxis) {
                        if (gradAxis == edgeAxis) continue;
                        const int adjFaceAxis = 3 - gradAxis - edgeAxis;
                        const Field<exint>& AFR = reducedIdx[S_FACEX + adjFaceAxis];
                        for (int gradDir = 0; gradDir < 2; ++gradDir) {
                            I3 adjFace = edgeToFaceMap(edge, edgeAxis, adjFaceAxis, gradDir);
                            const Real gradSign = (gradDir == 0) ? -1 : 1;
                            const Real contribution = -0.5 * divSign * gradSign * visc / (dx * dx);
                            const int adjR = (int)AFR.get(adjFace);
                            if (adjR < 0) continue;
                            Real adjOff[3] = {(Real)adjFace[0], (Real)adjFace[1], (Real)adjFace[2]};
                            adjOff[adjFaceAxis] -= 0.5;
                            for (int a = 0; a < 3; ++a) { adjOff[a] *= dx; adjOff[a] -= com[adjR][a]; }
                            Real rowVec[RDOF]; buildConversionCoefficients(adjOff, adjFaceAxis, rowVec);
                            addOuter(Visc[self].data(), contribution, colVec, rowVec);
                        }
                    }
                }
            }
        });
    }
}

static inline Real sysClamp(Real v, Real lo, Real hi) { return std::min(std::max(v, lo), hi); }

// S_CMB:9-292 (dimensions + setFromTriplets) and 294-868 (buildMatrixBlocksByTriplets), executed in
// the single-thread order: face axes 0,1,2 (tile order), then edge axes 0,1,2, then centres.
void Oracle::constructMatrixBlocks() {
    const Real MINWEIGHT = 0.1;   // S.h:226
    nActiveVs = nFace[0] + nFace[1] + nFace[2];
    nReducedVs = regionCount * RDOF;
    nPressures = nCenter;
    nStresses = 3 * nCenter + nEdge[0] + nEdge[1] + nEdge[2];
    nTotalDOFs = nActiveVs + nReducedVs + nPressures + nStresses;

    std::vector<Triplet> tMc, tMcInv, tUInv, tU, tG, tJG, tDt, tJDt;
    activeRHS.assign(nActiveVs, 0.); pressureRHS.assign(nPressures, 0.); stressRHS.assign(nStresses, 0.); oldActiveVs.assign(nActiveVs, 0.);

    const Field<exint>& CL = labels[S_CENTER];
    const Field<exint>& CA = activeIdx[S_CENTER];
    for (int faceAxis = 0; faceAxis < 3; ++faceAxis) {
        const Field<exint>& FL = labels[S_FACEX + faceAxis];
        const Field<exint>& FA = activeIdx[S_FACEX + faceAxis];
        const Field<exint>& FR = reducedIdx[S_FACEX + faceAxis];
        const Field<float>& FFW = fluidW[S_FACEX + faceAxis];
        const Field<float>& FLW = liquidW[S_FACEX + faceAxis];
        forEachTileOrder(FL.r, [&](int i, int j, int k) {
            I3 face{{i, j, k}};
            const exint selfLabel = FL.at(face);
            const exint selfActiveIndex = faceVelocityDOF(FA.at(face), faceAxis);
            const exint selfReducedIndex = FR.at(face);
            const Real localDensity = P.density;
            Real volume = (Real)FFW.at(face) * (Real)FLW.at(face);
            volume = sysClamp(volume, MINWEIGHT * MINWEIGHT, 1.0);
            const Real localVelocity = (Real)vel[faceAxis].at(face);

            if (isActive(selfLabel)) {
                tMc.push_back({selfActiveIndex, selfActiveIndex, volume * localDensity});
                tMcInv.push_back({selfActiveIndex, selfActiveIndex, 1. / (volume * localDensity)});
                activeRHS[selfActiveIndex] += localVelocity * volume * localDensity;
                oldActiveVs[selfActiveIndex] += localVelocity;
            }
            if (!(isActive(selfLabel) || isReduced(selfLabel))) return;

            Real colVec[RDOF];
            if (!isActive(selfLabel)) {
                Real off[3] = {(Real)i, (Real)j, (Real)k};
                off[faceAxis] -= 0.5;
                for (int a = 0; a < 3; ++a) { off[a] *= dx; off[a] -= com[selfReducedIndex][a]; }
                buildConversionCoefficients(off, faceAxis, colVec);
            }
            // pressure stencils (S_CMB:393-460)
            for (int gradDir = 0; gradDir < 2; ++gradDir) {
                const Real gradSign = (gradDir == 0) ? -1 : 1;
                I3 cell = faceToCellMap(face, faceAxis, gradDir);
                if (cell[faceAxis] < 0 || cell[faceAxis] >= FA.r[faceAxis]) continue;
                const exint cellPressureIndex = CA.get(cell);
                if (cellPressureIndex < 0) continue;
                const Real coeff = (Real)FFW.at(face) * (Real)liquidW[S_CENTER].get(cell) * (Real)invDx;
                const Real contribution = gradSign * coeff;
                if (coeff <= 0.) continue;
                if (isActive(selfLabel)) {
                    tG.push_back({selfActiveIndex, cellPressureIndex, contribution});
                    if (fluidW[S_CENTER].get(cell) < 1.f) {
                        const Real solidContribution = gradSign * coeff;
                        const Real svel = (Real)colVel[faceAxis].at(face);
                        pressureRHS[cellPressureIndex] += -1. * solidContribution * svel;
                    }
                    if (FFW.at(face) < 1.f) {
                        const Real solidContribution = gradSign * coeff;
                        const Real svel = (Real)colVel[faceAxis].at(face);
                        pressureRHS[cellPressureIndex] += solidContribution * svel;
                    }
                } else {
                    for (int n = 0; n < RDOF; ++n) tJG.push_back({RDOF * selfReducedIndex + n, cellPressureIndex, contribution * colVec[n]});
                }
            }
            // stress stencils: centres (S_CMB:465-550)
            for (int divDir = 0; divDir < 2; ++divDir) {
                const Real divSign = (divDir == 0) ? -1 : 1;
                I3 cell = faceToCellMap(face, faceAxis, divDir);
                if (cell[faceAxis] < 0 || cell[faceAxis] >= FA.r[faceAxis]) continue;
                const exint cellLabel = CL.get(cell);
                const exint cellStressIndex = centerStressDOF(CA.get(cell), faceAxis);
                if (!isActive(cellLabel)) continue;
                const Real coeff = (Real)FFW.at(face) * (Real)liquidW[S_CENTER].get(cell) * (Real)invDx;
                const Real contribution = -1. * divSign * coeff;
                if (coeff <= 0.) continue;
                if (isActive(selfLabel)) {
                    tDt.push_back({selfActiveIndex, cellStressIndex, contribution});
                    if (fluidW[S_CENTER].get(cell) < 1.f) {
                        const Real solidContribution = divSign * coeff;
                        const Real svel = (Real)colVel[faceAxis].at(face);
                        stressRHS[cellStressIndex] += -1. * solidContribution * svel;
                    }
                    if (FFW.at(face) < 1.f) {
                        const Real solidContribution = divSign * coeff;
                        const Real svel = (Real)colVel[faceAxis].at(face);
                        stressRHS[cellStressIndex] += solidContribution * svel;
                    }
                } else {
                    for (int n = 0; n < RDOF; ++n) tJDt.push_back({RDOF * selfReducedIndex + n, cellStressIndex, contribution * colVec[n]});
                }
            }
            // stress stencils: edges (S_CMB:552-639)
            for (int edgeAxis = 0; edgeAxis < 3; ++edgeAxis) {
                if (edgeAxis == faceAxis) continue;
                for (int divDir = 0; divDir < 2; ++divDir) {
                    const Real divSign = (divDir == 0) ? -1 : 1;
                    I3 edge = faceToEdgeMap(face, faceAxis, edgeAxis, divDir);
                    const exint edgeLabel = labels[S_EDGEYZ + edgeAxis].get(edge);
                    const exint edgeStressIndex = edgeStressDOF(activeIdx[S_EDGEYZ + edgeAxis].get(edge), edgeAxis);
                    if (!isActive(edgeLabel)) continue;
                    const Real coeff = (Real)FFW.at(face) * (Real)liquidW[S_EDGEYZ + edgeAxis].get(edge) * (Real)invDx;
                    const Real contribution = -1. * divSign * coeff;
                    if (coeff <= 0.) continue;
                    if (isActive(selfLabel)) {
                        tDt.push_back({selfActiveIndex, edgeStressIndex, contribution});
                        if (fluidW[S_EDGEYZ + edgeAxis].get(edge) < 1.f) {
                            const Real solidContribution = divSign * coeff;
                            const Real svel = (Real)colVel[faceAxis].at(face);
                            stressRHS[edgeStressIndex] += -1. * solidContribution * svel;
                        }
                        if (FFW.at(face) < 1.f) {
                            const Real solidContribution = divSign * coeff;
                            const Real svel = (Real)colVel[faceAxis].at(face);
                            stressRHS[edgeStressIndex] += solidContribution * svel;
                        }
                    } else {
                        for (int n = 0; n < RDOF; ++n) tJDt.push_back({RDOF * selfReducedIndex + n, edgeStressIndex, contribution * colVec[n]});
                    }
                }
            }
        });
    }
    // edge stress diagonal terms (S_CMB:650-735)
    for (int edgeAxis = 0; edgeAxis < 3; ++edgeAxis) {
        const Field<exint>& EL = labels[S_EDGEYZ + edgeAxis];
        forEachTileOrder(EL.r, [&](int i, int j, int k) {
            I3 edge{{i, j, k}};
            const exint edgeLabel = EL.at(edge);
            if (!isActive(edgeLabel)) return;
            const exint edgeStressIndex = edgeStressDOF(activeIdx[S_EDGEYZ + edgeAxis].at(edge), edgeAxis);
            const Real volumeWeight = sysClamp((Real)fluidW[S_EDGEYZ + edgeAxis].at(edge), MINWEIGHT, 1.0) * (Real)liquidW[S_EDGEYZ + edgeAxis].at(edge);
            const Real localViscosity = (Real)getLocalViscosity(S_EDGEYZ + edgeAxis, edge);
            const Real invLocalViscosity = sysClamp(1. / localViscosity, 0., 1e10);
            tUInv.push_back({edgeStressIndex, edgeStressIndex, 2. * invLocalViscosity * volumeWeight});
            tU.push_back({edgeStressIndex, edgeStressIndex, 0.5 * localViscosity * sysClamp(1. / volumeWeight, 0., 1.e2)});
        });
    }
    // centre stress diagonal terms (S_CMB:736-867)
    {
        const int res[3] = {nx, ny, nz};
        forEachTileOrder(res, [&](int i, int j, int k) {
            I3 cell{{i, j, k}};
            if (!isActive(CL.at(cell))) return;
            const exint ci = CA.at(cell);
            const Real volumeWeight = sysClamp((Real)fluidW[S_CENTER].at(cell), MINWEIGHT, 1.0) * (Real)liquidW[S_CENTER].at(cell);
            const Real localViscosity = (Real)getLocalViscosity(S_CENTER, cell);
            const Real invLocalViscosity = sysClamp(1. / localViscosity, 0., 1.e10);
            for (int a = 0; a < 3; ++a) {
                const exint s = centerStressDOF(ci, a);
                tUInv.push_back({s, s, invLocalViscosity * sysClamp(volumeWeight, 1.e-2, 1.)});
                tU.push_back({s, s, localViscosity * sysClamp(1. / volumeWeight, 0., 1.e2)});
            }
        });
    }
    Mc.resize(nActiveVs, nActiveVs); Mc.setFromTriplets(tMc);
    McInv.resize(nActiveVs, nActiveVs); McInv.setFromTriplets(tMcInv);
    uInv.resize(nStresses, nStresses); uInv.setFromTriplets(tUInv);
    uMat.resize(nStresses, nStresses); uMat.setFromTriplets(tU);
    G.resize(nActiveVs, nPressures); G.setFromTriplets(tG);
    JG.resize(nReducedVs, nPressures); JG.setFromTriplets(tJG);
    Dt.resize(nActiveVs, nStresses); Dt.setFromTriplets(tDt);
    JDt.resize(nReducedVs, nStresses); JDt.setFromTriplets(tJDt);
}

// S_AS:432-470 with S_AB:3-48 (Mr), 147-193 (B), 195-244 (B^-1), 356-367 (reduced rhs)
void Oracle::assembleSystemPressureStressFactored() {
    nSystemSize = nPressures + nStresses;
    Binv.assign(regionCount, {});
    reducedRHS.assign(nReducedVs, 0.);
    std::vector<Triplet> tMr, tB, tBinv;
    tMr.reserve(regionCount * RDOF * RDOF); tB.reserve(regionCount * RDOF * RDOF); tBinv.reserve(regionCount * RDOF * RDOF);
    for (exint r = 0; r < regionCount; ++r) {
        std::array<Real, RDOF * RDOF> B;
        for (int q = 0; q < RDOF * RDOF; ++q) B[q] = invDt * Mr[r][q] + 2. * Visc[r][q];
        inversePartialPivLU(B.data(), Binv[r].data(), RDOF);
        for (int i = 0; i < RDOF; ++i) {
            Real s = 0;
            for (int j = 0; j < RDOF; ++j) s += Mr[r][i * RDOF + j] * bestFit[r][j];
            reducedRHS[RDOF * r + i] = s;
            for (int j = 0; j < RDOF; ++j) {
                tMr.push_back({RDOF * r + i, RDOF * r + j, Mr[r][i * RDOF + j]});
                tB.push_back({RDOF * r + i, RDOF * r + j, B[i * RDOF + j]});
                tBinv.push_back({RDOF * r + i, RDOF * r + j, Binv[r][i * RDOF + j]});
            }
        }
    }
    MrMat.resize(nReducedVs, nReducedVs); MrMat.setFromTriplets(tMr);
    Bmat.resize(nReducedVs, nReducedVs); Bmat.setFromTriplets(tB);
    BinvMat.resize(nReducedVs, nReducedVs); BinvMat.setFromTriplets(tBinv);

    // b = -[G^T; D] McInv rhs_u - (1/dt) [JG^T; DJ^T] B^-1 rhs_r + [rhs_p; rhs_tau]   (S_AS:448-459)
    Gt = G.transpose(); D = Dt.transpose(); JGt = JG.transpose(); DJt = JDt.transpose();
    mcInvDiag.assign(nActiveVs, 0.);
    for (exint r = 0; r < McInv.rows; ++r) for (exint p = McInv.ptr[r]; p < McInv.ptr[r + 1]; ++p) if (McInv.idx[p] == r) mcInvDiag[r] = McInv.val[p];
    uInvDiag.assign(nStresses, 0.);
    for (exint r = 0; r < uInv.rows; ++r) for (exint p = uInv.ptr[r]; p < uInv.ptr[r + 1]; ++p) if (uInv.idx[p] == r) uInvDiag[r] = uInv.val[p];

    std::vector<Real> t(nActiveVs), s(nReducedVs), b1(nPressures), b2(nStresses), tmp;
    for (exint i = 0; i < nActiveVs; ++i) t[i] = mcInvDiag[i] * activeRHS[i];
    BinvMat.mulVec(reducedRHS.data(), s.data());
    Gt.mulVec(t.data(), b1.data());
    for (auto& v : b1) v = -v;
    JGt.mulVecAdd(s.data(), b1.data(), -invDt);
    D.mulVec(t.data(), b2.data());
    for (auto& v : b2) v = -v;
    DJt.mulVecAdd(s.data(), b2.data(), -invDt);
    b.assign(nSystemSize, 0.);
    for (exint i = 0; i < nPressures; ++i) b[i] = b1[i] + pressureRHS[i];
    for (exint i = 0; i < nStresses; ++i) b[nPressures + i] = b2[i] + stressRHS[i];
    solution.assign(nSystemSize, 0.);
}

// S_AS:381-397: explicit A by sparse triple products (structure = union, no pruning)
void Oracle::assembleExplicitA() {
    Csr McInvG_ = scaleRows(mcInvDiag, G), McInvDt_ = scaleRows(mcInvDiag, Dt);
    Csr BinvJG = spgemm(BinvMat, JG), BinvJDt = spgemm(BinvMat, JDt);
    Csr A11 = addScaled(-dt, spgemm(Gt, McInvG_), -1., spgemm(JGt, BinvJG));
    Csr A12 = addScaled(-dt, spgemm(Gt, McInvDt_), -1., spgemm(JGt, BinvJDt));
    Csr A21 = addScaled(-dt, spgemm(D, McInvG_), -1., spgemm(DJt, BinvJG));
    Csr A22 = addScaled(1., addScaled(-dt, spgemm(D, McInvDt_), -1., spgemm(DJt, BinvJDt)), -0.5, uInv);
    std::vector<Triplet> t;
    auto addBlock = [&](const Csr& M, exint ro, exint co) {   // util.h:105-120
        for (exint r = 0; r < M.rows; ++r) for (exint p = M.ptr[r]; p < M.ptr[r + 1]; ++p) t.push_back({r + ro, M.idx[p] + co, 1. * M.val[p]});
    };
    addBlock(A11, 0, 0); addBlock(A12, 0, nPressures); addBlock(A21, nPressures, 0); addBlock(A22, nPressures, nPressures);
    A.resize(nSystemSize, nSystemSize); A.setFromTriplets(t);
}

// exec/HDK_PolyStokes.C:344-476: the setup half of solveGasSubclass
void Oracle::runSetup() {
    auto t0 = std::chrono::steady_clock::now();
    buildIntegrationWeightsAlt();
    classifyCells();
    if (P.doReduced) constructReducedRegions(); else constructOnlyActiveRegions();
    classifyFaces();
    classifyEdges();
    if (P.doReduced) { constructCenterReducedIndices(); constructFacesReducedIndices(); constructEdgesReducedIndices(); }
    constructActiveIndices();
    if (P.doReduced) {
        computeCenterOfMasses(); computeLeastSquaresFits();
        computeReducedMassMatrices(); computeReducedViscosityMatricesInteriorOnly();
    } else { regionCount = 0; com.clear(); bestFit.clear(); Mr.clear(); Visc.clear(); }
    constructMatrixBlocks();
    assembleSystemPressureStressFactored();
    setupMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace orc
