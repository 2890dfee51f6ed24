// ps_oracle_classify.cpp -- TEST INFRASTRUCTURE ONLY (see ps_oracle.hpp header).
// Integration weights + the Classifier (exec/HDK_PolyStokesSolver_Classifier.cpp), restated.
#include "ps_oracle.hpp"
#include <limits>

namespace orc {

Oracle::Oracle(const Params& p) : P(p), nx(p.nx), ny(p.ny), nz(p.nz), dx(p.dx), invDx(1. / p.dx), dt(p.dt), invDt(1. / p.dt) {
    // S.cpp:69-152: every field UNASSIGNED / 0, index fields with constant -1 border
    for (int s = 0; s < 7; ++s) {
        int r[3]; resOf(s, r);
        liquidW[s].init(r[0], r[1], r[2], 0.f); liquidW[s].clampBorder = true;
        fluidW[s].init(r[0], r[1], r[2], 0.f); fluidW[s].clampBorder = true;
        labels[s].init(r[0], r[1], r[2], (exint)UNASSIGNED);
        activeIdx[s].init(r[0], r[1], r[2], (exint)UNASSIGNED);
        reducedIdx[s].init(r[0], r[1], r[2], (exint)UNASSIGNED);
    }
}

void Oracle::resOf(int samp, int r[3]) const {
    r[0] = nx; r[1] = ny; r[2] = nz;
    if (samp >= S_FACEX && samp <= S_FACEZ) r[samp - S_FACEX] += 1;
    else if (samp >= S_EDGEYZ) { int e = samp - S_EDGEYZ; for (int a = 0; a < 3; ++a) if (a != e) r[a] += 1; }
}

void Oracle::setInputs(const float* surf, const float* col, const float* visc, const float* const v[3], const float* const cv[3]) {
    surface.init(nx, ny, nz, 0.f); collision.init(nx, ny, nz, 0.f); viscosity.init(nx, ny, nz, 0.f);
    surface.clampBorder = collision.clampBorder = viscosity.clampBorder = true;
    std::copy(surf, surf + surface.size(), surface.d.begin());
    std::copy(col, col + collision.size(), collision.d.begin());
    std::copy(visc, visc + viscosity.size(), viscosity.d.begin());
    for (int a = 0; a < 3; ++a) {
        int r[3]; resOf(S_FACEX + a, r);
        vel[a].init(r[0], r[1], r[2], 0.f); colVel[a].init(r[0], r[1], r[2], 0.f);
        vel[a].clampBorder = colVel[a].clampBorder = true;
        std::copy(v[a], v[a] + vel[a].size(), vel[a].d.begin());
        std::copy(cv[a], cv[a] + colVel[a].size(), colVel[a].d.begin());
    }
}

// S.h:193-222 SamplingOffset, doubled so it stays integral: position of sample idx in
// half-cell units relative to the cell-centre lattice is 2*idx + off2 - 1.
static inline void samplingOffset2(int samp, int off2[3]) {
    off2[0] = off2[1] = off2[2] = 1;                      // CENTER (.5,.5,.5)
    if (samp >= S_FACEX && samp <= S_FACEZ) off2[samp - S_FACEX] = 0;         // FACEX (0,.5,.5) ...
    else if (samp >= S_EDGEYZ) { int e = samp - S_EDGEYZ; for (int a = 0; a < 3; ++a) if (a != e) off2[a] = 0; }  // EDGEXY (0,0,.5) ...
}

// HDK SIM_RawField::computeSDFWeightsSampled(sdf, 2, false, 0) -- SHIM:
// weight = #{2x2x2 sub-samples at +-dx/4 around the sample position with trilinear sdf < 0} / 8
// (for `invert`: sdf >= 0).  All interpolation fractions are exactly 1/4 or 3/4, the trilinear
// sum is evaluated in double where every product is exact.
static void sdfWeightsSampled(Field<float>& out, const Field<float>& sdf, int samp, bool countNonNegative) {
    int off2[3]; samplingOffset2(samp, off2);
    const int rx = out.r[0], ry = out.r[1], rz = out.r[2];
#pragma omp parallel for schedule(static)
    for (int k = 0; k < rz; ++k)
        for (int j = 0; j < ry; ++j)
            for (int i = 0; i < rx; ++i) {
                const int idx[3] = {i, j, k};
                int count = 0;
                for (int sz = 0; sz < 2; ++sz) for (int sy = 0; sy < 2; ++sy) for (int sx = 0; sx < 2; ++sx) {
                    const int s[3] = {sx, sy, sz};
                    int base[3]; double fr[3];
                    for (int a = 0; a < 3; ++a) {
                        // quarter-cell coordinate relative to cell centres: 4*idx + 2*off2 - 2 + (-1 | +1)
                        const int q = 4 * idx[a] + 2 * off2[a] - 2 + (s[a] ? 1 : -1);
                        const int fl = (q >= 0) ? q / 4 : -((-q + 3) / 4);
                        base[a] = fl; fr[a] = (q - 4 * fl) * 0.25;
                    }
                    double acc = 0.0;
                    for (int dz = 0; dz < 2; ++dz) for (int dy = 0; dy < 2; ++dy) for (int dxx = 0; dxx < 2; ++dxx) {
                        const double w = (dxx ? fr[0] : 1.0 - fr[0]) * (dy ? fr[1] : 1.0 - fr[1]) * (dz ? fr[2] : 1.0 - fr[2]);
                        acc += w * (double)sdf.get(base[0] + dxx, base[1] + dy, base[2] + dz);
                    }
                    if (countNonNegative ? (acc >= 0.0) : (acc < 0.0)) ++count;
                }
                out.at(i, j, k) = (float)count * 0.125f;
            }
}

// S.cpp:238-289 buildIntegrationWeightsAlt -> computeIntegrationWeights (291-305) and
// computeSolidIntegrationWeights (307-326).  Liquid weights: 1 inside liquid; fluid weights: 1 in
// fluid, 0 in solid (S.cpp:172, S_Cls:106-107) => counted where the collision SDF is >= 0.
void Oracle::buildIntegrationWeightsAlt() {
    for (int s = 0; s < 7; ++s) {
        sdfWeightsSampled(liquidW[s], surface, s, false);
        sdfWeightsSampled(fluidW[s], collision, s, true);
    }
}

// S_Cls:56-128
void Oracle::classifyCells() {
    Field<exint>& L = labels[S_CENTER];
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
        bool isInSolve = false, isInFluid = true;
        I3 cell{{i, j, k}};
        if (liquidW[S_CENTER].get(cell) > 0.f) isInSolve = true;
        if (!isInSolve)
            for (int axis = 0; axis < 3; ++axis) for (int dir = 0; dir < 2; ++dir) {
                I3 face = cellToFaceMap(cell, axis, dir);
                if (liquidW[S_FACEX + axis].get(face) > 0.f) isInSolve = true;
            }
        if (fluidW[S_CENTER].get(cell) == 0.f) isInFluid = false;
        if (isInSolve) L.at(cell) = isInFluid ? (exint)GENERICFLUID : (exint)SOLID;
        else L.at(cell) = UNSOLVED;
    }
}

void Oracle::overwriteIndices(Field<exint>& f, exint search, exint replace) {
    for (auto& v : f.d) if (v == search) v = replace;
}

// S_Cls:179-190
void Oracle::constructReducedRegions() {
    constructAirBoundaryLayer();
    constructSolidBoundaryLayer();
    if (P.doTile) constructTiles();
    overwriteIndices(labels[S_CENTER], GENERICFLUID, REDUCED);
}
// S_Cls:192-199
void Oracle::constructOnlyActiveRegions() { overwriteIndices(labels[S_CENTER], GENERICFLUID, ACTIVEFLUID); }

// S_Cls:291-362 (driver), 364-430 (initial layer), 432-508 (next layer), S.cpp:2022-2060 (mark)
void Oracle::constructAirBoundaryLayer() {
    Field<exint>& L = labels[S_CENTER];
    const int res[3] = {nx, ny, nz};
    std::vector<I3> layer;
    // buildInitialAirBoundaryLayer
    forEachTileOrder(res, [&](int i, int j, int k) {
        if (L.at(i, j, k) != GENERICFLUID) return;
        I3 cell{{i, j, k}};
        bool isBoundaryCell = false;
        for (int axis = 0; axis < 3; ++axis) for (int dir = 0; dir < 2; ++dir) {
            I3 adj = cellToCellMap(cell, axis, dir);
            if (adj[axis] < 0 || adj[axis] >= res[axis]) continue;
            I3 face = cellToFaceMap(cell, axis, dir);
            if (L.get(adj) == UNSOLVED) isBoundaryCell = true;
            if (liquidW[S_FACEX + axis].get(face) < 1.f) isBoundaryCell = true;
        }
        if (isBoundaryCell) layer.push_back(cell);
    });
    for (int layerNo = 0; layerNo < P.liquidLayers - 1; ++layerNo) {
        // setActiveLayerCells (duplicates are skipped there; marking twice is idempotent)
        for (const I3& c : layer) L.at(c) = ACTIVEFLUID;
        if (layerNo < P.liquidLayers - 2) {
            std::vector<I3> next;
            for (const I3& cell : layer)
                for (int axis = 0; axis < 3; ++axis) for (int dir = 0; dir < 2; ++dir) {
                    I3 adj = cellToCellMap(cell, axis, dir);
                    if (adj[axis] < 0 || adj[axis] >= res[axis]) continue;
                    I3 face = cellToFaceMap(cell, axis, dir);
                    if (liquidW[S_FACEX + axis].at(face) > 0.f && L.get(adj) == GENERICFLUID) next.push_back(adj);
                }
            layer.swap(next);
        }
    }
}

// S_Cls:510-571 (driver), 573-641 (initial), 643-703 (next)
void Oracle::constructSolidBoundaryLayer() {
    Field<exint>& L = labels[S_CENTER];
    const int res[3] = {nx, ny, nz};
    std::vector<I3> layer;
    forEachTileOrder(res, [&](int i, int j, int k) {
        const exint lab = L.at(i, j, k);
        if (lab != GENERICFLUID && lab != ACTIVEFLUID) return;
        I3 cell{{i, j, k}};
        bool isBoundaryCell = false;
        for (int axis = 0; axis < 3; ++axis) for (int dir = 0; dir < 2; ++dir) {
            I3 adj = cellToCellMap(cell, axis, dir);
            if (adj[axis] < 0 || adj[axis] >= res[axis]) { isBoundaryCell = true; continue; }
            if (L.get(adj) == SOLID) isBoundaryCell = true;
        }
        if (isBoundaryCell) layer.push_back(cell);
    });
    Field<exint> visited; visited.init(nx, ny, nz, (exint)UNVISITED);
    for (int layerNo = 0; layerNo < P.solidLayers; ++layerNo) {
        for (const I3& c : layer) { L.at(c) = ACTIVEFLUID; visited.at(c) = VISITED; }
        if (layerNo < P.solidLayers - 1) {
            std::vector<I3> next;
            for (const I3& cell : layer)
                for (int axis = 0; axis < 3; ++axis) for (int dir = 0; dir < 2; ++dir) {
                    I3 adj = cellToCellMap(cell, axis, dir);
                    if (adj[axis] < 0 || adj[axis] >= res[axis]) continue;
                    I3 face = cellToFaceMap(cell, axis, dir);
                    if (liquidW[S_FACEX + axis].get(face) > 0.f) {
                        if (visited.get(adj) == UNVISITED && (L.get(adj) == ACTIVEFLUID || L.get(adj) == GENERICFLUID)) next.push_back(adj);
                    }
                }
            layer.swap(next);
        }
    }
}

// S_Cls:705-746
void Oracle::constructTiles() {
    Field<exint>& L = labels[S_CENTER];
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
        if (L.at(i, j, k) != GENERICFLUID) continue;
        for (int t = 0; t < P.tilePadding; ++t)
            if (i % P.tileSize == t || j % P.tileSize == t || k % P.tileSize == t) L.at(i, j, k) = ACTIVEFLUID;
    }
}

// S_Cls:784-832 findFaceLabelFromCenter
void Oracle::classifyFaces() {
    for (int axis = 0; axis < 3; ++axis) {
        Field<exint>& FL = labels[S_FACEX + axis];
        for (int k = 0; k < FL.r[2]; ++k) for (int j = 0; j < FL.r[1]; ++j) for (int i = 0; i < FL.r[0]; ++i) {
            exint retval = UNSOLVED;
            I3 face{{i, j, k}};
            bool isActiveVelocity = false;
            for (int dir = 0; dir < 2; ++dir) {
                I3 cell = faceToCellMap(face, axis, dir);
                if (!liquidW[S_CENTER].inb(cell[0], cell[1], cell[2])) continue;   // c_oob, S.h:294-296
                if (liquidW[S_CENTER].get(cell) > 0.f) isActiveVelocity = true;
            }
            if (!isActiveVelocity)
                for (int edgeAxis = 0; edgeAxis < 3 && !isActiveVelocity; ++edgeAxis) {
                    if (edgeAxis == axis) continue;
                    for (int dir = 0; dir < 2; ++dir) {
                        I3 edge = faceToEdgeMap(face, axis, edgeAxis, dir);
                        if (liquidW[S_EDGEYZ + edgeAxis].get(edge) > 0.f) { isActiveVelocity = true; break; }
                    }
                }
            if (isActiveVelocity) retval = (fluidW[S_FACEX + axis].get(face) < 0.5f) ? (exint)SOLID : (exint)GENERICFLUID;
            FL.at(face) = retval;
        }
    }
}

// S_Cls:1021-1067 findEdgeLabelFromFaceAlt.  For edge axis e the two face axes are the
// other two; each contributes face(i,j,k) and the face one step back along the *other* axis,
// the latter guarded by the face oob predicate (S.h:306-314).
void Oracle::classifyEdges() {
    for (int e = 0; e < 3; ++e) {
        Field<exint>& EL = labels[S_EDGEYZ + e];
        // reference order of the two face axes: XY -> (X,Y); XZ -> (X,Z); YZ -> (Y,Z)
        const int fa0 = (e == 0) ? 1 : 0, fa1 = (e == 2) ? 1 : 2;
        for (int k = 0; k < EL.r[2]; ++k) for (int j = 0; j < EL.r[1]; ++j) for (int i = 0; i < EL.r[0]; ++i) {
            I3 edge{{i, j, k}};
            bool insystem = (liquidW[S_EDGEYZ + e].get(edge) != 0.f) && (fluidW[S_EDGEYZ + e].get(edge) != 0.f);
            if (!insystem) { EL.at(edge) = UNSOLVED; continue; }
            auto faceOk = [&](int fa) {
                // face(i,j,k) && !oob(back) && face(back); back = one step back along the other face axis
                const int other = 3 - fa - e;
                I3 back = edge; back[other] -= 1;
                const Field<float>& W = liquidW[S_FACEX + fa];
                return (W.get(edge) != 0.f) && W.inb(back[0], back[1], back[2]) && (W.get(back) != 0.f);
            };
            insystem = faceOk(fa0) && faceOk(fa1);
            EL.at(edge) = insystem ? (exint)GENERICFLUID : (exint)UNSOLVED;
        }
    }
}

// HDK SIM_VolumetricConnectedComponentBuilder::buildConnectedComponents -- SHIM (BASELINE.md
// section 3): 6-connected components of label==REDUCED cells, two cells connected only through a
// face with liquid weight > 0, ids in order of first encounter in tile iteration order.
void Oracle::buildConnectedComponents() {
    Field<exint>& L = labels[S_CENTER];
    Field<exint>& R = reducedIdx[S_CENTER];
    const int res[3] = {nx, ny, nz};
    std::fill(R.d.begin(), R.d.end(), (exint)UNASSIGNED);
    exint count = 0;
    std::vector<I3> stack;
    forEachTileOrder(res, [&](int i, int j, int k) {
        if (L.at(i, j, k) != REDUCED || R.at(i, j, k) != UNASSIGNED) return;
        const exint id = count++;
        R.at(i, j, k) = id; stack.push_back(I3{{i, j, k}});
        while (!stack.empty()) {
            I3 c = stack.back(); stack.pop_back();
            for (int axis = 0; axis < 3; ++axis) for (int dir = 0; dir < 2; ++dir) {
                I3 adj = cellToCellMap(c, axis, dir);
                if (!L.inb(adj[0], adj[1], adj[2])) continue;
                if (L.at(adj) != REDUCED || R.at(adj) != UNASSIGNED) continue;
                I3 face = cellToFaceMap(c, axis, dir);
                if (!(liquidW[S_FACEX + axis].at(face) > 0.f)) continue;
                R.at(adj) = id; stack.push_back(adj);
            }
        }
    });
    regionCount = count;
}

// S_Cls:217-239
void Oracle::constructCenterReducedIndices() {
    buildConnectedComponents();
    fixReducedRegionBoundaries();
    fixSmallReducedRegions();
}

// S_Cls:1073-1172: serial, in place, in tile order, repeated until a sweep changes nothing.
void Oracle::fixReducedRegionBoundaries() {
    Field<exint>& L = labels[S_CENTER];
    Field<exint>& R = reducedIdx[S_CENTER];
    const int res[3] = {nx, ny, nz};
    bool done = false;
    fixLoops = 0; fixRemoved = 0;
    while (!done) {
        done = true; fixLoops++;
        forEachTileOrder(res, [&](int i, int j, int k) {
            if (L.at(i, j, k) != ACTIVEFLUID) return;
            I3 cell{{i, j, k}};
            bool applyFix = false, isBoundaryCell = false;
            exint adjacentInteriorRegion = 0;
            for (int axis = 0; axis < 3; ++axis) for (int dir = 0; dir < 2; ++dir) {
                I3 adj = cellToCellMap(cell, axis, dir);
                if (isReduced(L.get(adj))) {
                    if (!isBoundaryCell) { isBoundaryCell = true; adjacentInteriorRegion = R.get(adj); }
                    else if (R.get(adj) != adjacentInteriorRegion) applyFix = true;
                }
            }
            if (applyFix) {
                done = false;
                for (int axis = 0; axis < 3; ++axis) for (int dir = 0; dir < 2; ++dir) {
                    I3 adj = cellToCellMap(cell, axis, dir);
                    if (isReduced(L.get(adj))) { fixRemoved++; L.at(adj) = ACTIVEFLUID; R.at(adj) = UNASSIGNED; }
                }
            }
        });
    }
}

// S_Cls:1174-1262 (+1418-1467 bounding boxes, 1264-1313 remap)
void Oracle::fixSmallReducedRegions() {
    Field<exint>& L = labels[S_CENTER];
    Field<exint>& R = reducedIdx[S_CENTER];
    const exint n = regionCount;
    std::vector<std::array<exint, 3>> bbMin(n, {std::numeric_limits<exint>::max(), std::numeric_limits<exint>::max(), std::numeric_limits<exint>::max()});
    std::vector<std::array<exint, 3>> bbMax(n, {std::numeric_limits<exint>::min(), std::numeric_limits<exint>::min(), std::numeric_limits<exint>::min()});
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
        if (!isReduced(L.at(i, j, k))) continue;
        const exint r = R.at(i, j, k);
        const int c[3] = {i, j, k};
        for (int a = 0; a < 3; ++a) { bbMin[r][a] = std::min<exint>(bbMin[r][a], c[a]); bbMax[r][a] = std::max<exint>(bbMax[r][a], c[a]); }
    }
    std::vector<char> doRemove(n, 0);
    for (exint r = 0; r < n; ++r)
        for (int a = 0; a < 3; ++a) {
            if (bbMax[r][a] == bbMin[r][a]) doRemove[r] = 1;
            // guard the (max - 3) underflow of an empty region (all of whose cells were converted by
            // fixReducedRegionBoundaries): reference compares min > max-3 with min=INT64_MAX, max=INT64_MIN
            if (bbMax[r][a] == std::numeric_limits<exint>::min() || bbMin[r][a] > bbMax[r][a] - 3) doRemove[r] = 1;
        }
    std::vector<exint> remap(n, -1);
    exint regionMap = 0;
    for (exint r = 0; r < n; ++r) if (!doRemove[r]) remap[r] = regionMap++;
    if (n > regionMap) {
        regionCount = regionMap;
        for (size_t q = 0; q < L.d.size(); ++q) {
            if (!isReduced(L.d[q])) continue;
            const exint r = R.d[q];
            if (doRemove[r]) { L.d[q] = ACTIVEFLUID; R.d[q] = UNASSIGNED; }
            else if (remap[r] < r) R.d[q] = remap[r];
        }
    }
}

// S_Cls:1473-1528
void Oracle::constructFacesReducedIndices() {
    const Field<exint>& CL = labels[S_CENTER];
    const Field<exint>& CR = reducedIdx[S_CENTER];
    for (int axis = 0; axis < 3; ++axis) {
        Field<exint>& FL = labels[S_FACEX + axis];
        Field<exint>& FR = reducedIdx[S_FACEX + axis];
        for (int k = 0; k < FL.r[2]; ++k) for (int j = 0; j < FL.r[1]; ++j) for (int i = 0; i < FL.r[0]; ++i) {
            exint idx = UNASSIGNED;
            I3 c1{{i, j, k}}; I3 c0 = c1; c0[axis] -= 1;
            if (CL.inb(c1[0], c1[1], c1[2]) && CL.at(c1) == REDUCED) idx = CR.at(c1);
            else if (CL.inb(c0[0], c0[1], c0[2]) && CL.at(c0) == REDUCED) idx = CR.at(c0);
            if (idx != UNASSIGNED) { FL.at(i, j, k) = REDUCED; FR.at(i, j, k) = idx; }
        }
    }
}

// S_Cls:1534-1659.  Per edge axis the four faces are visited in the reference's fixed order;
// REDUCED edges take their index from the quirky slot listed at S_Cls:1569/1599/1629.
void Oracle::constructEdgesReducedIndices() {
    for (int e = 0; e < 3; ++e) {
        Field<exint>& EL = labels[S_EDGEYZ + e];
        Field<exint>& ER = reducedIdx[S_EDGEYZ + e];
        const int fa0 = (e == 0) ? 1 : 0, fa1 = (e == 2) ? 1 : 2;
        for (int k = 0; k < EL.r[2]; ++k) for (int j = 0; j < EL.r[1]; ++j) for (int i = 0; i < EL.r[0]; ++i) {
            I3 edge{{i, j, k}};
            // the 4 candidate faces in reference order: fa0(i,j,k), fa0(back), fa1(i,j,k), fa1(back)
            struct Cand { int fa; I3 f; } cand[4];
            cand[0] = {fa0, edge}; cand[1] = {fa0, edge}; cand[1].f[3 - fa0 - e] -= 1;
            cand[2] = {fa1, edge}; cand[3] = {fa1, edge}; cand[3].f[3 - fa1 - e] -= 1;
            bool red[4];
            for (int q = 0; q < 4; ++q) {
                const Field<exint>& FL = labels[S_FACEX + cand[q].fa];
                red[q] = FL.inb(cand[q].f[0], cand[q].f[1], cand[q].f[2]) && FL.at(cand[q].f) == REDUCED;
            }
            exint label = UNASSIGNED, idx = UNASSIGNED;
            if (red[0] && red[1] && red[2] && red[3]) {
                // XY: faceX(i,j,k); XZ: faceX(i,j,k); YZ: faceY(i,j-1,k)  (S_Cls:1569,1599,1629)
                if (e == 0) { I3 f = edge; f[1] -= 1; idx = reducedIdx[S_FACEY].get(f); }
                else idx = reducedIdx[S_FACEX].get(edge);
                label = REDUCED;
            } else {
                for (int q = 0; q < 4; ++q) if (red[q]) { idx = reducedIdx[S_FACEX + cand[q].fa].get(cand[q].f); label = BOUNDARY; break; }
            }
            if (idx != UNASSIGNED) { EL.at(edge) = label; ER.at(edge) = idx; }
        }
    }
}

// S_Cls:1738-1770
exint Oracle::serialAssignFieldIndices(Field<exint>& idx, const Field<exint>& lab) {
    exint n = 0;
    forEachTileOrder(idx.r, [&](int i, int j, int k) { if (isActive(lab.at(i, j, k))) idx.at(i, j, k) = n++; });
    return n;
}

// S_Cls:257-284
void Oracle::constructActiveIndices() {
    for (int s = 0; s < 7; ++s) overwriteIndices(labels[s], GENERICFLUID, ACTIVEFLUID);
    nCenter = serialAssignFieldIndices(activeIdx[S_CENTER], labels[S_CENTER]);
    for (int a = 0; a < 3; ++a) nFace[a] = serialAssignFieldIndices(activeIdx[S_FACEX + a], labels[S_FACEX + a]);
    for (int e = 0; e < 3; ++e) nEdge[e] = serialAssignFieldIndices(activeIdx[S_EDGEYZ + e], labels[S_EDGEYZ + e]);
}

// S_Cls:4-54
void Oracle::buildValidFaces(float* const valid[3]) {
    for (int axis = 0; axis < 3; ++axis) {
        const Field<exint>& FL = labels[S_FACEX + axis];
        for (size_t q = 0; q < FL.d.size(); ++q) valid[axis][q] = (FL.d[q] == UNSOLVED || FL.d[q] == UNASSIGNED) ? 0.f : 1.f;
    }
}

}  // namespace orc
