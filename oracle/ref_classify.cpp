// ref_classify.cpp -- TEST INFRASTRUCTURE (parity oracle).  C entry point around the REFERENCE'S OWN classifier:
//   exec/HDK_PolyStokesSolver_Classifier.cpp  (classifyCells, the air / solid boundary floods, constructTiles, classifyFaces /
//   classifyEdges, constructCenterReducedIndices with fixReducedRegionBoundaries / fixSmallReducedRegions, the face / edge reduced
//   indices, serialAssignFieldIndices, buildValidFaces)
// compiled UNMODIFIED from /root/reference (oracle/Makefile target `ref`) against oracle/hdk_shim (the HDK stand-in of BASELINE.md
// section 3) and oracle/eigen_facade; class layout from the reference's own exec/HDK_PolyStokesSolver.h / HDK_PolyStokes.h.
// Not compiled: exec/HDK_PolyStokesSolver.cpp (2183 lines of weights / region algebra / dumps that need far more of the HDK and of
// Eigen).  The classifier needs six small members that live in that file; they are written here from their contract, not copied:
// the constructor (fields sized per sample position, filled with UNASSIGNED, constant border UNASSIGNED: S.cpp:18-155), the
// destructor, overwriteIndices / copyMaterialLabel (replace / copy one label value, S.cpp:1925-2005), setActiveLayerCells
// (GENERICFLUID -> ACTIVEFLUID on a cell list, S.cpp:2022-2059) and findOccupiedIndexTiles (mark the tiles a cell list touches,
// S.cpp:2061-2103).  The integration weights (HDK computeSDFWeightsSampled, closed source) are INPUTS here.
// Also compiled unmodified: exec/HDK_PolyStokesSolver_ConstructMatrixBlocks.cpp (constructMatrixBlocks / buildMatrixBlocksByTriplets:
// M_c, M_c^-1, mu, mu^-1, G, D^T, JG, JD^T and the right-hand sides).  Its inputs from exec/HDK_PolyStokesSolver.cpp are passed in by
// the caller: the region centres of mass (computeCenterOfMasses) and the basis evaluation buildConversionCoefficients (see below).
// And exec/HDK_PolyStokesSolver_AssembleBlocks.cpp + _AssembleSystem.cpp (assemble(): M_r, B, B^-1, reduced right-hand side, b, and the
// explicit A of assembleSystemPressureStress); their inputs from the region algebra of exec/HDK_PolyStokesSolver.cpp (per-region mass /
// viscosity matrices, best-fit vectors) are passed in by the caller.
// The call sequence is HDK_PolyStokes::solveGasSubclass, exec/HDK_PolyStokes.C:344-446.
#include <cstring>
#include "hdk_shim.h"
#include <Eigen/Sparse>
#include <tbb/tbb.h>
// the harness reads the classifier's private result fields
#define private public
#define protected public
#include "HDK_PolyStokesSolver.h"
#undef private
#undef protected

typedef HDK_PolyStokes::Solver Solver;

// ---- the members of exec/HDK_PolyStokesSolver.cpp the classifier links against (see the header comment) ----
Solver::Solver(HDK_PolyStokes& _parent, const fpreal _dx, const fpreal _dt, SIM_VectorField* _velocityField, const SIM_VectorField* _collisionVelocityField,
               const SIM_ScalarField* _surfaceField, const SIM_VectorField* _surfaceWeights, const SIM_ScalarField* _collisionField, const SIM_VectorField* _collisionWeights,
               const SIM_ScalarField* _densityField, const fpreal _constantDensity, const SIM_ScalarField* _viscosityField)
    : myParent(_parent), myVelocityField(_velocityField), myCollisionVelocityField(_collisionVelocityField),
      myUseSurfaceWeights(_parent.getUseInputSurfaceWeights()), mySurfaceField(_surfaceField), mySurfaceWeights(_surfaceWeights),
      myUseCollisionWeights(_parent.getUseInputCollisionWeights()), myCollisionField(_collisionField), myCollisionWeights(_collisionWeights),
      mySurfaceFieldData(_surfaceField->getField()), myCollisionFieldData(_collisionField->getField()), myViscosityFieldData(_viscosityField->getField()),
      myDensityFieldData(_densityField->getField()), myConstantDensity(_constantDensity), myMinDensity(_parent.getMinDensity()), myMaxDensity(_parent.getMaxDensity()),
      myMatrixScheme(_parent.getMatrixScheme()), mySolverType(_parent.getSolverType()), mySolverTolerance(_parent.getSolverTolerance()),
      mySolverMaxIterations(_parent.getSolverMaxIterations()), dx(_dx), invDx(1. / _dx), invDx2(1. / (_dx * _dx)), dt(_dt), invDt(1. / _dt),
      size(_velocityField->getSize()), orig(_velocityField->getOrig()), resolution(_velocityField->getTotalVoxelRes()),
      nx((exint)resolution.x()), ny((exint)resolution.y()), nz((exint)resolution.z())
{
    myLiquidBoundaryLayerSize = _parent.getActiveLiquidBoundaryLayerSize();
    mySolidBoundaryLayerSize = _parent.getActiveSolidBoundaryLayerSize();
    myDoReducedRegions = _parent.getDoReducedRegions();
    myDoTile = _parent.getDoTile();
    myTileSize = _parent.getTileSize();
    myTilePadding = _parent.getTilePadding();
    const SIM_RawField& ref = *mySurfaceFieldData;
    int res[3]; ref.getVoxelRes(res[0], res[1], res[2]);
    struct Slot { SIM_FieldSample sample; SIM_RawField* liquid; SIM_RawField* fluid; SIM_RawIndexField* index[3]; };
    const Slot slots[7] = {
        {SIM_SAMPLE_CENTER, &centerLiquidWeights, &centerFluidWeights, {&centerLabels, &centerReducedIndices, &centerActiveIndices}},
        {SIM_SAMPLE_EDGEXY, &edgeXYLiquidWeights, &edgeXYFluidWeights, {&edgeXYLabels, &edgeXYReducedIndices, &edgeXYActiveIndices}},
        {SIM_SAMPLE_EDGEXZ, &edgeXZLiquidWeights, &edgeXZFluidWeights, {&edgeXZLabels, &edgeXZReducedIndices, &edgeXZActiveIndices}},
        {SIM_SAMPLE_EDGEYZ, &edgeYZLiquidWeights, &edgeYZFluidWeights, {&edgeYZLabels, &edgeYZReducedIndices, &edgeYZActiveIndices}},
        {SIM_SAMPLE_FACEX, &faceXLiquidWeights, &faceXFluidWeights, {&faceXLabels, &faceXReducedIndices, &faceXActiveIndices}},
        {SIM_SAMPLE_FACEY, &faceYLiquidWeights, &faceYFluidWeights, {&faceYLabels, &faceYReducedIndices, &faceYActiveIndices}},
        {SIM_SAMPLE_FACEZ, &faceZLiquidWeights, &faceZFluidWeights, {&faceZLabels, &faceZReducedIndices, &faceZActiveIndices}}};
    for (const Slot& s : slots) {
        for (SIM_RawField* w : {s.liquid, s.fluid}) { w->init(s.sample, ref.getOrig(), ref.getSize(), res[0], res[1], res[2]); w->makeConstant(0.f); }
        for (SIM_RawIndexField* f : s.index) {
            f->init(s.sample, ref.getOrig(), ref.getSize(), res[0], res[1], res[2]);
            f->makeConstant(MaterialLabels::UNASSIGNED);
            f->setBorder(UT_VOXELBORDER_CONSTANT, MaterialLabels::UNASSIGNED);
        }
    }
    myThreadCount = UT_Thread::getNumProcessors();
}
Solver::~Solver() {}

void Solver::overwriteIndices(SIM_RawIndexField& indices, const exint searchValue, const exint replaceValue) {
    UT_VoxelArrayI& a = *indices.fieldNC();
    for (size_t i = 0; i < a.d.size(); ++i) if (a.d[i] == searchValue) a.d[i] = replaceValue;
    a.expandAllTiles();
}
void Solver::copyMaterialLabel(SIM_RawIndexField& source, SIM_RawIndexField& dest, const exint searchValue) {
    const UT_VoxelArrayI& s = *source.field(); UT_VoxelArrayI& d = *dest.fieldNC();
    for (size_t i = 0; i < s.d.size(); ++i) if (s.d[i] == searchValue) d.d[i] = searchValue;
    d.expandAllTiles();
}
void Solver::setActiveLayerCells(const UT_Array<UT_Vector3I>& activeCellLayer) {
    for (exint i = 0; i < activeCellLayer.size(); ++i) {
        const UT_Vector3I c = activeCellLayer[i];
        if (SIM::FieldUtils::getFieldValue(centerLabels, c) == MaterialLabels::GENERICFLUID) SIM::FieldUtils::setFieldValue(centerLabels, c, MaterialLabels::ACTIVEFLUID);
    }
}
void Solver::findOccupiedIndexTiles(UT_Array<bool>& isTileOccupiedList, const UT_Array<UT_Vector3I>& indexCellList, const SIM_RawIndexField& indexCellLabels) {
    for (exint i = 0; i < indexCellList.size(); ++i) {
        const UT_Vector3I c = indexCellList[i];
        isTileOccupiedList[indexCellLabels.field()->indexToLinearTile((int)c[0], (int)c[1], (int)c[2])] = true;
    }
}

// getLocalDensity / getLocalViscosity (S.cpp:1914-1924) are one-line accessors: the constant density and the trilinear viscosity sample.
fpreal Solver::getLocalDensity(UT_Vector3I) { return myConstantDensity; }
fpreal Solver::getLocalViscosity(UT_Vector3 point) { return myViscosityFieldData->getValue(point); }
// buildConversionCoefficients (S.cpp:2107-2149, the 26 basis functions evaluated at a face offset) is NOT retyped here: the caller
// passes a function (tests: the oracle's orc_conversion_coefficients, which has its own analytic known-answer tests).
typedef void (*refcls_coeff_fn)(const double* offset3, int axis, double* out26);
static refcls_coeff_fn g_coeff = nullptr;
ColumnVector Solver::buildConversionCoefficients(UT_Vector3T<SolveReal> offset, int axis) {
    ColumnVector v;
    double o[3] = {offset[0], offset[1], offset[2]}, c[REDUCED_DOF];
    g_coeff(o, axis, c);
    for (int i = 0; i < REDUCED_DOF; ++i) v(i) = c[i];
    return v;
}

// ---- the node class: declared by the reference's HDK_PolyStokes.h, its bodies live in HDK_PolyStokes.C (HDK node plumbing) ----
HDK_PolyStokes::HDK_PolyStokes(const SIM_DataFactory* factory) : GAS_SubSolver(factory) {}
HDK_PolyStokes::~HDK_PolyStokes() {}
bool HDK_PolyStokes::solveGasSubclass(SIM_Engine&, SIM_Object*, SIM_Time, SIM_Time) { return false; }
const SIM_DopDescription* HDK_PolyStokes::getDopDescription() { return nullptr; }
namespace { struct Node : HDK_PolyStokes { Node() : HDK_PolyStokes(nullptr) {} }; }

extern "C" {

struct refcls_params {
    int32_t nx, ny, nz;
    double dx, dt;
    int32_t liquidLayers, solidLayers, doReducedRegions, doTile, tileSize, tilePadding;
    double density;
    int32_t solverType;        // units.h:76-94: 0 PCG_MATRIX_VECTOR_PRODUCTS (factored), 1 EIGEN (explicit A)
};

struct RefCls {
    SIM_VectorField velocity, collisionVelocity, validFaces;
    SIM_ScalarField surface, collision, density, viscosity;
    Node node;
    Solver* S = nullptr;
    ~RefCls() { delete S; }
};

static SIM_RawIndexField* index_field(Solver& S, int kind, int slot) {
    SIM_RawIndexField* fields[3][7] = {
        {&S.centerLabels, &S.faceXLabels, &S.faceYLabels, &S.faceZLabels, &S.edgeYZLabels, &S.edgeXZLabels, &S.edgeXYLabels},
        {&S.centerActiveIndices, &S.faceXActiveIndices, &S.faceYActiveIndices, &S.faceZActiveIndices, &S.edgeYZActiveIndices, &S.edgeXZActiveIndices, &S.edgeXYActiveIndices},
        {&S.centerReducedIndices, &S.faceXReducedIndices, &S.faceYReducedIndices, &S.faceZReducedIndices, &S.edgeYZReducedIndices, &S.edgeXZReducedIndices, &S.edgeXYReducedIndices}};
    return fields[kind][slot];
}

// weights: 14 float arrays in the repository's slot order (0 centre, 1..3 faces x / y / z, 4..6 edges of axis 0 (YZ) / 1 (XZ) / 2 (XY)),
// liquid first then fluid, each x-fastest with the slot's resolution.  Runs the classification; returns a handle.
void* refcls_create(const refcls_params* P, const float* const* weights) {
    std::map<std::string, double>& prm = hdk_shim::params();
    prm.clear();
    prm["matrixSetup"] = 0; prm["solverType"] = P->solverType; prm["useInputSurfaceWeights"] = 0; prm["useInputCollisionWeights"] = 0;
    prm["minDensity"] = 0; prm["maxDensity"] = 1e30; prm["activeLiquidBoundaryLayerSize"] = P->liquidLayers; prm["activeSolidBoundaryLayerSize"] = P->solidLayers;
    prm["doReducedRegions"] = P->doReducedRegions; prm["doTile"] = P->doTile; prm["tileSize"] = P->tileSize; prm["tilePadding"] = P->tilePadding;
    prm[SIM_NAME_TOLERANCE] = 1e-3; prm["maxSolverIterations"] = 1; prm["useWarmStart"] = 0; prm["exportMatrices"] = 0; prm["exportComponentMatrices"] = 0;
    prm["exportStats"] = 0; prm["doSolve"] = 0; prm["keepNonConvergedResults"] = 0;

    RefCls* H = new RefCls;
    const UT_Vector3 orig(0.f, 0.f, 0.f), size((float)(P->nx * P->dx), (float)(P->ny * P->dx), (float)(P->nz * P->dx));
    const SIM_FieldSample faceSample[3] = {SIM_SAMPLE_FACEX, SIM_SAMPLE_FACEY, SIM_SAMPLE_FACEZ};
    for (int a = 0; a < 3; ++a)
        for (SIM_VectorField* v : {&H->velocity, &H->collisionVelocity, &H->validFaces}) v->getField(a)->init(faceSample[a], orig, size, P->nx, P->ny, P->nz);
    for (SIM_ScalarField* s : {&H->surface, &H->collision, &H->density, &H->viscosity}) s->getField()->init(SIM_SAMPLE_CENTER, orig, size, P->nx, P->ny, P->nz);

    H->S = new Solver(H->node, P->dx, P->dt, &H->velocity, &H->collisionVelocity, &H->surface, nullptr, &H->collision, nullptr, &H->density, P->density, &H->viscosity);
    Solver& S = *H->S;

    // buildIntegrationWeightsAlt (S.cpp:238-287) with the weights supplied by the caller
    SIM_RawField* liquid[7] = {&S.centerLiquidWeights, &S.faceXLiquidWeights, &S.faceYLiquidWeights, &S.faceZLiquidWeights, &S.edgeYZLiquidWeights, &S.edgeXZLiquidWeights, &S.edgeXYLiquidWeights};
    SIM_RawField* fluid[7] = {&S.centerFluidWeights, &S.faceXFluidWeights, &S.faceYFluidWeights, &S.faceZFluidWeights, &S.edgeYZFluidWeights, &S.edgeXZFluidWeights, &S.edgeXYFluidWeights};
    for (int slot = 0; slot < 7; ++slot)
        for (int k = 0; k < 2; ++k) {
            UT_VoxelArrayF& a = *(k == 0 ? liquid : fluid)[slot]->fieldNC();
            memcpy(a.d.data(), weights[k * 7 + slot], a.d.size() * sizeof(float));
            a.expandAllTiles();
        }
    S.liquidWeights = {&S.centerLiquidWeights, &S.edgeXYLiquidWeights, &S.edgeXZLiquidWeights, &S.edgeYZLiquidWeights, &S.faceXLiquidWeights, &S.faceYLiquidWeights, &S.faceZLiquidWeights};
    S.fluidWeights = {&S.centerFluidWeights, &S.edgeXYFluidWeights, &S.edgeXZFluidWeights, &S.edgeYZFluidWeights, &S.faceXFluidWeights, &S.faceYFluidWeights, &S.faceZFluidWeights};

    // exec/HDK_PolyStokes.C:344-395
    S.classifyCells();
    if (S.doReducedRegions()) S.constructReducedRegions(); else S.constructOnlyActiveRegions();
    S.classifyFaces();
    S.classifyEdges();
    if (S.doReducedRegions()) {
        S.constructCenterReducedIndices();
        S.constructFacesReducedIndices();
        S.constructEdgesReducedIndices();
    }
    S.constructCenterActiveIndices();
    S.constructFacesActiveIndices();
    S.constructEdgesActiveIndices();
    return H;
}
void refcls_destroy(void* hv) { delete (RefCls*)hv; }

// out: 21 int64 arrays, [kind][slot] with kind 0 labels, 1 active indices, 2 reduced indices.  counts[8]: nCenter, nFaceX, nFaceY, nFaceZ,
// nEdgeYZ, nEdgeXZ, nEdgeXY, regionCount.  valid (may be NULL): 3 float arrays (face x / y / z) written by buildValidFaces.
void refcls_results(void* hv, int64_t* const* out, int64_t* counts, float* const* valid) {
    RefCls* H = (RefCls*)hv; Solver& S = *H->S;
    for (int kind = 0; kind < 3; ++kind)
        for (int slot = 0; slot < 7; ++slot) {
            const UT_VoxelArrayI& a = *index_field(S, kind, slot)->field();
            for (size_t i = 0; i < a.d.size(); ++i) out[kind * 7 + slot][i] = (int64_t)a.d[i];
        }
    counts[0] = S.nCenter; counts[1] = S.nFaceX; counts[2] = S.nFaceY; counts[3] = S.nFaceZ; counts[4] = S.nEdgeYZ; counts[5] = S.nEdgeXZ; counts[6] = S.nEdgeXY;
    counts[7] = S.myInteriorRegionCount;
    if (valid) {
        S.buildValidFaces(H->validFaces);      // exec/HDK_PolyStokes.C:562-570
        for (int a = 0; a < 3; ++a) { const UT_VoxelArrayF& v = *H->validFaces.getField(a)->field(); memcpy(valid[a], v.d.data(), v.d.size() * sizeof(float)); }
    }
}

// constructMatrixBlocks (exec/HDK_PolyStokes.C:432-435) on the classified grid.  vel / colvel: 3 face-sampled float arrays each,
// viscosity: centre-sampled float array, com: regionCount x 3 doubles (computeCenterOfMasses), coeff: see buildConversionCoefficients above.
int refcls_construct_blocks(void* hv, const float* const* vel, const float* const* colvel, const float* viscosity, const double* com, refcls_coeff_fn coeff) {
    RefCls* H = (RefCls*)hv; Solver& S = *H->S;
    for (int a = 0; a < 3; ++a) {
        UT_VoxelArrayF& v = *H->velocity.getField(a)->fieldNC(); memcpy(v.d.data(), vel[a], v.d.size() * sizeof(float)); v.expandAllTiles();
        UT_VoxelArrayF& c = *H->collisionVelocity.getField(a)->fieldNC(); memcpy(c.d.data(), colvel[a], c.d.size() * sizeof(float)); c.expandAllTiles();
    }
    UT_VoxelArrayF& mu = *H->viscosity.getField()->fieldNC(); memcpy(mu.d.data(), viscosity, mu.d.size() * sizeof(float)); mu.expandAllTiles();
    S.reducedRegionCOM.setSize(S.myInteriorRegionCount);
    for (exint r = 0; r < S.myInteriorRegionCount; ++r) S.reducedRegionCOM[r] = UT_Vector3T<SolveReal>(com[3 * r], com[3 * r + 1], com[3 * r + 2]);
    g_coeff = coeff;
    S.constructMatrixBlocks();
    return 0;
}

// assemble() (exec/HDK_PolyStokes.C:436-446 with initializeGuessVectors, S.cpp:512-519) from exec/HDK_PolyStokesSolver_AssembleSystem.cpp and
// _AssembleBlocks.cpp: M_r, B = M_r / dt + 2 JD^T mu D J^T, B^-1, the reduced right-hand side, b, and for solverType EIGEN the explicit A.
// Inputs that exec/HDK_PolyStokesSolver.cpp would have computed (region algebra D3-D5, not compiled): per region the 26 x 26 mass and
// viscosity matrices (row-major) and the 26 best-fit coefficients.
int refcls_assemble(void* hv, const double* massDense, const double* viscDense, const double* bestFit) {
    RefCls* H = (RefCls*)hv; Solver& S = *H->S;
    const exint R = S.myInteriorRegionCount;
    S.reducedMassMatrices.setSize(R); S.reducedViscosityMatrices.setSize(R); S.reducedRegionBestFitVectors.setSize(R);
    for (exint r = 0; r < R; ++r) {
        for (int i = 0; i < REDUCED_DOF; ++i) {
            for (int j = 0; j < REDUCED_DOF; ++j) {
                S.reducedMassMatrices[r](i, j) = massDense[(r * REDUCED_DOF + i) * REDUCED_DOF + j];
                S.reducedViscosityMatrices[r](i, j) = viscDense[(r * REDUCED_DOF + i) * REDUCED_DOF + j];
            }
            S.reducedRegionBestFitVectors[r](i) = bestFit[r * REDUCED_DOF + i];
        }
    }
    S.activeGuessVector = Vector::Zero(S.nActiveVs); S.reducedGuessVector = Vector::Zero(S.nReducedVs);
    S.pressureGuessVector = Vector::Zero(S.nPressures); S.stressGuessVector = Vector::Zero(S.nStresses);
    S.assemble();
    return 0;
}

static const SparseMatrix* block(Solver& S, const char* name) {
    const std::string n(name);
    if (n == "Mr") return &S.Mr_Matrix; if (n == "B") return &S.Mr_plus_2JDtuDJ_Matrix; if (n == "BInv") return &S.Inv_Mr_plus_2JDtuDJ_Matrix; if (n == "A") return &S.A;
    if (n == "Mc") return &S.Mc_Matrix; if (n == "McInv") return &S.McInv_Matrix; if (n == "u") return &S.u_Matrix; if (n == "uInv") return &S.uInv_Matrix;
    if (n == "G") return &S.G_Matrix; if (n == "Dt") return &S.Dt_Matrix; if (n == "JG") return &S.JG_Matrix; if (n == "JDt") return &S.JDt_Matrix;
    if (n == "oldActiveVs") return &S.oldActiveVs;
    return nullptr;
}
int refcls_csr_dims(void* hv, const char* name, int64_t* rows, int64_t* cols, int64_t* nnz) {
    const SparseMatrix* m = block(*((RefCls*)hv)->S, name);
    if (!m) return -1;
    *rows = m->rows(); *cols = m->cols(); *nnz = m->nonZeros();
    return 0;
}
int refcls_csr_copy(void* hv, const char* name, int64_t* ptr, int32_t* idx, double* val) {
    const SparseMatrix* m = block(*((RefCls*)hv)->S, name);
    if (!m) return -1;
    for (Eigen::Index r = 0; r <= m->outerSize(); ++r) ptr[r] = m->outerIndexPtr()[r];
    for (Eigen::Index k = 0; k < m->nonZeros(); ++k) { idx[k] = m->innerIndexPtr()[k]; val[k] = m->valuePtr()[k]; }
    return 0;
}
int64_t refcls_vector(void* hv, const char* name, double* out) {
    Solver& S = *((RefCls*)hv)->S; const std::string n(name);
    const Vector* v = n == "activeRHS" ? &S.activeRHSVector : n == "pressureRHS" ? &S.pressureRHSVector : n == "stressRHS" ? &S.stressRHSVector :
                      n == "reducedRHS" ? &S.reducedRHSVector : n == "b" ? &S.b : nullptr;
    if (!v) return -1;
    if (out) for (Eigen::Index i = 0; i < v->size(); ++i) out[i] = (*v)(i);
    return (int64_t)v->size();
}

}  // extern "C"
