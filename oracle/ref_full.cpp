// ref_full.cpp -- TEST INFRASTRUCTURE (parity oracle / CPU reference).  C entry points around the REFERENCE'S COMPLETE per-step solver:
// all six exec/HDK_PolyStokesSolver*.cpp translation units (5 789 lines: constructor, integration weights, classifier, centres of mass,
// least-squares fits, reduced mass / viscosity matrices, matrix blocks, assembly, the solvers, velocity recovery and write-back, valid
// faces, exports), lib/src/Preconditioner.cpp and the lib/include headers, compiled UNMODIFIED from /root/reference
// (oracle/Makefile target `ref` -> oracle/_ref/libps_ref_full.so) on oracle/hdk_shim (the HDK stand-in of BASELINE.md section 3) and
// oracle/eigen_facade (the checkout's Eigen lacks Eigen/Core).  No reference source is copied; the class layout comes from the
// reference's own exec/HDK_PolyStokesSolver.h / HDK_PolyStokes.h.  The ONLY reference code not compiled is exec/HDK_PolyStokes.C -- the
// Houdini node (parameter templates, field fetching, addError): its stage sequence, exec/HDK_PolyStokes.C:329-608, is what reffull_step
// below calls, and the four out-of-line members of the node class are defined here as stubs.
// What remains the stand-ins' (not the reference's): tile iteration order, connected components, computeSDFWeightsSampled, trilinear
// getValue, border modes, the job pool behind UT_ThreadedAlgorithm / tbb::parallel_for (hdk_shim.h; one job unless reffull_set_threads), and the per-operation arithmetic of the Eigen calls (eigen_facade/Eigen/Core).
#include <cstring>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "hdk_shim.h"
#include <Eigen/Sparse>
#include <tbb/tbb.h>
// the harness reads the solver's private state
#define private public
#define protected public
#include "HDK_PolyStokesSolver.h"
#undef private
#undef protected

typedef HDK_PolyStokes::Solver Solver;

// the node class: declared by the reference's HDK_PolyStokes.h; its bodies live in HDK_PolyStokes.C (HDK node plumbing, not compiled)
HDK_PolyStokes::HDK_PolyStokes(const SIM_DataFactory* factory) : GAS_SubSolver(factory) {}
HDK_PolyStokes::~HDK_PolyStokes() {}
bool HDK_PolyStokes::solveGasSubclass(SIM_Engine&, SIM_Object*, SIM_Time, SIM_Time) { return false; }
const SIM_DopDescription* HDK_PolyStokes::getDopDescription() { return nullptr; }
namespace { struct Node : HDK_PolyStokes { Node() : HDK_PolyStokes(nullptr) { myObj = nullptr; } }; }

extern "C" {

struct reffull_params {
    int32_t nx, ny, nz;
    double dx, dt, density, tolerance;
    int32_t maxIterations, liquidLayers, solidLayers, doReducedRegions, doTile, tileSize, tilePadding, solverType, useWarmStart, keepNonConvergedResults;
};

struct RefFull {
    SIM_VectorField velocity, collisionVelocity, validFaces;
    SIM_ScalarField surface, collision, density, viscosity;
    Node node;
    Solver* S = nullptr;
    int result = -3;
    ~RefFull() { delete S; }
};

static void fill(SIM_RawField* f, const float* src) { UT_VoxelArrayF& a = *f->fieldNC(); memcpy(a.d.data(), src, a.d.size() * sizeof(float)); a.expandAllTiles(); }

// fields: x-fastest float arrays; surface / collision / viscosity centre-sampled, vel / colvel face-sampled (x, y, z)
void* reffull_create(const reffull_params* P, const float* surface, const float* collision, const float* viscosity, const float* const* vel, const float* const* colvel) {
    std::map<std::string, double>& prm = hdk_shim::params();
    prm.clear();
    prm["matrixSetup"] = 0; prm["solverType"] = P->solverType; prm["useInputSurfaceWeights"] = 0; prm["useInputCollisionWeights"] = 0;
    prm["minDensity"] = 0; prm["maxDensity"] = 1e30; prm["activeLiquidBoundaryLayerSize"] = P->liquidLayers; prm["activeSolidBoundaryLayerSize"] = P->solidLayers;
    prm["doReducedRegions"] = P->doReducedRegions; prm["doTile"] = P->doTile; prm["tileSize"] = P->tileSize; prm["tilePadding"] = P->tilePadding;
    prm[SIM_NAME_TOLERANCE] = P->tolerance; prm["maxSolverIterations"] = P->maxIterations; prm["useWarmStart"] = P->useWarmStart;
    prm["exportMatrices"] = 0; prm["exportComponentMatrices"] = 0; prm["exportStats"] = 0; prm["doSolve"] = 1; prm["keepNonConvergedResults"] = P->keepNonConvergedResults;

    RefFull* H = new RefFull;
    const UT_Vector3 orig(0.f, 0.f, 0.f), size((float)(P->nx * P->dx), (float)(P->ny * P->dx), (float)(P->nz * P->dx));
    const SIM_FieldSample faceSample[3] = {SIM_SAMPLE_FACEX, SIM_SAMPLE_FACEY, SIM_SAMPLE_FACEZ};
    for (int a = 0; a < 3; ++a)
        for (SIM_VectorField* v : {&H->velocity, &H->collisionVelocity, &H->validFaces}) v->getField(a)->init(faceSample[a], orig, size, P->nx, P->ny, P->nz);
    for (SIM_ScalarField* s : {&H->surface, &H->collision, &H->density, &H->viscosity}) s->getField()->init(SIM_SAMPLE_CENTER, orig, size, P->nx, P->ny, P->nz);
    fill(H->surface.getField(), surface); fill(H->collision.getField(), collision); fill(H->viscosity.getField(), viscosity);
    for (int a = 0; a < 3; ++a) { fill(H->velocity.getField(a), vel[a]); fill(H->collisionVelocity.getField(a), colvel[a]); }
    H->S = new Solver(H->node, P->dx, P->dt, &H->velocity, &H->collisionVelocity, &H->surface, nullptr, &H->collision, nullptr, &H->density, P->density, &H->viscosity);
    return H;
}
void reffull_destroy(void* hv) { delete (RefFull*)hv; }
// jobs of the HDK stand-in's thread pool (UT_ThreadedAlgorithm / UTparallelFor* / tbb::parallel_for); 0 = all hardware threads.  Set BEFORE
// reffull_create (the solver reads UT_Thread::getNumProcessors() in its constructor, S.cpp:154).  Returns the job count in effect.
int reffull_set_threads(int n) { hdk_shim::setThreads(n); return hdk_shim::threads(); }
int reffull_get_threads() { return hdk_shim::threads(); }

// exec/HDK_PolyStokes.C:344-476: everything up to and including assemble() (the reference's "setup" clock)
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
// PS_REF_TIMING=1: per-stage wall clock of the reference's setup on stderr (diagnostic)
#define STAGE(name, stmt) do { const double t0_ = now_s(); stmt; if (timing) fprintf(stderr, "[ref stage] %-40s %8.3f s\n", name, now_s() - t0_); } while (0)
int reffull_setup(void* hv) {
    RefFull* H = (RefFull*)hv; Solver& S = *H->S; HDK_PolyStokes& node = H->node;
    const bool timing = getenv("PS_REF_TIMING") && atoi(getenv("PS_REF_TIMING"));
    S.setupClockStart();
    STAGE("buildIntegrationWeightsAlt", S.buildIntegrationWeightsAlt());
    STAGE("classifyCells", S.classifyCells());
    STAGE("constructReducedRegions", if (S.doReducedRegions()) S.constructReducedRegions(); else S.constructOnlyActiveRegions());
    STAGE("classifyFaces", S.classifyFaces());
    STAGE("classifyEdges", S.classifyEdges());
    if (S.doReducedRegions()) {
        STAGE("constructCenterReducedIndices", S.constructCenterReducedIndices());
        STAGE("constructFacesReducedIndices", S.constructFacesReducedIndices());
        STAGE("constructEdgesReducedIndices", S.constructEdgesReducedIndices());
    }
    STAGE("construct*ActiveIndices", S.constructCenterActiveIndices(); S.constructFacesActiveIndices(); S.constructEdgesActiveIndices());
    if (S.doReducedRegions()) {
        STAGE("computeCenterOfMasses", S.computeCenterOfMasses());
        STAGE("computeLeastSquaresFits", S.computeLeastSquaresFits());
        STAGE("computeReducedMassMatrices", S.computeReducedMassMatrices());
        if (node.getMatrixScheme() == HDK_PolyStokes_Options::MatrixScheme::ALL_DOFS_EXPLICIT_INTERIOR_STRESS) S.computeReducedViscosityMatrices();
        else STAGE("computeReducedViscosityMatricesInteriorOnly", S.computeReducedViscosityMatricesInteriorOnly());
    }
    STAGE("constructMatrixBlocks", S.constructMatrixBlocks());
    S.initializeGuessVectors();
    if (node.getUseWarmStart()) S.constructGuessVectors();
    STAGE("assemble", S.assemble());
    S.setupClockEnd();
    return 0;
}
// exec/HDK_PolyStokes.C:486-584: preconditioner, solve, valid faces, velocity recovery and write-back.  Returns the SolverResult (S.h:61-70).
int reffull_solve(void* hv) {
    RefFull* H = (RefFull*)hv; Solver& S = *H->S; HDK_PolyStokes& node = H->node;
    S.constructPreconditioner();
    Solver::SolverResult r = S.solve();
    H->result = (int)r;
    if (r == Solver::SolverResult::UNSUPPORTED_SOLVER) return H->result;
    S.buildValidFaces(H->validFaces);
    if (r == Solver::SolverResult::SUCCESS || node.getKeepNonConvergedResults()) {
        if (node.getMatrixScheme() == HDK_PolyStokes_Options::MatrixScheme::PRESSURE_STRESS) S.recoverVelocityFromPressureStress();
        for (int axis : {0, 1, 2}) S.applySolutionToVelocity(*H->velocity.getField(axis), *H->validFaces.getField(axis), axis);
    }
    return H->result;
}

// exec/HDK_PolyStokes.C:508-584 with "Do Solve" off: no solve (solverResult stays INCOMPLETE, the solution vector is the zero vector of
// assemble, S_AS:466), valid faces, and -- only with keepNonConvergedResults -- velocity recovery + write-back of that zero solution
int reffull_skip_solve(void* hv) {
    RefFull* H = (RefFull*)hv; Solver& S = *H->S; HDK_PolyStokes& node = H->node;
    S.constructPreconditioner();
    H->result = (int)Solver::SolverResult::INCOMPLETE;
    S.buildValidFaces(H->validFaces);
    if (node.getKeepNonConvergedResults()) {
        if (node.getMatrixScheme() == HDK_PolyStokes_Options::MatrixScheme::PRESSURE_STRESS) S.recoverVelocityFromPressureStress();
        for (int axis : {0, 1, 2}) S.applySolutionToVelocity(*H->velocity.getField(axis), *H->validFaces.getField(axis), axis);
    }
    return H->result;
}

// ---- read-out ----
static SIM_RawIndexField* index_field(Solver& S, int kind, int slot) {
    SIM_RawIndexField* fields[3][7] = {
        {&S.centerLabels, &S.faceXLabels, &S.faceYLabels, &S.faceZLabels, &S.edgeYZLabels, &S.edgeXZLabels, &S.edgeXYLabels},
        {&S.centerActiveIndices, &S.faceXActiveIndices, &S.faceYActiveIndices, &S.faceZActiveIndices, &S.edgeYZActiveIndices, &S.edgeXZActiveIndices, &S.edgeXYActiveIndices},
        {&S.centerReducedIndices, &S.faceXReducedIndices, &S.faceYReducedIndices, &S.faceZReducedIndices, &S.edgeYZReducedIndices, &S.edgeXZReducedIndices, &S.edgeXYReducedIndices}};
    return fields[kind][slot];
}
// kind 0 labels / 1 active indices / 2 reduced indices; slot 0 centre, 1..3 faces x / y / z, 4..6 edges of axis 0 (YZ) / 1 (XZ) / 2 (XY)
void reffull_index_field(void* hv, int kind, int slot, int64_t* out) {
    const UT_VoxelArrayI& a = *index_field(*((RefFull*)hv)->S, kind, slot)->field();
    for (size_t i = 0; i < a.d.size(); ++i) out[i] = (int64_t)a.d[i];
}
void reffull_weight_field(void* hv, int liquid, int slot, float* out) {
    Solver& S = *((RefFull*)hv)->S;
    SIM_RawField* liquidW[7] = {&S.centerLiquidWeights, &S.faceXLiquidWeights, &S.faceYLiquidWeights, &S.faceZLiquidWeights, &S.edgeYZLiquidWeights, &S.edgeXZLiquidWeights, &S.edgeXYLiquidWeights};
    SIM_RawField* fluidW[7] = {&S.centerFluidWeights, &S.faceXFluidWeights, &S.faceYFluidWeights, &S.faceZFluidWeights, &S.edgeYZFluidWeights, &S.edgeXZFluidWeights, &S.edgeXYFluidWeights};
    const UT_VoxelArrayF& a = *(liquid ? liquidW : fluidW)[slot]->field();
    memcpy(out, a.d.data(), a.d.size() * sizeof(float));
}
// which 0: velocity (after reffull_solve: the written-back field), 1: valid faces
void reffull_face_field(void* hv, int which, int axis, float* out) {
    RefFull* H = (RefFull*)hv;
    const UT_VoxelArrayF& a = *(which == 0 ? H->velocity : H->validFaces).getField(axis)->field();
    memcpy(out, a.d.data(), a.d.size() * sizeof(float));
}
int64_t reffull_count(void* hv, const char* name) {
    RefFull* H = (RefFull*)hv; Solver& S = *H->S; const std::string n(name);
    if (n == "nCenter") return S.nCenter; if (n == "nFaceX") return S.nFaceX; if (n == "nFaceY") return S.nFaceY; if (n == "nFaceZ") return S.nFaceZ;
    if (n == "nEdgeYZ") return S.nEdgeYZ; if (n == "nEdgeXZ") return S.nEdgeXZ; if (n == "nEdgeXY") return S.nEdgeXY;
    if (n == "nActiveVs") return S.nActiveVs; if (n == "nReducedVs") return S.nReducedVs; if (n == "nPressures") return S.nPressures; if (n == "nStresses") return S.nStresses;
    if (n == "nTotalDOFs") return S.nTotalDOFs; if (n == "nSystemSize") return S.nSystemSize; if (n == "regionCount") return S.myInteriorRegionCount;
    if (n == "iterations") return S.solveIterations; if (n == "result") return H->result;
    return INT64_MIN;
}
double reffull_real(void* hv, const char* name) {
    Solver& S = *((RefFull*)hv)->S; const std::string n(name);
    if (n == "solveError") return S.solveError; if (n == "setupWallclockMs") return S.setupWallclockTime; if (n == "solveWallclockMs") return S.solveWallclockTime;
    return NAN;
}
static const SparseMatrix* block(Solver& S, const std::string& n) {
    if (n == "Mc") return &S.Mc_Matrix; if (n == "McInv") return &S.McInv_Matrix; if (n == "u") return &S.u_Matrix; if (n == "uInv") return &S.uInv_Matrix;
    if (n == "G") return &S.G_Matrix; if (n == "Dt") return &S.Dt_Matrix; if (n == "JG") return &S.JG_Matrix; if (n == "JDt") return &S.JDt_Matrix;
    if (n == "Mr") return &S.Mr_Matrix; if (n == "B") return &S.Mr_plus_2JDtuDJ_Matrix; if (n == "BInv") return &S.Inv_Mr_plus_2JDtuDJ_Matrix; if (n == "A") return &S.A;
    return nullptr;
}
int reffull_csr_dims(void* hv, const char* name, int64_t* rows, int64_t* cols, int64_t* nnz) {
    const SparseMatrix* m = block(*((RefFull*)hv)->S, name);
    if (!m) return -1;
    *rows = m->rows(); *cols = m->cols(); *nnz = m->nonZeros();
    return 0;
}
int reffull_csr_copy(void* hv, const char* name, int64_t* ptr, int32_t* idx, double* val) {
    const SparseMatrix* m = block(*((RefFull*)hv)->S, name);
    if (!m) return -1;
    for (Eigen::Index r = 0; r <= m->outerSize(); ++r) ptr[r] = m->outerIndexPtr()[r];
    for (Eigen::Index k = 0; k < m->nonZeros(); ++k) { idx[k] = m->innerIndexPtr()[k]; val[k] = m->valuePtr()[k]; }
    return 0;
}
// vectors by the oracle's names; the dense per-region blocks are row-major 26 x 26 (MrDense, ViscDense) / 26 (bestFit) / 3 (com)
int64_t reffull_vector(void* hv, const char* name, double* out) {
    Solver& S = *((RefFull*)hv)->S; const std::string n(name);
    const exint R = S.myInteriorRegionCount;
    if (n == "com") { if (out) for (exint r = 0; r < (exint)S.reducedRegionCOM.size(); ++r) for (int a = 0; a < 3; ++a) out[3 * r + a] = S.reducedRegionCOM[r][a]; return 3 * (int64_t)S.reducedRegionCOM.size(); }
    if (n == "bestFit") { if (out) for (exint r = 0; r < (exint)S.reducedRegionBestFitVectors.size(); ++r) for (int i = 0; i < REDUCED_DOF; ++i) out[r * REDUCED_DOF + i] = S.reducedRegionBestFitVectors[r](i); return REDUCED_DOF * (int64_t)S.reducedRegionBestFitVectors.size(); }
    if (n == "MrDense" || n == "ViscDense") {
        UT_Array<ReducedMatrix>& M = n == "MrDense" ? S.reducedMassMatrices : S.reducedViscosityMatrices;
        if (out) for (exint r = 0; r < (exint)M.size(); ++r) for (int i = 0; i < REDUCED_DOF; ++i) for (int j = 0; j < REDUCED_DOF; ++j) out[(r * REDUCED_DOF + i) * REDUCED_DOF + j] = M[r](i, j);
        return REDUCED_DOF * REDUCED_DOF * (int64_t)M.size();
    }
    (void)R;
    const Vector* v = n == "activeRHS" ? &S.activeRHSVector : n == "pressureRHS" ? &S.pressureRHSVector : n == "stressRHS" ? &S.stressRHSVector :
                      n == "reducedRHS" ? &S.reducedRHSVector : n == "b" ? &S.b : n == "solution" ? &S.solutionVector : n == "guess" ? &S.guessVector : nullptr;
    if (!v) return -1;
    if (out) for (Eigen::Index i = 0; i < v->size(); ++i) out[i] = (*v)(i);
    return (int64_t)v->size();
}
// exportMatrices / exportComponentMatrices / exportMatricesPostSolve / exportStats (S.cpp:533-606); what: bit 0 matrices (+ post-solve), 1 component matrices, 2 stats
void reffull_export(void* hv, const char* prefix, int what) {
    Solver& S = *((RefFull*)hv)->S;
    if (what & 1) { S.exportMatrices(prefix); S.exportMatricesPostSolve(prefix); }
    if (what & 2) S.exportComponentMatrices(prefix);
    if (what & 4) S.exportStats(prefix);
}

}  // extern "C"
