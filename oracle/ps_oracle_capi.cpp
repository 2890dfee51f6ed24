// ps_oracle_capi.cpp -- TEST INFRASTRUCTURE ONLY (see ps_oracle.hpp header).
// Flat C entry points so tests / bench.py's cpu_baseline leg can drive the oracle via ctypes.
#include "ps_oracle.hpp"
#include <cstring>
#include <omp.h>

using namespace orc;

extern "C" {

struct orc_params {
    int nx, ny, nz;
    double dx, dt, density, tolerance;
    int maxIterations, liquidLayers, solidLayers;
    int doReduced, doTile, tileSize, tilePadding;
    int threads;
};

void* orc_create(const orc_params* p) {
    Params q;
    q.nx = p->nx; q.ny = p->ny; q.nz = p->nz; q.dx = p->dx; q.dt = p->dt; q.density = p->density; q.tolerance = p->tolerance;
    q.maxIterations = p->maxIterations; q.liquidLayers = p->liquidLayers; q.solidLayers = p->solidLayers;
    q.doReduced = p->doReduced; q.doTile = p->doTile; q.tileSize = p->tileSize; q.tilePadding = p->tilePadding; q.threads = p->threads;
    if (q.threads > 0) omp_set_num_threads(q.threads);
    return new Oracle(q);
}
void orc_destroy(void* h) { delete (Oracle*)h; }
int orc_num_threads() { return omp_get_max_threads(); }

void orc_set_inputs(void* h, const float* surf, const float* col, const float* visc,
                    const float* vx, const float* vy, const float* vz, const float* cx, const float* cy, const float* cz) {
    const float* v[3] = {vx, vy, vz}; const float* c[3] = {cx, cy, cz};
    ((Oracle*)h)->setInputs(surf, col, visc, v, c);
}
// overrides the 14 weight fields (slot order: centre, faceX..Z, edgeYZ, edgeXZ, edgeXY); then run
// orc_setup_from_weights instead of orc_setup
void orc_set_weights(void* h, int liquid, int slot, const float* w) {
    Oracle* o = (Oracle*)h; Field<float>& F = liquid ? o->liquidW[slot] : o->fluidW[slot];
    std::copy(w, w + F.size(), F.d.begin());
}
void orc_setup(void* h) { ((Oracle*)h)->runSetup(); }
void orc_build_weights(void* h) { ((Oracle*)h)->buildIntegrationWeightsAlt(); }
// individual classifier stages for fine-grained tests
void orc_classify(void* h) {
    Oracle* o = (Oracle*)h;
    o->classifyCells();
    if (o->P.doReduced) o->constructReducedRegions(); else o->constructOnlyActiveRegions();
    o->classifyFaces(); o->classifyEdges();
    if (o->P.doReduced) { o->constructCenterReducedIndices(); o->constructFacesReducedIndices(); o->constructEdgesReducedIndices(); }
    o->constructActiveIndices();
}
void orc_assemble_explicit_A(void* h) { ((Oracle*)h)->assembleExplicitA(); }
void orc_construct_guess(void* h) { ((Oracle*)h)->constructGuessVectors(); }
int orc_solve_eigen_cg(void* h) { Oracle* o = (Oracle*)h; if (o->A.rows != o->nSystemSize || o->A.nnz() == 0) o->assembleExplicitA(); return o->solveEigenCG(); }
int orc_solve(void* h) { return ((Oracle*)h)->solveSPDwithMatrixVectorPCG(); }
void orc_apply(void* h, const double* x, double* y) { Oracle* o = (Oracle*)h; if (o->McInvG.rows != o->nActiveVs || o->McInvG.nnz() != o->G.nnz()) o->setupMatrixVectorProducts(); o->applyMatrixVectorProducts(x, y); }
void orc_writeback(void* h, float* vx, float* vy, float* vz, float* valx, float* valy, float* valz) {
    Oracle* o = (Oracle*)h;
    float* valid[3] = {valx, valy, valz}; float* v[3] = {vx, vy, vz};
    o->buildValidFaces(valid);
    if (o->solverResult == SUCCESS || o->P.keepNonConverged) {
        o->recoverVelocityFromPressureStress();
        for (int a = 0; a < 3; ++a) o->applySolutionToVelocity(v[a], valid[a], a);
    }
}

int64_t orc_count(void* h, const char* name) {
    Oracle* o = (Oracle*)h; std::string n(name);
    if (n == "nCenter") return o->nCenter;
    if (n == "nFaceX") return o->nFace[0]; if (n == "nFaceY") return o->nFace[1]; if (n == "nFaceZ") return o->nFace[2];
    if (n == "nEdgeYZ") return o->nEdge[0]; if (n == "nEdgeXZ") return o->nEdge[1]; if (n == "nEdgeXY") return o->nEdge[2];
    if (n == "nActiveVs") return o->nActiveVs; if (n == "nReducedVs") return o->nReducedVs;
    if (n == "nPressures") return o->nPressures; if (n == "nStresses") return o->nStresses;
    if (n == "nTotalDOFs") return o->nTotalDOFs; if (n == "nSystemSize") return o->nSystemSize;
    if (n == "regionCount") return o->regionCount;
    if (n == "iterations") return o->solveIterations; if (n == "result") return o->solverResult;
    if (n == "usedBiCGStab") return o->usedBiCGStab; if (n == "fixLoops") return o->fixLoops; if (n == "fixRemoved") return o->fixRemoved;
    return INT64_MIN;
}
double orc_real(void* h, const char* name) {
    Oracle* o = (Oracle*)h; std::string n(name);
    if (n == "solveError") return o->solveError; if (n == "setupMs") return o->setupMs; if (n == "solveMs") return o->solveMs;
    return NAN;
}

// kind: 0 labels, 1 activeIdx, 2 reducedIdx (int64 out); slot = sample slot
int64_t orc_index_field(void* h, int kind, int slot, int64_t* out) {
    Oracle* o = (Oracle*)h;
    const Field<exint>& F = kind == 0 ? o->labels[slot] : kind == 1 ? o->activeIdx[slot] : o->reducedIdx[slot];
    if (out) std::copy(F.d.begin(), F.d.end(), out);
    return (int64_t)F.size();
}
int64_t orc_weight_field(void* h, int liquid, int slot, float* out) {
    Oracle* o = (Oracle*)h; const Field<float>& F = liquid ? o->liquidW[slot] : o->fluidW[slot];
    if (out) std::copy(F.d.begin(), F.d.end(), out);
    return (int64_t)F.size();
}

static Csr* findCsr(Oracle* o, const std::string& n) {
    if (n == "Mc") return &o->Mc; if (n == "McInv") return &o->McInv; if (n == "uInv") return &o->uInv; if (n == "u") return &o->uMat;
    if (n == "G") return &o->G; if (n == "Dt") return &o->Dt; if (n == "JG") return &o->JG; if (n == "JDt") return &o->JDt;
    if (n == "Mr") return &o->MrMat; if (n == "B") return &o->Bmat; if (n == "BInv") return &o->BinvMat; if (n == "A") return &o->A;
    return nullptr;
}
int orc_csr_dims(void* h, const char* name, int64_t* rows, int64_t* cols, int64_t* nnz) {
    Csr* m = findCsr((Oracle*)h, name); if (!m) return -1;
    *rows = m->rows; *cols = m->cols; *nnz = m->nnz(); return 0;
}
int orc_csr_copy(void* h, const char* name, int64_t* ptr, int32_t* idx, double* val) {
    Csr* m = findCsr((Oracle*)h, name); if (!m) return -1;
    std::copy(m->ptr.begin(), m->ptr.end(), ptr); std::copy(m->idx.begin(), m->idx.end(), idx); std::copy(m->val.begin(), m->val.end(), val); return 0;
}
int orc_csr_save(void* h, const char* name, const char* path) {
    Csr* m = findCsr((Oracle*)h, name); if (!m) return -1;
    return saveMarket(*m, path) ? 0 : -2;
}

static std::vector<Real>* findVec(Oracle* o, const std::string& n) {
    if (n == "activeRHS") return &o->activeRHS; if (n == "reducedRHS") return &o->reducedRHS;
    if (n == "pressureRHS") return &o->pressureRHS; if (n == "stressRHS") return &o->stressRHS;
    if (n == "b") return &o->b; if (n == "solution") return &o->solution; if (n == "velSolution") return &o->velSolution;
    if (n == "oldActiveVs") return &o->oldActiveVs; if (n == "guess") return &o->guess;
    return nullptr;
}
int64_t orc_vector(void* h, const char* name, double* out) {
    Oracle* o = (Oracle*)h; std::string n(name);
    if (n == "com") { if (out) for (size_t r = 0; r < o->com.size(); ++r) for (int a = 0; a < 3; ++a) out[3 * r + a] = o->com[r][a]; return (int64_t)o->com.size() * 3; }
    if (n == "bestFit") { if (out) for (size_t r = 0; r < o->bestFit.size(); ++r) std::copy(o->bestFit[r].begin(), o->bestFit[r].end(), out + RDOF * r); return (int64_t)o->bestFit.size() * RDOF; }
    if (n == "MrDense" || n == "ViscDense" || n == "BinvDense") {
        auto& V = n == "MrDense" ? o->Mr : n == "ViscDense" ? o->Visc : o->Binv;
        if (out) for (size_t r = 0; r < V.size(); ++r) std::copy(V[r].begin(), V[r].end(), out + (size_t)RDOF * RDOF * r);
        return (int64_t)V.size() * RDOF * RDOF;
    }
    std::vector<Real>* v = findVec(o, n); if (!v) return -1;
    if (out) std::copy(v->begin(), v->end(), out);
    return (int64_t)v->size();
}
int orc_vector_save(void* h, const char* name, const char* path) {
    std::vector<Real>* v = findVec((Oracle*)h, name); if (!v) return -1;
    return saveMarketVector(*v, path) ? 0 : -2;
}

// dense helpers exposed for known-answer tests
void orc_inverse_partial_piv_lu(const double* A, double* Ainv, int n) { inversePartialPivLU(A, Ainv, n); }
int orc_solve_full_piv_lu(const double* A, const double* rhs, double* x, int n) { int rank = 0; solveFullPivLU(A, rhs, x, n, &rank); return rank; }
void orc_conversion_coefficients(const double* off, int axis, double* out) { Params p; p.nx = p.ny = p.nz = 1; Oracle o(p); o.buildConversionCoefficients(off, axis, out); }

}  // extern "C"
