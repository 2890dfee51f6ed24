// ps_oracle.hpp -- TEST INFRASTRUCTURE ONLY (parity oracle + CPU baseline).
//
// CPU restatement of the per-step viscous Stokes solve of panuelosj/polystokes
// (classify -> reduced-region algebra -> assemble -> PCG -> write-back), written
// against the HDK shim documented in BASELINE.md section 3.  Every function cites the
// reference file:line it follows (paths relative to the reference checkout:
// exec/HDK_PolyStokesSolver*.cpp = S*.cpp, lib/include/*.h).
//
// PINNING.  The reference ships no tests / golden vectors, but its WHOLE solver is compiled here from its own files
// (oracle/_ref/libps_ref_full.so, `make ref`: all six exec/HDK_PolyStokesSolver*.cpp + lib/, unmodified, on the stand-ins
// oracle/hdk_shim for the closed-source HDK and oracle/eigen_facade for the incomplete Eigen checkout), and this restatement is
// checked against it stage by stage (tests/test_ref_full.py, test_ref_classify.py, test_ref_solve.py): weights, classification,
// centres of mass, least-squares fits, region matrices, matrix blocks, B / B^-1, right-hand sides BIT-EQUAL; b 2e-15; same CG /
// BiCGSTAB / Eigen-CG iteration counts; velocity within the solver tolerance.
// Pinned MODULO the stand-ins: the HDK primitives (tile iteration order, connected components, computeSDFWeightsSampled,
// trilinear getValue, border modes -- defined in BASELINE.md section 3, implemented once in hdk_shim.h and once here) and the
// per-operation rounding of the Eigen calls.  Against a real Houdini build only those could differ.
// Nothing in the product path (polystokes_b200/) may include, link or call this code.
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>
#include <string>
#include <array>
#include <algorithm>
#include <cmath>
#include <cassert>

namespace orc {

using exint = int64_t;
using Real = double;

// S.h:71-82
enum Label : exint {
    UNASSIGNED = -1, UNSOLVED = -2, GENERICFLUID = -3, ACTIVEFLUID = -4, SOLID = -5,
    REDUCED = -6, UNVISITED = -7, VISITED = -8, BOUNDARY = -9
};
// S.h:61-70
enum SolverResult : int {
    UNSUPPORTED_SOLVER = -4, INCOMPLETE = -3, INVALID = -2, FAILED = -1,
    NOCONVERGE = 0, SUCCESS = 1, NOCHANGE = 2
};

// sample slots used by this oracle: 0 centre, 1..3 face x/y/z, 4..6 edge axis 0/1/2
// (edge axis 0 = YZ, 1 = XZ, 2 = XY; S.h:397-411, 541-554)
enum Samp { S_CENTER = 0, S_FACEX = 1, S_FACEY = 2, S_FACEZ = 3, S_EDGEYZ = 4, S_EDGEXZ = 5, S_EDGEXY = 6 };

constexpr int RDOF = 26;  // units.h:10-15 (QUADRATIC_REGIONS)

struct I3 {
    int v[3];
    int& operator[](int a) { return v[a]; }
    int operator[](int a) const { return v[a]; }
    bool operator==(const I3& o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2]; }
};

// SIM::FieldUtils index maps (HDK, shim: BASELINE.md section 3)
inline I3 cellToFaceMap(I3 c, int axis, int dir) { if (dir == 1) c[axis] += 1; return c; }
inline I3 faceToCellMap(I3 f, int axis, int dir) { if (dir == 0) f[axis] -= 1; return f; }
inline I3 cellToCellMap(I3 c, int axis, int dir) { c[axis] += (dir == 0) ? -1 : 1; return c; }
inline I3 faceToEdgeMap(I3 f, int faceAxis, int edgeAxis, int dir) { if (dir == 1) f[3 - faceAxis - edgeAxis] += 1; return f; }
inline I3 edgeToFaceMap(I3 e, int edgeAxis, int faceAxis, int dir) { if (dir == 0) e[3 - faceAxis - edgeAxis] -= 1; return e; }

// dense x-fastest stand-in for SIM_RawField / SIM_RawIndexField
template <typename T>
struct Field {
    int r[3] = {0, 0, 0};
    std::vector<T> d;
    T border = T(-1);      // index fields: UT_VOXELBORDER_CONSTANT, UNASSIGNED (S.cpp:101-152)
    bool clampBorder = false;  // float fields: clamp-to-edge (shim)
    void init(int rx, int ry, int rz, T fill) { r[0] = rx; r[1] = ry; r[2] = rz; d.assign((size_t)rx * ry * rz, fill); }
    size_t size() const { return d.size(); }
    bool inb(int i, int j, int k) const { return i >= 0 && i < r[0] && j >= 0 && j < r[1] && k >= 0 && k < r[2]; }
    size_t lin(int i, int j, int k) const { return (size_t)i + (size_t)r[0] * ((size_t)j + (size_t)r[1] * (size_t)k); }
    T& at(int i, int j, int k) { return d[lin(i, j, k)]; }
    const T& at(int i, int j, int k) const { return d[lin(i, j, k)]; }
    T& at(const I3& c) { return d[lin(c[0], c[1], c[2])]; }
    const T& at(const I3& c) const { return d[lin(c[0], c[1], c[2])]; }
    // getFieldValue / operator(): border-aware read
    T get(int i, int j, int k) const {
        if (inb(i, j, k)) return d[lin(i, j, k)];
        if (!clampBorder) return border;
        i = std::min(std::max(i, 0), r[0] - 1); j = std::min(std::max(j, 0), r[1] - 1); k = std::min(std::max(k, 0), r[2] - 1);
        return d[lin(i, j, k)];
    }
    T get(const I3& c) const { return get(c[0], c[1], c[2]); }
};

// UT_VoxelArray iteration order (shim): 16^3 tiles, tile-linear x->y->z, x fastest inside.
template <typename F>
inline void forEachTileOrder(const int r[3], F&& f) {
    const int T = 16;
    for (int tk = 0; tk < r[2]; tk += T)
        for (int tj = 0; tj < r[1]; tj += T)
            for (int ti = 0; ti < r[0]; ti += T) {
                const int ke = std::min(tk + T, r[2]), je = std::min(tj + T, r[1]), ie = std::min(ti + T, r[0]);
                for (int k = tk; k < ke; ++k)
                    for (int j = tj; j < je; ++j)
                        for (int i = ti; i < ie; ++i) f(i, j, k);
            }
}

// own CSR honouring the Eigen semantics cited in SURVEY.md section 8c
struct Triplet { exint r, c; Real v; };
struct Csr {
    exint rows = 0, cols = 0;
    std::vector<exint> ptr;   // rows+1
    std::vector<int> idx;     // column indices (Eigen StorageIndex = int)
    std::vector<Real> val;
    exint nnz() const { return (exint)idx.size(); }
    void resize(exint r, exint c) { rows = r; cols = c; ptr.assign(r + 1, 0); idx.clear(); val.clear(); }
    // extern/eigen/Eigen/src/SparseCore/SparseMatrix.h:1024-1158: duplicates summed in
    // list order, explicit zeros kept, rows sorted by column.
    void setFromTriplets(const std::vector<Triplet>& t);
    Csr transpose() const;
    void mulVec(const Real* x, Real* y) const;           // y = M x
    void mulVecAdd(const Real* x, Real* y, Real s) const; // y += s M x
};
// extern/eigen/Eigen/src/SparseCore/ConservativeSparseSparseProduct.h:19-72
// (structural union, numerical zeros kept, result rows sorted)
Csr spgemm(const Csr& A, const Csr& B);
Csr scaleRows(const std::vector<Real>& diag, const Csr& B);  // diag * B
Csr addScaled(Real sa, const Csr& A, Real sb, const Csr& B); // sa*A + sb*B (union)

struct Params {
    int nx = 0, ny = 0, nz = 0;
    double dx = 1, dt = 1, density = 1;
    double tolerance = 1e-3;
    int maxIterations = 5000;
    int liquidLayers = 2, solidLayers = 2;
    int doReduced = 1, doTile = 1, tileSize = 16, tilePadding = 2;
    int doSolve = 1, keepNonConverged = 1;
    int threads = 0;   // OpenMP threads for SpMV / dots (0 = default)
};

struct Oracle {
    Params P;
    int nx, ny, nz;
    Real dx, invDx, dt, invDt;
    // inputs
    Field<float> surface, collision, viscosity;   // centre sampled
    Field<float> vel[3], colVel[3];               // face sampled
    // 14 weights, slot = Samp
    Field<float> liquidW[7], fluidW[7];
    // labels + indices, slot = Samp
    Field<exint> labels[7], activeIdx[7], reducedIdx[7];
    exint nCenter = 0, nFace[3] = {0, 0, 0}, nEdge[3] = {0, 0, 0};  // nEdge[edgeAxis]
    exint nActiveVs = 0, nReducedVs = 0, nPressures = 0, nStresses = 0, nTotalDOFs = 0, nSystemSize = 0;
    exint regionCount = 0;
    int fixLoops = 0; exint fixRemoved = 0;
    // region algebra
    std::vector<std::array<Real, 3>> com;
    std::vector<std::array<Real, RDOF>> bestFit;
    std::vector<std::array<Real, RDOF * RDOF>> Mr, Visc, Binv;
    // blocks
    Csr Mc, McInv, uInv, uMat, G, Dt, JG, JDt, MrMat, Bmat, BinvMat, A;
    std::vector<Real> activeRHS, reducedRHS, pressureRHS, stressRHS, oldActiveVs;
    std::vector<Real> b, solution, velSolution;
    std::vector<Real> guess;                    // guessVector (S_AS:413-419); zero unless constructGuessVectors ran
    // derived operator factors (Apply.h:24-68)
    Csr Gt, D, JGt, DJt, McInvG, McInvDt;
    std::vector<Real> mcInvDiag, uInvDiag;
    // results
    int solveIterations = -1; Real solveError = -1; int solverResult = INCOMPLETE; int usedBiCGStab = 0;
    double setupMs = 0, solveMs = 0;

    explicit Oracle(const Params& p);
    void resOf(int samp, int r[3]) const;
    void setInputs(const float* surf, const float* col, const float* visc, const float* const v[3], const float* const cv[3]);

    // --- the per-step driver sequence (exec/HDK_PolyStokes.C:344-584) ---
    void buildIntegrationWeightsAlt();          // S.cpp:238-326 (+ HDK computeSDFWeightsSampled, shim)
    void classifyCells();                       // S_Cls:56-128
    void constructReducedRegions();             // S_Cls:179-190
    void constructOnlyActiveRegions();          // S_Cls:192-199
    void classifyFaces();                       // S_Cls:201-207, 784-832
    void classifyEdges();                       // S_Cls:209-215, 1021-1067
    void constructCenterReducedIndices();       // S_Cls:217-239
    void constructFacesReducedIndices();        // S_Cls:241-247, 1473-1528
    void constructEdgesReducedIndices();        // S_Cls:249-255, 1534-1659
    void constructActiveIndices();              // S_Cls:257-284, 1738-1770
    void computeCenterOfMasses();               // S.cpp:328-372, 1274-1324
    void computeLeastSquaresFits();             // S.cpp:374-417, 1330-1399
    void computeReducedMassMatrices();          // S.cpp:419-441, 1405-1482
    void computeReducedViscosityMatricesInteriorOnly();  // S.cpp:468-490, 1484-1694
    void constructMatrixBlocks();               // S_CMB:9-868
    void assembleSystemPressureStressFactored();// S_AS:432-470 (+ S_AB:3-48,147-244,356-367)
    void assembleExplicitA();                   // S_AS:381-397
    void setupMatrixVectorProducts();           // Apply.h:24-68
    void applyMatrixVectorProducts(const Real* x, Real* y);  // Apply.h:102-179
    int solveSPDwithMatrixVectorPCG();          // S.cpp:734-812 (+ pcg.h:268-340, 134-200)
    void constructGuessVectors();               // S.cpp:512-531 (+ S_AS:413-419), the useWarmStart branch of PS.C:465-467
    int solveEigenCG();                         // S.cpp:814-862: Eigen::ConjugateGradient<Lower|Upper> + DiagonalPreconditioner on explicit A
    void buildValidFaces(float* const valid[3]);// S_Cls:4-54
    void recoverVelocityFromPressureStress();   // S.cpp:492-510
    void applySolutionToVelocity(float* velOut, const float* valid, int axis); // S.cpp:937-1028

    void runSetup();                            // PS.C:344-476
    // helpers
    void constructAirBoundaryLayer();           // S_Cls:291-508
    void constructSolidBoundaryLayer();         // S_Cls:510-703
    void constructTiles();                      // S_Cls:705-746
    void buildConnectedComponents();            // HDK SIM_VolumetricConnectedComponentBuilder (shim)
    void fixReducedRegionBoundaries();          // S_Cls:1073-1172
    void fixSmallReducedRegions();              // S_Cls:1174-1313, 1418-1467
    exint serialAssignFieldIndices(Field<exint>& idx, const Field<exint>& lab);  // S_Cls:1738-1770
    void overwriteIndices(Field<exint>& f, exint search, exint replace);          // S.cpp:1969-2004
    float getLocalViscosity(int samp, const I3& idx) const;   // S.cpp:1920-1924 + indexToPos (shim)
    void buildConversionCoefficients(const Real off[3], int axis, Real out[RDOF]) const;  // S.cpp:2107-2149
    exint faceVelocityDOF(exint i, int axis) const { return axis == 0 ? i : axis == 1 ? i + nFace[0] : i + nFace[0] + nFace[1]; }  // S.h:628-642
    exint centerStressDOF(exint i, int axis) const { return i + (exint)axis * nCenter; }   // S.h:586-606, 643-657
    exint edgeStressDOF(exint i, int edgeAxis) const {                                       // S.h:673-687: YZ, XZ, XY after 3 centre blocks
        exint o = 3 * nCenter;
        if (edgeAxis >= 1) o += nEdge[0];
        if (edgeAxis >= 2) o += nEdge[1];
        return i + o;
    }
    static bool isActive(exint l) { return l == ACTIVEFLUID || l == BOUNDARY; }     // S.h:708-710
    static bool isReduced(exint l) { return l == REDUCED || l == BOUNDARY; }        // S.h:711-713
    static bool isSolved(exint l) { return l == GENERICFLUID || l == ACTIVEFLUID || l == REDUCED || l == BOUNDARY; }  // S.h:717-722
};

// dense 26x26 helpers (Eigen semantics: S_AB:209 .inverse() = PartialPivLU; S.cpp:415 fullPivLu().solve)
void inversePartialPivLU(const Real* A, Real* Ainv, int n);
void solveFullPivLU(const Real* A, const Real* rhs, Real* x, int n, int* rankOut);

// extern/eigen/unsupported/Eigen/src/SparseExtra/MarketIO.h:311-372, 68-90
bool saveMarket(const Csr& m, const std::string& path);
bool saveMarketVector(const std::vector<Real>& v, const std::string& path);

}  // namespace orc
