// ref_adaptor.cpp -- TEST INFRASTRUCTURE.  Compiles the node-side binding integration/hdk_polystokes_b200_adaptor.cpp against the HDK
// stand-in (oracle/hdk_shim) and the reference's OWN node class declaration (exec/HDK_PolyStokes.h), and exposes one C call that builds
// the SIM fields of a scene, runs the adaptor and returns the fields it wrote back -- the drop-in boundary of INTEGRATION.md exercised
// with the reference's types.  In the plugin the adaptor links against libpolystokes_b200.so directly; here its ps_* calls are routed
// (by renaming them for this translation unit only -- the adaptor source is compiled as it stands) through pointers that refadp_bind()
// resolves with dlopen(RTLD_LOCAL) from the library the test names: the product, or its emulation twin in the CPU tests.  Nothing is loaded
// RTLD_GLOBAL, so the two libraries never see each other's symbols.
#include <cstring>
#include <dlfcn.h>
#define ps_create refadp_ps_create
#define ps_destroy refadp_ps_destroy
#define ps_step refadp_ps_step
#define ps_last_error refadp_ps_last_error
#define ps_create_multi refadp_ps_create_multi
#define ps_set_params refadp_ps_set_params
#define ps_alloc_pinned refadp_ps_alloc_pinned
#define ps_free_pinned refadp_ps_free_pinned
#include "polystokes_b200.h"
#include "hdk_shim.h"
#include <Eigen/Sparse>
#include <tbb/tbb.h>
#define private public
#define protected public
#include "HDK_PolyStokes.h"
#undef private
#undef protected
#include "../integration/hdk_polystokes_b200_adaptor.cpp"
#undef ps_create
#undef ps_destroy
#undef ps_step
#undef ps_last_error
#undef ps_create_multi
#undef ps_set_params
#undef ps_alloc_pinned
#undef ps_free_pinned

namespace {
struct Backend {
    void* lib = nullptr;
    int (*create)(const ps_params*, ps_handle*) = nullptr;
    void (*destroy)(ps_handle) = nullptr;
    int (*step)(ps_handle, const ps_fields_in*, ps_fields_out*, ps_stats*) = nullptr;
    const char* (*last_error)(void) = nullptr;
    int (*create_multi)(const ps_params*, int, const int*, ps_handle*) = nullptr;
    int (*set_params)(ps_handle, const ps_params*) = nullptr;
    void* (*alloc_pinned)(size_t) = nullptr;
    void (*free_pinned)(void*) = nullptr;
} g_backend;
}
extern "C" {
int refadp_ps_create(const ps_params* p, ps_handle* out) { return g_backend.create ? g_backend.create(p, out) : PS_FAILED; }
void refadp_ps_destroy(ps_handle h) { if (g_backend.destroy) g_backend.destroy(h); }
int refadp_ps_step(ps_handle h, const ps_fields_in* in, ps_fields_out* out, ps_stats* st) { return g_backend.step ? g_backend.step(h, in, out, st) : PS_FAILED; }
const char* refadp_ps_last_error(void) { return g_backend.last_error ? g_backend.last_error() : "refadp_bind was not called"; }
int refadp_ps_create_multi(const ps_params* p, int n, const int* devs, ps_handle* out) { return g_backend.create_multi ? g_backend.create_multi(p, n, devs, out) : PS_FAILED; }
int refadp_ps_set_params(ps_handle h, const ps_params* p) { return g_backend.set_params ? g_backend.set_params(h, p) : PS_FAILED; }
void* refadp_ps_alloc_pinned(size_t n) { return g_backend.alloc_pinned ? g_backend.alloc_pinned(n) : nullptr; }
void refadp_ps_free_pinned(void* p) { if (g_backend.free_pinned) g_backend.free_pinned(p); }
// which library stands behind the C ABI (0 on success)
int refadp_bind(const char* path) {
    void* lib = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!lib) return -1;
    g_backend.lib = lib;
    g_backend.create = (int (*)(const ps_params*, ps_handle*))dlsym(lib, "ps_create");
    g_backend.destroy = (void (*)(ps_handle))dlsym(lib, "ps_destroy");
    g_backend.step = (int (*)(ps_handle, const ps_fields_in*, ps_fields_out*, ps_stats*))dlsym(lib, "ps_step");
    g_backend.last_error = (const char* (*)(void))dlsym(lib, "ps_last_error");
    g_backend.create_multi = (int (*)(const ps_params*, int, const int*, ps_handle*))dlsym(lib, "ps_create_multi");
    g_backend.set_params = (int (*)(ps_handle, const ps_params*))dlsym(lib, "ps_set_params");
    g_backend.alloc_pinned = (void* (*)(size_t))dlsym(lib, "ps_alloc_pinned");
    g_backend.free_pinned = (void (*)(void*))dlsym(lib, "ps_free_pinned");
    return (g_backend.create && g_backend.destroy && g_backend.step && g_backend.last_error && g_backend.create_multi && g_backend.set_params && g_backend.alloc_pinned && g_backend.free_pinned) ? 0 : -2;
}
}

HDK_PolyStokes::HDK_PolyStokes(const SIM_DataFactory* factory) : GAS_SubSolver(factory) {}
HDK_PolyStokes::~HDK_PolyStokes() {}
bool HDK_PolyStokes::solveGasSubclass(SIM_Engine&, SIM_Object*, SIM_Time, SIM_Time) { return false; }
const SIM_DopDescription* HDK_PolyStokes::getDopDescription() { return nullptr; }
namespace { struct Node : HDK_PolyStokes { Node() : HDK_PolyStokes(nullptr) {} }; }

extern "C" {
struct refadp_params {
    int32_t nx, ny, nz;
    double dx, dt, density, tolerance;
    int32_t maxIterations, liquidLayers, solidLayers, doReducedRegions, doTile, tileSize, tilePadding, solverType, useWarmStart, keepNonConvergedResults;
};
static void fill(SIM_RawField* f, const float* src) { UT_VoxelArrayF& a = *f->fieldNC(); memcpy(a.d.data(), src, a.d.size() * sizeof(float)); a.expandAllTiles(); }

// velOut / validOut: 3 face-sampled float arrays each; steps: how many times the node is cooked on the same handle
// numDevices > 1: the node state asks for one ps_create_multi handle; doSolve: the node's "Do Solve" toggle
int refadp_run2(const refadp_params* P, const float* surface, const float* collision, const float* viscosity, const float* const* vel, const float* const* colvel,
                int steps, float* const* velOut, float* const* validOut, char* errorOut, int errorLen, int numDevices, int doSolve);
int refadp_run(const refadp_params* P, const float* surface, const float* collision, const float* viscosity, const float* const* vel, const float* const* colvel,
               int steps, float* const* velOut, float* const* validOut, char* errorOut, int errorLen) {
    return refadp_run2(P, surface, collision, viscosity, vel, colvel, steps, velOut, validOut, errorOut, errorLen, 1, 1);
}
int refadp_run2(const refadp_params* P, const float* surface, const float* collision, const float* viscosity, const float* const* vel, const float* const* colvel,
                int steps, float* const* velOut, float* const* validOut, char* errorOut, int errorLen, int numDevices, int doSolve) {
    std::map<std::string, double>& prm = hdk_shim::params();
    prm.clear();
    prm["matrixSetup"] = 0; prm["solverType"] = P->solverType; prm["useInputSurfaceWeights"] = 0; prm["useInputCollisionWeights"] = 0;
    prm["minDensity"] = 0; prm["maxDensity"] = 1e30; prm["activeLiquidBoundaryLayerSize"] = P->liquidLayers; prm["activeSolidBoundaryLayerSize"] = P->solidLayers;
    prm["doReducedRegions"] = P->doReducedRegions; prm["doTile"] = P->doTile; prm["tileSize"] = P->tileSize; prm["tilePadding"] = P->tilePadding;
    prm[SIM_NAME_TOLERANCE] = P->tolerance; prm["maxSolverIterations"] = P->maxIterations; prm["useWarmStart"] = P->useWarmStart;
    prm["exportMatrices"] = 0; prm["exportComponentMatrices"] = 0; prm["exportStats"] = 0; prm["doSolve"] = doSolve; prm["keepNonConvergedResults"] = P->keepNonConvergedResults;
    SIM_VectorField velocity, collisionVelocity, validFaces;
    SIM_ScalarField surf, coll, visc;
    const UT_Vector3 orig(0.f, 0.f, 0.f), size((float)(P->nx * P->dx), (float)(P->ny * P->dx), (float)(P->nz * P->dx));
    const SIM_FieldSample faceSample[3] = {SIM_SAMPLE_FACEX, SIM_SAMPLE_FACEY, SIM_SAMPLE_FACEZ};
    for (int a = 0; a < 3; ++a)
        for (SIM_VectorField* v : {&velocity, &collisionVelocity, &validFaces}) v->getField(a)->init(faceSample[a], orig, size, P->nx, P->ny, P->nz);
    for (SIM_ScalarField* s : {&surf, &coll, &visc}) s->getField()->init(SIM_SAMPLE_CENTER, orig, size, P->nx, P->ny, P->nz);
    fill(surf.getField(), surface); fill(coll.getField(), collision); fill(visc.getField(), viscosity);
    for (int a = 0; a < 3; ++a) fill(collisionVelocity.getField(a), colvel[a]);
    Node node;
    polystokes_b200_node_state state;
    state.numDevices = numDevices;
    std::string error;
    int result = PS_INCOMPLETE;
    for (int s = 0; s < steps; ++s) {
        for (int a = 0; a < 3; ++a) fill(velocity.getField(a), vel[a]);      // every cook starts from the same input velocity
        // substeps come with different dt (the handle must survive that: ps_set_params); the last cook uses the scene's dt
        const double dt = P->dt * (s + 1 < steps ? 0.5 + 0.25 * s : 1.0);
        const ps_handle before = state.handle;
        result = polystokes_b200_step(node, state, P->dx, dt, P->density, &velocity, &collisionVelocity, &surf, &coll, &visc, &validFaces, &error);
        if (s > 0 && state.handle != before) { error = "the adaptor rebuilt the handle although the grid did not change"; result = PS_FAILED; break; }
    }
    polystokes_b200_release(state);
    for (int a = 0; a < 3; ++a) {
        const UT_VoxelArrayF& v = *velocity.getField(a)->field(); memcpy(velOut[a], v.d.data(), v.d.size() * sizeof(float));
        const UT_VoxelArrayF& w = *validFaces.getField(a)->field(); memcpy(validOut[a], w.d.data(), w.d.size() * sizeof(float));
    }
    if (errorOut && errorLen > 0) { strncpy(errorOut, error.c_str(), (size_t)errorLen - 1); errorOut[errorLen - 1] = 0; }
    return result;
}
}
