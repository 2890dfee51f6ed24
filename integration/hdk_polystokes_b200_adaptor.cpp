// hdk_polystokes_b200_adaptor.cpp -- see the header.  Include order in the plugin: the HDK headers and exec/HDK_PolyStokes.h first.
#include <cstring>
#include <vector>
#include "hdk_polystokes_b200_adaptor.h"

namespace {
size_t voxels(const SIM_RawField& f) { int rx, ry, rz; f.getVoxelRes(rx, ry, rz); return (size_t)rx * ry * rz; }
// page-locked staging array `slot` of the node state, grown on demand
float* stage(polystokes_b200_node_state& st, int slot, size_t count) {
    if (st.stageCount[slot] < count) {
        ps_free_pinned(st.stage[slot]);
        st.stage[slot] = (float*)ps_alloc_pinned(count * sizeof(float)); st.stageCount[slot] = st.stage[slot] ? count : 0;
    }
    return st.stage[slot];
}
// dense x-fastest copy of a SIM_RawField (face / edge fields are already sized +1 along their axes)
void toDense(const SIM_RawField& f, float* out) {
    int rx, ry, rz; f.getVoxelRes(rx, ry, rz);
    const UT_VoxelArrayF& a = *f.field();
    size_t q = 0;
    for (int z = 0; z < rz; ++z) for (int y = 0; y < ry; ++y) for (int x = 0; x < rx; ++x) out[q++] = a.getValue(x, y, z);
}
void fromDense(SIM_RawField& f, const float* in) {
    int rx, ry, rz; f.getVoxelRes(rx, ry, rz);
    UT_VoxelArrayF& a = *f.fieldNC();
    size_t q = 0;
    for (int z = 0; z < rz; ++z) for (int y = 0; y < ry; ++y) for (int x = 0; x < rx; ++x) a.setValue(x, y, z, in[q++]);
}
int cancelRequested(void*) { return UTgetInterrupt()->opInterrupt() ? 1 : 0; }       // same polling as the reference's loops (S_Cls:73)
}  // namespace

int polystokes_b200_step(HDK_PolyStokes& node, polystokes_b200_node_state& state, fpreal dx, fpreal dt, fpreal constantDensity,
                         SIM_VectorField* velocityField, const SIM_VectorField* collisionVelocityField, const SIM_ScalarField* surfaceField,
                         const SIM_ScalarField* collisionField, const SIM_ScalarField* viscosityField, SIM_VectorField* validField, std::string* error)
{
    ps_params P;
    memset(&P, 0, sizeof P);
    const UT_Vector3 res = velocityField->getTotalVoxelRes();                            // S.cpp:41-44
    P.nx = (int32_t)res.x(); P.ny = (int32_t)res.y(); P.nz = (int32_t)res.z();
    P.dx = dx; P.dt = dt;                                                                 // PS.C:319-320
    P.constantDensity = constantDensity;                                                  // PS.C:298-304
    P.tolerance = node.getSolverTolerance();                                              // PS.h:19-40
    P.maxSolverIterations = node.getSolverMaxIterations();
    P.activeLiquidBoundaryLayerSize = node.getActiveLiquidBoundaryLayerSize();
    P.activeSolidBoundaryLayerSize = node.getActiveSolidBoundaryLayerSize();
    P.doReducedRegions = node.getDoReducedRegions(); P.doTile = node.getDoTile();
    P.tileSize = node.getTileSize();                 P.tilePadding = node.getTilePadding();
    P.exportMatrices = node.getExportMatrices();     P.exportComponentMatrices = node.getExportComponentMatrices();
    P.exportStats = node.getExportStats();
    UT_String prefix; node.getExportDataPrefix(prefix);
    strncpy(P.exportDataPrefix, prefix.c_str(), sizeof P.exportDataPrefix - 1);
    P.doSolve = node.getDoSolve();                   P.keepNonConvergedResults = node.getKeepNonConvergedResults();
    P.useWarmStart = node.getUseWarmStart();
    P.matrixSetup = (int32_t)node.getMatrixScheme(); P.solverType = (int32_t)node.getSolverType();
    P.useInputSurfaceWeights = node.getUseInputSurfaceWeights();
    P.useInputCollisionWeights = node.getUseInputCollisionWeights();
    P.minDensity = node.getMinDensity();             P.maxDensity = node.getMaxDensity();
    P.cancel_cb = cancelRequested;

    // a handle is rebuilt only when the grid changes; dt, density and every node parameter are per-step values (ps_set_params)
    const ps_params& Q = state.params;
    const bool sameGrid = state.handle && Q.nx == P.nx && Q.ny == P.ny && Q.nz == P.nz && Q.dx == P.dx;
    if (!sameGrid) {
        ps_destroy(state.handle); state.handle = nullptr;
        const int rc = state.numDevices > 1 ? ps_create_multi(&P, state.numDevices, nullptr, &state.handle) : ps_create(&P, &state.handle);
        if (rc != PS_SUCCESS) { if (error) *error = ps_last_error(); return PS_FAILED; }
    } else if (ps_set_params(state.handle, &P) != PS_SUCCESS) { if (error) *error = ps_last_error(); return PS_FAILED; }
    state.params = P;

    const SIM_RawField* inF[9] = {surfaceField->getField(), collisionField->getField(), viscosityField->getField(),
                                  velocityField->getField(0), velocityField->getField(1), velocityField->getField(2),
                                  collisionVelocityField->getField(0), collisionVelocityField->getField(1), collisionVelocityField->getField(2)};
    float* buf[15];
    for (int i = 0; i < 9; ++i) { buf[i] = stage(state, i, voxels(*inF[i])); if (!buf[i]) { if (error) *error = ps_last_error(); return PS_FAILED; } toDense(*inF[i], buf[i]); }
    for (int a = 0; a < 3; ++a) {        // outputs: velocity (invalid faces keep the input value: the library starts from the input field) and valid
        buf[9 + a] = stage(state, 9 + a, voxels(*inF[3 + a])); buf[12 + a] = stage(state, 12 + a, voxels(*inF[3 + a]));
        if (!buf[9 + a] || !buf[12 + a]) { if (error) *error = ps_last_error(); return PS_FAILED; }
    }

    ps_fields_in in;
    in.memory = PS_MEM_HOST; in.surface = buf[0]; in.collision = buf[1]; in.viscosity = buf[2];
    for (int a = 0; a < 3; ++a) { in.velocity[a] = buf[3 + a]; in.collisionvel[a] = buf[6 + a]; }
    ps_fields_out out;
    out.memory = PS_MEM_HOST;
    for (int a = 0; a < 3; ++a) { out.velocity[a] = buf[9 + a]; out.valid[a] = buf[12 + a]; }
    ps_stats stats;
    const int result = ps_step(state.handle, &in, &out, &stats);                          // == Solver::SolverResult (S.h:61-70)

    if (result == PS_UNSUPPORTED_SOLVER) { if (error) *error = "Unsupported Solver."; return result; }      // PS.C:530-534
    if (result == PS_FAILED || result == PS_INVALID) { if (error) *error = ps_last_error(); return result; }
    for (int a = 0; a < 3; ++a) fromDense(*validField->getField(a), buf[12 + a]);         // PS.C:562
    if (result == PS_SUCCESS || node.getKeepNonConvergedResults())                        // PS.C:565-583 (also without doSolve: the zero solution is written back)
        for (int a = 0; a < 3; ++a) fromDense(*velocityField->getField(a), buf[9 + a]);
    if (node.getDoSolve()) {                                                              // PS.C:588-605: reporting only when a solve was asked for
        if (result == PS_SUCCESS || node.getKeepNonConvergedResults()) { velocityField->pubHandleModification(); validField->pubHandleModification(); }
        else if (error) *error = (result == PS_NOCONVERGE) ? "Solver did not converge, exiting..." : "Solver failed, exiting...";
    }
    return result;
}

void polystokes_b200_release(polystokes_b200_node_state& state) {
    ps_destroy(state.handle); state.handle = nullptr;
    for (int i = 0; i < 15; ++i) { ps_free_pinned(state.stage[i]); state.stage[i] = nullptr; state.stageCount[i] = 0; }
}
