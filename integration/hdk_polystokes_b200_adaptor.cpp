// hdk_polystokes_b200_adaptor.cpp -- see the header.  Include order in the plugin: the HDK headers and exec/HDK_PolyStokes.h first.
#include <cstring>
#include <vector>
#include "hdk_polystokes_b200_adaptor.h"

namespace {
// dense x-fastest copy of a SIM_RawField (face / edge fields are already sized +1 along their axes)
void toDense(const SIM_RawField& f, std::vector<float>& out) {
    int rx, ry, rz; f.getVoxelRes(rx, ry, rz);
    out.resize((size_t)rx * ry * rz);
    const UT_VoxelArrayF& a = *f.field();
    size_t q = 0;
    for (int z = 0; z < rz; ++z) for (int y = 0; y < ry; ++y) for (int x = 0; x < rx; ++x) out[q++] = a.getValue(x, y, z);
}
void fromDense(SIM_RawField& f, const std::vector<float>& in) {
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

    if (!state.handle || memcmp(&P, &state.params, sizeof P) != 0) {
        ps_destroy(state.handle); state.handle = nullptr; state.params = P;
        if (ps_create(&P, &state.handle) != PS_SUCCESS) { if (error) *error = ps_last_error(); return PS_FAILED; }
    }

    std::vector<float> surf, coll, visc, vel[3], cvel[3], valid[3];
    toDense(*surfaceField->getField(), surf); toDense(*collisionField->getField(), coll); toDense(*viscosityField->getField(), visc);
    for (int a = 0; a < 3; ++a) { toDense(*velocityField->getField(a), vel[a]); toDense(*collisionVelocityField->getField(a), cvel[a]); valid[a].resize(vel[a].size()); }

    ps_fields_in in;
    in.memory = PS_MEM_HOST; in.surface = surf.data(); in.collision = coll.data(); in.viscosity = visc.data();
    for (int a = 0; a < 3; ++a) { in.velocity[a] = vel[a].data(); in.collisionvel[a] = cvel[a].data(); }
    ps_fields_out out;
    out.memory = PS_MEM_HOST;
    for (int a = 0; a < 3; ++a) { out.velocity[a] = vel[a].data(); out.valid[a] = valid[a].data(); }        // velocity is overwritten in place on valid faces
    ps_stats stats;
    const int result = ps_step(state.handle, &in, &out, &stats);                          // == Solver::SolverResult (S.h:61-70)

    if (result == PS_UNSUPPORTED_SOLVER) { if (error) *error = "Unsupported Solver."; return result; }      // PS.C:530-534
    if (result == PS_FAILED || result == PS_INVALID) { if (error) *error = ps_last_error(); return result; }
    for (int a = 0; a < 3; ++a) fromDense(*validField->getField(a), valid[a]);            // PS.C:562
    if (result == PS_SUCCESS || node.getKeepNonConvergedResults()) {                      // PS.C:566-595
        for (int a = 0; a < 3; ++a) fromDense(*velocityField->getField(a), vel[a]);
        velocityField->pubHandleModification(); validField->pubHandleModification();
    } else if (error) *error = (result == PS_NOCONVERGE) ? "Solver did not converge, exiting..." : "Solver failed, exiting...";       // PS.C:597-604
    return result;
}

void polystokes_b200_release(polystokes_b200_node_state& state) { ps_destroy(state.handle); state.handle = nullptr; }
