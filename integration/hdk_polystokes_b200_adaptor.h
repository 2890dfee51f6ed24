// hdk_polystokes_b200_adaptor.h -- the node-side binding of INTEGRATION.md section 2 as a compilable unit.
//
// What a maintainer of panuelosj/polystokes adds to exec/: HDK_PolyStokes::solveGasSubclass (exec/HDK_PolyStokes.C:222-609) keeps its
// field fetching and validation (PS.C:235-328) and then, instead of constructing HDK_PolyStokes::Solver and running its 24 stages
// (PS.C:329-584), calls polystokes_b200_step() with the same field objects.  The function marshals the SIM fields into dense x-fastest
// arrays, makes ONE C call (ps_step, include/polystokes_b200.h) and writes the velocity and `valid` fields back (PS.C:562-584).
// It uses HDK types only through the members the reference itself uses (SIM_VectorField::getField, SIM_RawField::field / fieldNC /
// getVoxelRes, UT_VoxelArray::getValue / setValue), so the same source compiles against the real HDK in the plugin and against
// oracle/hdk_shim here, where tests/test_zz_adaptor.py runs it next to the compiled reference solver.
#pragma once
#include <string>
#include "polystokes_b200.h"

struct polystokes_b200_node_state {      // lives in the node instance (one handle per node; the grid rarely changes between substeps)
    ps_handle handle = nullptr;
    ps_params params = {};
    int numDevices = 1;                  // > 1: one ps_create_multi handle over devices 0 .. numDevices-1 (the cook thread stays single)
    // page-locked staging of the 9 input and 6 output fields (ps_alloc_pinned), reused while the grid keeps its size
    float* stage[15] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t stageCount[15] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
};

// Returns the SolverResult of the step (S.h:61-70 == PS_* of polystokes_b200.h); on PS_FAILED / PS_INVALID *error holds ps_last_error().
// `node` supplies the parameters through the reference's own accessors (exec/HDK_PolyStokes.h:19-40).
int polystokes_b200_step(HDK_PolyStokes& node, polystokes_b200_node_state& state, fpreal dx, fpreal dt, fpreal constantDensity,
                         SIM_VectorField* velocityField, const SIM_VectorField* collisionVelocityField, const SIM_ScalarField* surfaceField,
                         const SIM_ScalarField* collisionField, const SIM_ScalarField* viscosityField, SIM_VectorField* validField, std::string* error);
void polystokes_b200_release(polystokes_b200_node_state& state);
