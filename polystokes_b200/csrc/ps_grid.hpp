// ps_grid.hpp -- grid geometry, sample slots, labels and index maps shared by all kernels.
//
// Data layout in HBM: every grid field is a dense x-fastest array sized by its
// sample slot.  Labels are int8, DOF / region indices int32, the 14 volume-fraction weights are
// stored as uint8 "eighths" (the 2x2x2 sub-sampling of the reference can only yield k/8).
#pragma once
#include "ps_rt.hpp"

namespace ps {

// reference: exec/HDK_PolyStokesSolver.h:71-82 (MaterialLabels)
enum : int8_t {
    L_UNASSIGNED = -1, L_UNSOLVED = -2, L_GENERICFLUID = -3, L_ACTIVEFLUID = -4, L_SOLID = -5,
    L_REDUCED = -6, L_UNVISITED = -7, L_VISITED = -8, L_BOUNDARY = -9
};
// reference: exec/HDK_PolyStokesSolver.h:61-70 (SolverResult)
enum { R_UNSUPPORTED_SOLVER = -4, R_INCOMPLETE = -3, R_INVALID = -2, R_FAILED = -1, R_NOCONVERGE = 0, R_SUCCESS = 1, R_NOCHANGE = 2 };

// sample slots: 0 centre; 1..3 face x,y,z; 4..6 edge axis 0 (YZ), 1 (XZ), 2 (XY)
// (edge axis = direction the edge runs along; exec/HDK_PolyStokesSolver.h:397-411)
enum { SL_CENTER = 0, SL_FACE = 1, SL_EDGE = 4, N_SLOTS = 7 };
constexpr int RDOF = 26;   // lib/include/units.h:10-15

struct Geom {
    int nx, ny, nz;
    int r[N_SLOTS][3];
    int64_t n[N_SLOTS];
    double dx, invDx, dt, invDt, density;
    int zLo, zHi;          // owned z-slab [zLo, zHi) of this rank (whole grid on one GPU)
    // Slab-local setup (ps_part.hpp): the grid fields stay addressed by GLOBAL voxel index (so no stencil, position or tile
    // rule changes), but a rank only fills the cell layers [wzLo, wzHi) -- its slab plus a halo -- and the sweeps only visit
    // that window.  One GPU / replicated setup: the window is the whole grid.
    int wzLo, wzHi;
    int slabLocal;         // 1: slab-local setup (the fields exist on the window only, regions / numbering are per slab)
};

inline Geom make_geom(int nx, int ny, int nz, double dx, double dt, double density) {
    Geom g;
    g.nx = nx; g.ny = ny; g.nz = nz; g.dx = dx; g.invDx = 1. / dx; g.dt = dt; g.invDt = 1. / dt; g.density = density;
    for (int s = 0; s < N_SLOTS; ++s) {
        int r[3] = {nx, ny, nz};
        if (s >= SL_FACE && s < SL_EDGE) r[s - SL_FACE] += 1;
        else if (s >= SL_EDGE) { int e = s - SL_EDGE; for (int a = 0; a < 3; ++a) if (a != e) r[a] += 1; }
        for (int a = 0; a < 3; ++a) g.r[s][a] = r[a];
        g.n[s] = (int64_t)r[0] * r[1] * r[2];
    }
    g.zLo = 0; g.zHi = nz; g.wzLo = 0; g.wzHi = nz; g.slabLocal = 0;
    return g;
}

// voxel index range [lo, hi) of the layers z in [zA, zB) of a slot (clipped to the slot's extent; zB >= nz includes the slot's
// extra top layer); dense x-fastest layout: a z-range is one contiguous index range
PS_HD void z_range(const Geom& g, int slot, int zA, int zB, int64_t& lo, int64_t& hi) {
    const int rz = g.r[slot][2];
    const int a = zA < 0 ? 0 : (zA > rz ? rz : zA), b = zB >= g.nz ? rz : (zB < a ? a : zB);
    const int64_t plane = (int64_t)g.r[slot][0] * g.r[slot][1];
    lo = plane * a; hi = plane * b;
}
// does this rank own layer z of a slot?  (the extra top layer of a slot belongs to the last slab)
PS_HD bool owns_z(const Geom& g, int z) { return z >= g.zLo && (z < g.zHi || g.zHi >= g.nz); }

struct I3 { int x, y, z; };
PS_HD int comp(const I3& c, int a) { return a == 0 ? c.x : a == 1 ? c.y : c.z; }
PS_HD I3 shifted(I3 c, int a, int d) { if (a == 0) c.x += d; else if (a == 1) c.y += d; else c.z += d; return c; }

PS_HD bool in_bounds(const Geom& g, int slot, const I3& c) {
    return c.x >= 0 && c.x < g.r[slot][0] && c.y >= 0 && c.y < g.r[slot][1] && c.z >= 0 && c.z < g.r[slot][2];
}
PS_HD int64_t lin(const Geom& g, int slot, const I3& c) { return (int64_t)c.x + (int64_t)g.r[slot][0] * ((int64_t)c.y + (int64_t)g.r[slot][1] * (int64_t)c.z); }
PS_HD I3 delin(const Geom& g, int slot, int64_t q) {
    const int rx = g.r[slot][0], ry = g.r[slot][1];
    I3 c; c.x = (int)(q % rx); const int64_t t = q / rx; c.y = (int)(t % ry); c.z = (int)(t / ry); return c;
}
PS_HD I3 clamped(const Geom& g, int slot, I3 c) {
    c.x = c.x < 0 ? 0 : (c.x >= g.r[slot][0] ? g.r[slot][0] - 1 : c.x);
    c.y = c.y < 0 ? 0 : (c.y >= g.r[slot][1] ? g.r[slot][1] - 1 : c.y);
    c.z = c.z < 0 ? 0 : (c.z >= g.r[slot][2] ? g.r[slot][2] - 1 : c.z);
    return c;
}

// Position of voxel c in the UT_VoxelArray iteration order of the reference (16^3 tiles visited
// x->y->z, x fastest inside a tile; consumer: serialAssignFieldIndices,
// exec/HDK_PolyStokesSolver_Classifier.cpp:1738-1770).  Closed form, no table.
PS_HD int64_t tile_key(const Geom& g, int slot, const I3& c) {
    const int rx = g.r[slot][0], ry = g.r[slot][1], rz = g.r[slot][2];
    const int ti = c.x >> 4, tj = c.y >> 4, tk = c.z >> 4;
    const int tw = min(16, rx - (ti << 4)), th = min(16, ry - (tj << 4)), td = min(16, rz - (tk << 4));
    return (int64_t)(tk << 4) * rx * ry + (int64_t)(tj << 4) * rx * td + (int64_t)(ti << 4) * th * td
         + (c.x & 15) + tw * ((c.y & 15) + th * (c.z & 15));
}

// border-aware reads.  Index/label fields: constant UNASSIGNED outside (reference:
// exec/HDK_PolyStokesSolver.cpp:101-152); float-like fields: clamp to edge (HDK shim)
PS_HD int label_at(const Geom& g, const int8_t* f, int slot, const I3& c) { return in_bounds(g, slot, c) ? (int)f[lin(g, slot, c)] : (int)L_UNASSIGNED; }
PS_HD int index_at(const Geom& g, const int32_t* f, int slot, const I3& c) { return in_bounds(g, slot, c) ? f[lin(g, slot, c)] : -1; }
PS_HD int weight8_at(const Geom& g, const uint8_t* f, int slot, const I3& c) { return (int)f[lin(g, slot, clamped(g, slot, c))]; }
PS_HD float float_at(const Geom& g, const float* f, int slot, const I3& c) { return f[lin(g, slot, clamped(g, slot, c))]; }

// reference predicates: exec/HDK_PolyStokesSolver.h:708-722
PS_HD bool is_active(int l) { return l == L_ACTIVEFLUID || l == L_BOUNDARY; }
PS_HD bool is_reduced(int l) { return l == L_REDUCED || l == L_BOUNDARY; }

// all device-resident grid fields of one solver instance
struct Fields {
    const float* surface; const float* collision; const float* viscosity;
    const float* vel[3]; const float* colvel[3];
    uint8_t* liqW[N_SLOTS]; uint8_t* fluW[N_SLOTS];
    int8_t* label[N_SLOTS];
    int32_t* aidx[N_SLOTS];
    int32_t* ridx[N_SLOTS];
    int32_t* krow[3];        // per face: row of K_ext (active DOF | nActiveVs + coupled reduced row | -1)
};

// the 26-term divergence-free quadratic basis row of a face sample (reference:
// exec/HDK_PolyStokesSolver.cpp:2107-2149, QUADRATIC_REGIONS)
PS_HD void conversion_coefficients(double ox, double oy, double oz, int axis, double* v) {
#pragma unroll
    for (int n = 0; n < RDOF; ++n) v[n] = 0.;
    if (axis == 0) {
        v[0] = 1.; v[3] = ox; v[4] = oy; v[5] = oz;
        v[6] = ox * ox; v[7] = ox * oy; v[8] = ox * oz; v[9] = oy * oy; v[10] = oy * oz; v[11] = oz * oz;
    } else if (axis == 1) {
        v[1] = 1.; v[12] = ox; v[13] = oy; v[14] = oz;
        v[15] = ox * ox; v[16] = ox * oy; v[17] = ox * oz; v[18] = oy * oy; v[19] = oy * oz; v[20] = oz * oz;
    } else {
        v[2] = 1.; v[3] = -oz;
        v[6] = -2. * ox * oz; v[7] = -1. * oy * oz; v[8] = -0.5 * oz * oz;
        v[13] = -oz; v[16] = -1. * ox * oz; v[18] = -2. * oy * oz; v[19] = -0.5 * oz * oz;
        v[21] = ox; v[22] = oy; v[23] = ox * ox; v[24] = ox * oy; v[25] = oy * oy;
    }
}
// offset of a face sample from a region's centre of mass: (index - 1/2 on the face axis) * dx - COM
// (reference: exec/HDK_PolyStokesSolver_ConstructMatrixBlocks.cpp:446-450)
PS_D void face_offset(const Geom& g, const I3& f, int axis, const double* com, double& ox, double& oy, double& oz) {
    double px = (double)f.x, py = (double)f.y, pz = (double)f.z;
    if (axis == 0) px -= 0.5; else if (axis == 1) py -= 0.5; else pz -= 0.5;
    ox = sub_rn(mul_rn(px, g.dx), com[0]); oy = sub_rn(mul_rn(py, g.dx), com[1]); oz = sub_rn(mul_rn(pz, g.dx), com[2]);
}
// the 10 monomials {1,x,y,z,xx,xy,xz,yy,yz,zz} of a coupled reduced face row's offset from its region's centre of mass;
// `packed` = x | y<<10 | z<<20 | axis<<30 (RegionData::rowXYZ)
PS_D void row_monomials(double dx, uint32_t packed, const double* com, double* m) {
    const int axis = (int)(packed >> 30);
    double px = (double)(packed & 1023u), py = (double)((packed >> 10) & 1023u), pz = (double)((packed >> 20) & 1023u);
    if (axis == 0) px -= 0.5; else if (axis == 1) py -= 0.5; else pz -= 0.5;
    const double ox = sub_rn(mul_rn(px, dx), com[0]), oy = sub_rn(mul_rn(py, dx), com[1]), oz = sub_rn(mul_rn(pz, dx), com[2]);
    m[0] = 1.; m[1] = ox; m[2] = oy; m[3] = oz; m[4] = ox * ox; m[5] = ox * oy; m[6] = ox * oz; m[7] = oy * oy; m[8] = oy * oz; m[9] = oz * oz;
}

}  // namespace ps
