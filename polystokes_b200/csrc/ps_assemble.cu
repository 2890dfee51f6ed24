// ps_assemble.cu -- ConstructMatrixBlocks on the GPU.
//
// The reference emits ~15 triplet lists and sorts them into CSR (exec/HDK_PolyStokesSolver_
// ConstructMatrixBlocks.cpp:9-292); every row has a fixed stencil, so here each thread writes its
// row straight into slot-major ELL -- no triplets, no sort.
//   K_ext  = [G D^T] on active face rows, followed by the same stencil on the *coupled reduced* face
//            rows (the J = C * K_red factorisation: JG/JD^T entries are `contribution * c_f(n)`,
//            S_CMB:443-456, 513-526, 601-614, i.e. a K-style row scaled by the basis row c_f)
//   K_ext^T as cell rows (<= 6 faces: pressure + the three centre stresses) and edge rows (<= 4 faces)
// Every value is +-(weight8 * weight8) / (64 dx): stored as the signed integer product (CompactOp, ps_solver.hpp) and
// rebuilt in the kernels as (double)code * (invDx / 64) -- bit-equal to the reference's products (S_CMB:362-380, 408-412).
#include "ps_solver.hpp"

namespace ps {

PS_D float local_viscosity_a(const Geom& g, const float* visc, int slot, const I3& idx) {
    int o[3] = {1, 1, 1};
    if (slot >= SL_FACE && slot < SL_EDGE) o[slot - SL_FACE] = 0;
    else if (slot >= SL_EDGE) { for (int a = 0; a < 3; ++a) if (a != slot - SL_EDGE) o[a] = 0; }
    const int b0 = o[0] ? idx.x : idx.x - 1, b1 = o[1] ? idx.y : idx.y - 1, b2 = o[2] ? idx.z : idx.z - 1;
    const float t0 = o[0] ? 0.f : 0.5f, t1 = o[1] ? 0.f : 0.5f, t2 = o[2] ? 0.f : 0.5f;
    float cz[2];
    for (int dz = 0; dz < 2; ++dz) {
        float cy[2];
        for (int dy = 0; dy < 2; ++dy) {
            const float a = float_at(g, visc, SL_CENTER, I3{b0, b1 + dy, b2 + dz});
            const float b = float_at(g, visc, SL_CENTER, I3{b0 + 1, b1 + dy, b2 + dz});
            cy[dy] = fadd_rn(a, fmul_rn(t0, fsub_rn(b, a)));
        }
        cz[dz] = fadd_rn(cy[0], fmul_rn(t1, fsub_rn(cy[1], cy[0])));
    }
    return fadd_rn(cz[0], fmul_rn(t2, fsub_rn(cz[1], cz[0])));
}

PS_D double clampd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }   // SYSclamp

// coefficient of one (face, cell|edge) pair: faceFluidW * liquidW * invDx (S_CMB:408-411, 480-483, 568-571)
PS_D double pair_coeff(int ffw8, int lw8, double invDx) { return mul_rn(mul_rn((double)ffw8 * 0.125, (double)lw8 * 0.125), invDx); }
PS_D uint64_t pack_code(uint64_t word, int k, int code) { return word | ((uint64_t)(uint8_t)(int8_t)code << (8 * k)); }

// K_ext rows: one thread per face of each axis.  Slots: 0,1 pressure of cell(-),cell(+); 2,3 centre stress;
// 4..7 the two edge axes (ascending) x (dir 0, dir 1).  Empty slots carry value 0 and repeat a valid column.
// Only the rows in `own` are written (all rows on one GPU / with replicated setup): the faces of the rank's window are visited, a
// face whose row belongs to another rank is skipped.
void k_assemble_K(cudaStream_t st, const Geom& g, const Fields& F, const Counts& C, CompactOp& Op, double* mcInv, double* mc, double* rhsU, double* oldVs, const RowSet& ownRows) {
    const RowSet own = ownRows;
    const int64_t nRows = C.nRowsExt, nAct = C.nActiveVs, nP = C.nPressures;
    uint64_t* kcode = Op.kcode.p; int32_t* kcol = Op.kcol.p; uint8_t* kmc = Op.kmc.p;
    {   // M_c^-1 by weight product (S_CMB:361-380): 1 / (rho * clamp(k / 64, MINWEIGHT^2, 1))
        double* lut = Op.mcInvLut.p; const double density = g.density;
        ps_for(st, 65, PS_LAMBDA(int64_t k) { const double volume = fmin(fmax((double)k * 0.125 * 0.125, 0.1 * 0.1), 1.0); lut[k] = 1. / mul_rn(volume, density); });
    }
    for (int axis = 0; axis < 3; ++axis) {
        const int32_t* krow = F.krow[axis];
        const uint8_t* ffw = F.fluW[SL_FACE + axis]; const uint8_t* flw = F.liqW[SL_FACE + axis];
        const uint8_t* clw = F.liqW[SL_CENTER]; const int32_t* CA = F.aidx[SL_CENTER];
        const int e1 = axis == 0 ? 1 : 0, e2 = axis == 2 ? 1 : 2;
        const int8_t* EL1 = F.label[SL_EDGE + e1]; const int8_t* EL2 = F.label[SL_EDGE + e2];
        const int32_t* EA1 = F.aidx[SL_EDGE + e1]; const int32_t* EA2 = F.aidx[SL_EDGE + e2];
        const uint8_t* ew1 = F.liqW[SL_EDGE + e1]; const uint8_t* ew2 = F.liqW[SL_EDGE + e2];
        const int64_t eOff1 = nP + C.stressOff[3 + e1], eOff2 = nP + C.stressOff[3 + e2];
        const float* vel = F.vel[axis];
        const double density = g.density;
        int64_t lo, hi;
        z_range(g, SL_FACE + axis, g.wzLo, g.wzHi, lo, hi);
        ps_for_range(st, lo, hi, PS_LAMBDA(int64_t q) {
            const int64_t row = krow[q];
            if (row < 0 || !own.has(row)) return;
            const I3 f = delin(g, SL_FACE + axis, q);
            const int fw = ffw[q];
            uint64_t word = 0; int32_t c[6] = {0, 0, 0, 0, 0, 0};
            for (int dir = 0; dir < 2; ++dir) {
                const I3 cell = dir ? f : shifted(f, axis, -1);
                if (!in_bounds(g, SL_CENTER, cell)) continue;
                const int64_t qc = lin(g, SL_CENTER, cell);
                const int ci = CA[qc];
                if (ci < 0) continue;
                const int prod = fw * (int)clw[qc];
                if (prod <= 0) continue;
                word = pack_code(word, dir, dir ? prod : -prod);          // G: gradientSign * coeff
                word = pack_code(word, 2 + dir, dir ? -prod : prod);      // D^T: -divergenceSign * coeff (column = nP + axis nC + ci)
                c[dir] = ci;
            }
            for (int dir = 0; dir < 2; ++dir) {
                const I3 ed1 = dir ? shifted(f, 3 - axis - e1, 1) : f;
                const int64_t q1 = lin(g, SL_EDGE + e1, ed1);
                if (is_active(EL1[q1])) {
                    const int prod = fw * (int)ew1[q1];
                    if (prod > 0) { word = pack_code(word, 4 + dir, dir ? -prod : prod); c[2 + dir] = (int32_t)(eOff1 + EA1[q1]); }
                }
                const I3 ed2 = dir ? shifted(f, 3 - axis - e2, 1) : f;
                const int64_t q2 = lin(g, SL_EDGE + e2, ed2);
                if (is_active(EL2[q2])) {
                    const int prod = fw * (int)ew2[q2];
                    if (prod > 0) { word = pack_code(word, 6 + dir, dir ? -prod : prod); c[4 + dir] = (int32_t)(eOff2 + EA2[q2]); }
                }
            }
            kcode[row] = word;
            kcol[row] = c[0] | (int32_t)((uint32_t)axis << 30);
            for (int s = 1; s < 6; ++s) kcol[(int64_t)s * nRows + row] = c[s];
            if (row < nAct) {
                // M_c, M_c^-1, rhs_u, old velocities (S_CMB:361-391); MINWEIGHT = 0.1 (S.h:226)
                const double MINWEIGHT = 0.1;
                double volume = (double)fw * 0.125 * ((double)flw[q] * 0.125);
                volume = clampd(volume, MINWEIGHT * MINWEIGHT, 1.0);
                const double m = mul_rn(volume, density);
                const double lv = (double)vel[q];
                mc[row] = m; mcInv[row] = 1. / m;
                kmc[row] = (uint8_t)(fw * (int)flw[q]);
                rhsU[row] = mul_rn(mul_rn(lv, volume), density);
                oldVs[row] = lv;
            }
        });
    }
}

// K_ext^T rows + the stress diagonal + the moving-solid right-hand sides, one thread per cell / edge.
// Rows of the rank's own cells / edges only (voxels of its slab; everything on one GPU / with replicated setup).
void k_assemble_Kt(cudaStream_t st, const Geom& g, const Fields& F, const Counts& C, CompactOp& Op, double* uInv, double* uDiag, double* rhsPT, bool ownOnly) {
    const int zA = ownOnly ? g.zLo : 0, zB = ownOnly ? g.zHi : g.nz;
    const int64_t nP = C.nPressures, nC = C.nCenter;
    const double invDx = g.invDx;
    const double MINWEIGHT = 0.1;
    {
        uint64_t* ccode = Op.ccode.p; int32_t* ccol = Op.ccol.p;
        const int32_t* CA = F.aidx[SL_CENTER];
        const uint8_t* clw = F.liqW[SL_CENTER]; const uint8_t* cfw = F.fluW[SL_CENTER];
        const int32_t* kr0 = F.krow[0]; const int32_t* kr1 = F.krow[1]; const int32_t* kr2 = F.krow[2];
        const uint8_t* fw0 = F.fluW[SL_FACE + 0]; const uint8_t* fw1 = F.fluW[SL_FACE + 1]; const uint8_t* fw2 = F.fluW[SL_FACE + 2];
        const float* cv0 = F.colvel[0]; const float* cv1 = F.colvel[1]; const float* cv2 = F.colvel[2];
        const float* visc = F.viscosity;
        const int64_t nAct = C.nActiveVs;
        int64_t lo, hi;
        z_range(g, SL_CENTER, zA, zB, lo, hi);
        ps_for_range(st, lo, hi, PS_LAMBDA(int64_t q) {
            const int ci = CA[q];
            if (ci < 0) return;
            const I3 c = delin(g, SL_CENTER, q);
            const int lw = clw[q], fwc = cfw[q];
            double rhsP = 0.;
            uint64_t word = 0; int32_t pcols[6];
            for (int axis = 0; axis < 3; ++axis) {
                const int32_t* kr = axis == 0 ? kr0 : axis == 1 ? kr1 : kr2;
                const uint8_t* fwA = axis == 0 ? fw0 : axis == 1 ? fw1 : fw2;
                const float* cvel = axis == 0 ? cv0 : axis == 1 ? cv1 : cv2;
                double rhsT = 0.;
                for (int side = 0; side < 2; ++side) {
                    const I3 f = side ? shifted(c, axis, 1) : c;
                    const int64_t qf = lin(g, SL_FACE + axis, f);
                    const int slot = axis * 2 + side;
                    pcols[slot] = 0;
                    const int64_t row = kr[qf];
                    if (row < 0) continue;
                    const int prod = (int)fwA[qf] * lw;
                    if (prod <= 0) continue;
                    const double coeff = pair_coeff(fwA[qf], lw, invDx);
                    // low face: this cell is its (+) cell (dir 1); high face: its (-) cell (dir 0)
                    const double sign = side ? -1. : 1.;               // gradientSign = divergenceSign
                    word = pack_code(word, slot, side ? -prod : prod);  // pressure row: sign * coeff; the stress row aa carries -sign * coeff
                    pcols[slot] = (int32_t)row;
                    if (row < nAct) {   // moving-solid terms exist for active faces only (S_CMB:417-441, 489-511)
                        const double svel = (double)cvel[qf];
                        const double sc = sign * coeff;
                        if (fwc < 8) { rhsP += -1. * sc * svel; rhsT += -1. * sc * svel; }
                        if (fwA[qf] < 8) { rhsP += sc * svel; rhsT += sc * svel; }
                    }
                }
                const int64_t trow = (int64_t)axis * nC + ci;
                rhsPT[nP + trow] = rhsT;
            }
            ccode[ci] = word;
            for (int s = 0; s < 6; ++s) ccol[(int64_t)s * nC + ci] = pcols[s];
            rhsPT[ci] = rhsP;
            // centre stress diagonal (S_CMB:772-819)
            const double volumeWeight = clampd((double)fwc * 0.125, MINWEIGHT, 1.0) * ((double)lw * 0.125);
            const double localViscosity = (double)visc[q];
            const double invLocalViscosity = clampd(1. / localViscosity, 0., 1.e10);
            const double ui = mul_rn(invLocalViscosity, clampd(volumeWeight, 1.e-2, 1.));
            const double ud = mul_rn(localViscosity, clampd(1. / volumeWeight, 0., 1.e2));
            for (int axis = 0; axis < 3; ++axis) { uInv[(int64_t)axis * nC + ci] = ui; uDiag[(int64_t)axis * nC + ci] = ud; }
        });
    }
    const int64_t nERows = C.nEdge[0] + C.nEdge[1] + C.nEdge[2];
    for (int e = 0; e < 3; ++e) {
        uint32_t* ecode = Op.ecode.p; int32_t* ecol = Op.ecol.p;
        const int8_t* EL = F.label[SL_EDGE + e]; const int32_t* EA = F.aidx[SL_EDGE + e];
        const uint8_t* elw = F.liqW[SL_EDGE + e]; const uint8_t* efw = F.fluW[SL_EDGE + e];
        const int fa0 = (e == 0) ? 1 : 0, fa1 = (e == 2) ? 1 : 2;
        const int32_t* krA = F.krow[fa0]; const int32_t* krB = F.krow[fa1];
        const uint8_t* fwA = F.fluW[SL_FACE + fa0]; const uint8_t* fwB = F.fluW[SL_FACE + fa1];
        const float* cvA = F.colvel[fa0]; const float* cvB = F.colvel[fa1];
        const float* visc = F.viscosity;
        const int64_t eRowOff = C.stressOff[3 + e] - 3 * nC;     // row inside the edge block
        const int64_t tOff = C.stressOff[3 + e];                 // row inside the stress block
        const int64_t nAct = C.nActiveVs;
        int64_t lo, hi;
        z_range(g, SL_EDGE + e, zA, zB, lo, hi);
        ps_for_range(st, lo, hi, PS_LAMBDA(int64_t q) {
            if (!is_active(EL[q])) return;
            const int ei = EA[q];
            const I3 ed = delin(g, SL_EDGE + e, q);
            const int lw = elw[q], fwe = efw[q];
            uint64_t word = 0; int32_t cidx[4] = {0, 0, 0, 0};
            double rhs = 0.;
            for (int k = 0; k < 4; ++k) {
                const int fa = k < 2 ? fa0 : fa1;
                const bool back = (k & 1) == 0;                      // order: back face, then face at the edge index
                const I3 f = back ? shifted(ed, 3 - fa - e, -1) : ed;
                if (!in_bounds(g, SL_FACE + fa, f)) continue;
                const int64_t qf = lin(g, SL_FACE + fa, f);
                const int64_t row = (k < 2 ? krA : krB)[qf];
                if (row < 0) continue;
                const int ffw8 = (k < 2 ? fwA : fwB)[qf];
                const int prod = ffw8 * lw;
                if (prod <= 0) continue;
                const double coeff = pair_coeff(ffw8, lw, invDx);
                const double divSign = back ? 1. : -1.;              // back face sees this edge at dir 1
                word = pack_code(word, k, back ? -prod : prod);      // -divSign * coeff
                cidx[k] = (int32_t)row;
                if (row < nAct) {   // S_CMB:581-599
                    const double svel = (double)(k < 2 ? cvA : cvB)[qf];
                    const double sc = divSign * coeff;
                    if (fwe < 8) rhs += -1. * sc * svel;
                    if (ffw8 < 8) rhs += sc * svel;
                }
            }
            const int64_t erow = eRowOff + ei;
            ecode[erow] = (uint32_t)word;
            for (int k = 0; k < 4; ++k) ecol[(int64_t)k * nERows + erow] = cidx[k];
            rhsPT[nP + tOff + ei] = rhs;
            // edge stress diagonal (S_CMB:685-711)
            const double volumeWeight = clampd((double)fwe * 0.125, MINWEIGHT, 1.0) * ((double)lw * 0.125);
            const double localViscosity = (double)local_viscosity_a(g, visc, SL_EDGE + e, ed);
            const double invLocalViscosity = clampd(1. / localViscosity, 0., 1e10);
            uInv[tOff + ei] = mul_rn(mul_rn(2., invLocalViscosity), volumeWeight);
            uDiag[tOff + ei] = mul_rn(mul_rn(0.5, localViscosity), clampd(1. / volumeWeight, 0., 1.e2));
        });
    }
}

}  // namespace ps
