// ps_explicit.cu -- the explicit system matrix of assembleSystemPressureStress
// (exec/HDK_PolyStokesSolver_AssembleSystem.cpp:351-430), built on the GPU from the factors.
//
// Reference: four sparse triple products
//     A = -dt [G D^T]^T Mc^-1 [G D^T]  -  [JG JD^T]^T B^-1 [JG JD^T]  -  1/2 [0 0; 0 mu^-1]
// evaluated by Eigen (structural union of the products, no pruning of numerical zeros; ConservativeSparseSparseProduct.h:19-72)
// and merged by setFromTriplets (S_AS:381-397).  It is what solverType EIGEN solves with and what exportMatrices writes.
// On this path the matrix is needed for export / parity only (the solvers keep A factored), so the build favours a simple,
// exactly reproducible structure over speed -- still all of it runs on the device:
//   row i of A  =  U_i  u  R_i  (u {i} on stress rows)
//   U_i : columns of the ACTIVE face rows of K_ext adjacent to DOF i          (<= 6 faces x 8 slots, merged in registers)
//   R_i : every column touched by the region r that touches i (JG / JD^T store all 26 rows of a touched column, explicit
//         zeros included: S_CMB:443-456, 513-526, 601-614, so the block is structurally dense); value v_i^T B_r^-1 v_j with
//         v_j = column j of J_r = sum_f K_fj c_f (J = C K_red is never stored, ps_solver.hpp)
// Steps: (1) region of every DOF column (atomic max / min over the coupled reduced rows; C6 guarantees one region per column,
// checked), (2) per-region ascending column lists (select + stable radix sort by region), (3) V = J columns and W = B^-1 V
// per list entry, (4) one thread per row counts its merged length, exclusive scan, (5) the same thread writes its row.
#include "ps_solver.hpp"

namespace ps {

struct ExplCtx {
    OpArgs A;
    const int32_t* rowRegion; const uint32_t* rowXYZ; const double* com; const double* Binv;
    double dx, dt;
    const int32_t* regionOfCol;   // [n] region touching the DOF column, -1 if none
    const int32_t* regStart;      // [R+1] range of the region's columns inside listCol
    const int32_t* listCol;       // columns touched by a region, sorted by (region, column)
    const int32_t* posOfCol;      // [n] index of the column inside listCol
    const double* VW;             // [len(listCol)][52]: v_j (26) then B^-1 v_j (26)
};

// face rows of K_ext adjacent to system row i with their K_fi (the transposed blocks, CompactOp): returns the slot count
PS_D int row_faces(const OpArgs& A, int64_t i, int64_t* f, double* val) {
    if (i < A.nP) {
        const uint64_t word = A.ccode[i];
        for (int k = 0; k < 6; ++k) { const int code = op_code(word, k); val[k] = (double)code * A.valScale; f[k] = code ? A.ccol[(int64_t)k * A.nC + i] : -1; }
        return 6;
    }
    const int64_t j = i - A.nP;
    if (j < 3 * A.nC) {
        const int a = (int)(j / A.nC); const int64_t ci = j - (int64_t)a * A.nC;
        const uint64_t word = A.ccode[ci];
        for (int k = 0; k < 2; ++k) { const int code = op_code(word, 2 * a + k); val[k] = -((double)code * A.valScale); f[k] = code ? A.ccol[(int64_t)(2 * a + k) * A.nC + ci] : -1; }
        return 2;
    }
    const int64_t e = j - 3 * A.nC;
    const uint32_t word = A.ecode[e];
    for (int k = 0; k < 4; ++k) { const int code = op_code(word, k); val[k] = (double)code * A.valScale; f[k] = code ? A.ecol[(int64_t)k * A.nE + e] : -1; }
    return 4;
}
// the 8 (column, value) slots of face row r of K_ext (same decode as k_row, ps_pcg.cu); empty slots have value 0
PS_D void face_row_entries(const OpArgs& A, int64_t r, int64_t* col, double* val) {
    const uint64_t word = A.kcode[r];
    const int32_t c0w = A.kcol[r];
    const int64_t cOff = A.nP + (int64_t)((uint32_t)c0w >> 30) * A.nC;
    col[0] = c0w & OP_COL_MASK; col[1] = A.kcol[A.nRowsExt + r]; col[2] = cOff + col[0]; col[3] = cOff + col[1];
    for (int k = 0; k < 4; ++k) col[4 + k] = A.kcol[(int64_t)(2 + k) * A.nRowsExt + r];
    for (int k = 0; k < 8; ++k) val[k] = (double)op_code(word, k) * A.valScale;
}

constexpr int U_MAX = 49;   // 6 faces x 8 slots + the diagonal
// U_i: sorted, duplicate-free (column, sum_f K_fi (Mc^-1_f K_fj)) list of row i; stress rows always carry their diagonal
PS_D int build_u_list(const OpArgs& A, int64_t i, const int64_t* f, const double* kv, int ns, int32_t* ucol, double* uval) {
    int n = 0;
    auto insert = [&](int32_t c, double v) {
        int lo = 0;
        while (lo < n && ucol[lo] < c) ++lo;
        if (lo < n && ucol[lo] == c) { uval[lo] += v; return; }
        for (int q = n; q > lo; --q) { ucol[q] = ucol[q - 1]; uval[q] = uval[q - 1]; }
        ucol[lo] = c; uval[lo] = v; ++n;
    };
    for (int k = 0; k < ns; ++k) {
        if (kv[k] == 0. || f[k] >= A.nActiveVs) continue;
        int64_t col[8]; double val[8];
        face_row_entries(A, f[k], col, val);
        const double mcInv = A.mcInvLut[A.kmc[f[k]]];
        for (int s = 0; s < 8; ++s) if (val[s] != 0.) insert((int32_t)col[s], kv[k] * (mcInv * val[s]));
    }
    if (i >= A.nP) insert((int32_t)i, 0.);
    return n;
}

// (1) region of every column touched by a coupled reduced row
static void k_mark_region_columns(cudaStream_t st, const OpArgs& A, const int32_t* rowRegion, int32_t* regMax, int32_t* regMin) {
    ps_for(st, A.nRowsExt - A.nActiveVs, PS_LAMBDA(int64_t row) {
        int64_t col[8]; double val[8];
        face_row_entries(A, A.nActiveVs + row, col, val);
        const int r = rowRegion[row];
        for (int s = 0; s < 8; ++s) if (val[s] != 0.) { atomic_max(&regMax[col[s]], r); atomic_min(&regMin[col[s]], r); }
    });
}
static void k_region_column_flags(cudaStream_t st, int64_t n, const int32_t* regMax, int32_t* regMin, uint8_t* flag, int* conflict) {
    ps_for(st, n, PS_LAMBDA(int64_t c) {
        const int r = regMax[c];
        flag[c] = r >= 0 ? 1 : 0;
        if (r >= 0 && regMin[c] != r) atomic_or(conflict, 1);
    });
}
static void k_gather_regions(cudaStream_t st, int64_t m, const int32_t* listCol, const int32_t* regionOfCol, int32_t* keys, int* counts) {
    ps_for(st, m, PS_LAMBDA(int64_t l) { const int r = regionOfCol[listCol[l]]; keys[l] = r; atomic_add(&counts[r], 1); });
}
static void k_positions(cudaStream_t st, int64_t m, const int32_t* listCol, int32_t* posOfCol) {
    ps_for(st, m, PS_LAMBDA(int64_t l) { posOfCol[listCol[l]] = (int32_t)l; });
}
// (3) v_j = column j of J_r and w_j = B_r^-1 v_j
static void k_region_columns_vw(cudaStream_t st, int64_t m, const ExplCtx X, double* VW) {
    ps_for(st, m, PS_LAMBDA(int64_t l) {
        const OpArgs& A = X.A;
        const int64_t j = X.listCol[l];
        const int region = X.regionOfCol[j];
        int64_t f[6]; double kv[6];
        const int ns = row_faces(A, j, f, kv);
        double V[RDOF], c[RDOF], mono[10];
        for (int n = 0; n < RDOF; ++n) V[n] = 0.;
        for (int k = 0; k < ns; ++k) {
            if (kv[k] == 0. || f[k] < A.nActiveVs) continue;
            const uint32_t packed = X.rowXYZ[f[k] - A.nActiveVs];
            row_monomials(X.dx, packed, X.com + 3 * region, mono);
            conversion_coefficients(mono[1], mono[2], mono[3], (int)(packed >> 30), c);
            for (int n = 0; n < RDOF; ++n) V[n] += kv[k] * c[n];
        }
        const double* B = X.Binv + (size_t)region * RDOF * RDOF;
        double* out = VW + (size_t)l * 2 * RDOF;
        for (int n = 0; n < RDOF; ++n) out[n] = V[n];
        for (int n = 0; n < RDOF; ++n) { double t = 0.; for (int q = 0; q < RDOF; ++q) t += B[n * RDOF + q] * V[q]; out[RDOF + n] = t; }
    });
}
// (4) / (5): one thread per row merges U_i with the column list of its region
template <bool FILL>
static void k_rows(cudaStream_t st, const ExplCtx X, int64_t* len, const int64_t* ptr, int32_t* idx, double* val) {
    ps_for(st, X.A.nP + X.A.nT, PS_LAMBDA(int64_t i) {
        const OpArgs& A = X.A;
        int64_t f[6]; double kv[6];
        const int ns = row_faces(A, i, f, kv);
        int32_t ucol[U_MAX]; double uval[U_MAX];
        const int nU = build_u_list(A, i, f, kv, ns, ucol, uval);
        const int region = X.regionOfCol[i];
        const int32_t rLo = region >= 0 ? X.regStart[region] : 0, rHi = region >= 0 ? X.regStart[region + 1] : 0;
        if (!FILL) {
            int64_t n = rHi - rLo;
            for (int u = 0; u < nU; ++u) if (region < 0 || X.regionOfCol[ucol[u]] != region) ++n;
            len[i] = n;
            return;
        }
        const double diagTerm = i >= A.nP ? -0.5 * A.uInv[i - A.nP] : 0.;
        const double* Vi = region >= 0 ? X.VW + (size_t)X.posOfCol[i] * 2 * RDOF : nullptr;
        int64_t out = ptr[i];
        int u = 0; int32_t t = rLo;
        while (u < nU || t < rHi) {
            const int32_t cu = u < nU ? ucol[u] : INT32_MAX, ct = t < rHi ? X.listCol[t] : INT32_MAX;
            const int32_t c = cu < ct ? cu : ct;
            double v = 0.;
            if (cu == c) { v = -X.dt * uval[u]; ++u; }
            if (ct == c) {
                const double* Wj = X.VW + (size_t)t * 2 * RDOF + RDOF;
                double q = 0.;
                for (int n = 0; n < RDOF; ++n) q += Vi[n] * Wj[n];
                v += -1. * q; ++t;
            }
            if (c == (int32_t)i) v += diagTerm;
            idx[out] = c; val[out] = v; ++out;
        }
    });
}

void Solver::buildExplicitA() {
    if (haveA) return;
    if (part.multi()) throw Error("explicit A: single GPU only (the matrix couples regions of neighbouring slabs)");
    const OpArgs A = make_op_args();
    const int64_t n = C.nSystemSize;
    const int R = RG.count;
    Aexp.n = n; Aexp.nnz = 0;
    Aexp.ptr.alloc((size_t)n + 2);
    Aexp.ptr.zero(st, (size_t)n + 2);
    if (n == 0) { haveA = true; return; }
    DBuf<int32_t> regMax, regMin, posOfCol, listCol, keys, keysTmp, valsTmp, regStart;
    DBuf<uint8_t> flag; DBuf<int> counts; DBuf<double> VW; DBuf<int64_t> len;
    regMax.alloc((size_t)n); regMin.alloc((size_t)n); posOfCol.alloc((size_t)n); flag.alloc((size_t)n);
    regMax.fill_byte(st, 0xFF, (size_t)n);                                   // -1
    dev_memset(regMin.p, 0x7F, (size_t)n * sizeof(int32_t), st);             // 0x7f7f7f7f > any region id
    posOfCol.fill_byte(st, 0xFF, (size_t)n);
    int64_t m = 0;
    std::vector<int32_t> hStart((size_t)R + 1, 0);
    if (R > 0 && A.nRowsExt > A.nActiveVs) {
        counts.alloc((size_t)R + 1); counts.zero(st, (size_t)R + 1);
        k_mark_region_columns(st, A, RG.rowRegion.p, regMax.p, regMin.p);
        k_region_column_flags(st, n, regMax.p, regMin.p, flag.p, counts.p + R);
        m = select_flagged(st, n, flag.p, listCol, 0);                       // ascending columns
        if (m > 0) {
            keys.alloc((size_t)m);
            k_gather_regions(st, m, listCol.p, regMax.p, keys.p, counts.p);
            int bits = 1; while ((1ll << bits) < R + 1 && bits < 31) ++bits;
            sort_pairs_by_key(st, m, bits, keys, listCol, keysTmp, valsTmp);  // stable: (region, column) order
            k_positions(st, m, listCol.p, posOfCol.p);
        }
        std::vector<int> hc = counts.to_host(st, (size_t)R + 1);
        if (hc[R]) throw Error("explicit A: a DOF column is touched by two reduced regions (fixReducedRegionBoundaries should prevent this)");
        for (int r = 0; r < R; ++r) hStart[r + 1] = hStart[r] + hc[r];
    }
    regStart.from_host(st, hStart.data(), hStart.size());
    if (m == 0) listCol.alloc(1);
    VW.alloc((size_t)std::max<int64_t>(m, 1) * 2 * RDOF);
    ExplCtx X = {A, RG.rowRegion.p, RG.rowXYZ.p, RG.com.p, RG.Binv.p, g.dx, g.dt, regMax.p, regStart.p, listCol.p, posOfCol.p, VW.p};
    if (m > 0) k_region_columns_vw(st, m, X, VW.p);
    len.alloc((size_t)n + 1); len.zero(st, (size_t)n + 1);
    k_rows<false>(st, X, len.p, nullptr, nullptr, nullptr);
    const int64_t nnz = exclusive_scan_i64(st, n + 1, len.p, Aexp.ptr.p);
    if (nnz > (int64_t)INT32_MAX) throw Error("explicit A: more than 2^31-1 stored entries (Eigen's StorageIndex is int); use the factored solvers at this size");
    Aexp.nnz = nnz;
    Aexp.idx.alloc((size_t)std::max<int64_t>(nnz, 1)); Aexp.val.alloc((size_t)std::max<int64_t>(nnz, 1));
    k_rows<true>(st, X, nullptr, Aexp.ptr.p, Aexp.idx.p, Aexp.val.p);
    stream_sync(st);
    haveA = true;
}

}  // namespace ps
