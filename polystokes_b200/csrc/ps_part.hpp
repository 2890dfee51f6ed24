// ps_part.hpp -- z-slab decomposition of one Stokes step over several GPUs.
//
// One process per GPU.  The grid is cut along z at multiples of lcm(16, tileSize): the reference numbers DOFs
// in UT_VoxelArray order (16^3 tiles, tile index z-slowest; exec/HDK_PolyStokesSolver_Classifier.cpp:1738-1770),
// so every rank owns ONE contiguous index range per sample slot, and reduced tiles (aligned to multiples of
// tileSize, S_Cls:736-738) never straddle a cut.  Classification and the (cheap, stencil) matrix fill are
// computed redundantly by every rank in the global numbering -- no setup communication, bit-identical to the
// single-GPU result -- while the per-region Gram builds, the operator rows and all CG vector work are
// restricted to the owned ranges.  Per CG iteration: one halo exchange of p, one of w, two scalar all-reduces.
#pragma once
#include "ps_grid.hpp"

namespace ps {

// up to 7 disjoint ascending ranges of a global index space, passed by value to kernels: local item l in
// [0, total) <-> global index.  Neighbouring threads map to neighbouring indices except at the <= 6 seams.
template <int N>
struct RangeSetN {
    int n = 0;
    int64_t lo[N] = {};
    int64_t pre[N + 1] = {};     // pre[k] = items before range k; pre[n] = total
    void add(int64_t a, int64_t b) { if (b < a) b = a; lo[n] = a; pre[n + 1] = pre[n] + (b - a); ++n; }
    PS_HD int64_t total() const { return pre[n]; }
    PS_HD int64_t count(int k) const { return pre[k + 1] - pre[k]; }
    PS_HD int64_t at(int64_t l) const {
        int64_t base = lo[0];
#pragma unroll
        for (int i = 1; i < N; ++i) if (i < n && l >= pre[i]) base = lo[i] - pre[i];
        return base + l;
    }
    PS_HD bool has(int64_t gidx) const {
        bool in = false;
#pragma unroll
        for (int i = 0; i < N; ++i) if (i < n && gidx >= lo[i] && gidx < lo[i] + (pre[i + 1] - pre[i])) in = true;
        return in;
    }
};
typedef RangeSetN<7> RangeSet;    // the system vector: p | xx | yy | zz | yz | xz | xy
typedef RangeSetN<4> RowSet;      // operator rows: K_ext (x, y, z faces + coupled reduced rows) and the K_ext^T blocks
inline RowSet one_range(int64_t a, int64_t b) { RowSet r; r.add(a, b); return r; }

// host description of who owns what (identical on every rank: derived from the replicated classification)
struct Partition {
    int rank = 0, nranks = 1;
    std::vector<int> zCut;                              // [nranks+1]; rank k owns cells z in [zCut[k], zCut[k+1])
    std::vector<int64_t> slotCut[N_SLOTS];              // [nranks+1] active-index cut of every sample slot
    std::vector<int32_t> regionCut;                     // [nranks+1] region-id cut
    std::vector<int64_t> redRowCut;                     // [nranks+1] coupled-reduced-row cut (relative to nActiveVs)
    bool multi() const { return nranks > 1; }
};

}  // namespace ps
