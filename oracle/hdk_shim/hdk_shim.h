// hdk_shim.h -- TEST INFRASTRUCTURE (parity oracle, `make -C oracle ref`).  NOT the HDK.
//
// The reference's classifier (exec/HDK_PolyStokesSolver_Classifier.cpp) is written against SideFX's closed-source Houdini
// Development Kit.  This header is the "minimal shim for the HDK voxel-array types" of BASELINE.md section 3: it declares,
// from scratch, just the HDK names that exec/HDK_PolyStokesSolver.h, exec/HDK_PolyStokes.h and the classifier touch, with
// the semantics tabulated there, so that the reference's OWN classifier source compiles unmodified into
// oracle/_ref/libps_ref_classify.so.  What is the reference's: every flood, stencil rule, sweep order, remap and numbering
// loop of the classifier.  What is this shim's (and therefore the same definition the oracle restates): the 16^3 tile
// iteration order of UT_VoxelArray, the connected-component labelling of SIM_VolumetricConnectedComponentBuilder, the
// SIM::FieldUtils index maps, border modes, and threading (HDK_SHIM_THREADS jobs on std::thread; default one job, serial).
#pragma once
#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <ctime>
#include <deque>
#include <functional>
#include <limits>
#include <map>
#include <string>
#include <thread>
#include <vector>

typedef int64_t exint;
typedef double fpreal;
typedef float fpreal32;
typedef double fpreal64;
typedef unsigned int uint;

template <class T> inline T SYSclamp(T v, T lo, T hi) { return v < lo ? lo : (v > hi ? hi : v); }
template <class T> inline T SYSmin(T a, T b) { return a < b ? a : b; }
template <class T> inline T SYSmax(T a, T b) { return a > b ? a : b; }

// ---- parameters of the node (read by the GET_DATA_FUNC_* getters of HDK_PolyStokes.h) ----
namespace hdk_shim {
inline std::map<std::string, double>& params() { static std::map<std::string, double> p; return p; }
inline double param(const char* name) { std::map<std::string, double>::const_iterator it = params().find(name); assert(it != params().end()); return it == params().end() ? 0. : it->second; }
}

class UT_String { public: std::string s; UT_String() {} UT_String(const char* c) : s(c ? c : "") {} const char* c_str() const { return s.c_str(); } operator const char*() const { return s.c_str(); } std::string toStdString() const { return s; } };
typedef UT_String UT_StringHolder;

// ---- small vectors ----
template <class T>
class UT_Vector3T {
public:
    T v[3];
    UT_Vector3T() { v[0] = v[1] = v[2] = T(0); }
    UT_Vector3T(T a, T b, T c) { v[0] = a; v[1] = b; v[2] = c; }
    explicit UT_Vector3T(T a) { v[0] = v[1] = v[2] = a; }
    template <class U> UT_Vector3T(const UT_Vector3T<U>& o) { v[0] = (T)o.v[0]; v[1] = (T)o.v[1]; v[2] = (T)o.v[2]; }
    T& x() { return v[0]; } T& y() { return v[1]; } T& z() { return v[2]; }
    const T& x() const { return v[0]; } const T& y() const { return v[1]; } const T& z() const { return v[2]; }
    T& operator[](int i) { return v[i]; } const T& operator[](int i) const { return v[i]; }
    T& operator()(int i) { return v[i]; } const T& operator()(int i) const { return v[i]; }
    bool operator==(const UT_Vector3T& o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2]; }
    bool operator!=(const UT_Vector3T& o) const { return !(*this == o); }
    UT_Vector3T operator+(const UT_Vector3T& o) const { return UT_Vector3T(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
    UT_Vector3T operator-(const UT_Vector3T& o) const { return UT_Vector3T(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
    UT_Vector3T& operator+=(const UT_Vector3T& o) { v[0] += o.v[0]; v[1] += o.v[1]; v[2] += o.v[2]; return *this; }
    UT_Vector3T& operator-=(const UT_Vector3T& o) { v[0] -= o.v[0]; v[1] -= o.v[1]; v[2] -= o.v[2]; return *this; }
    UT_Vector3T& operator*=(T s) { v[0] *= s; v[1] *= s; v[2] *= s; return *this; }
    UT_Vector3T& operator/=(T s) { v[0] /= s; v[1] /= s; v[2] /= s; return *this; }
    UT_Vector3T operator*(T s) const { return UT_Vector3T(v[0] * s, v[1] * s, v[2] * s); }
    UT_Vector3T operator/(T s) const { return UT_Vector3T(v[0] / s, v[1] / s, v[2] / s); }
    T dot(const UT_Vector3T& o) const { return v[0] * o.v[0] + v[1] * o.v[1] + v[2] * o.v[2]; }
    T length2() const { return dot(*this); }
};
template <class T> UT_Vector3T<T> operator*(T s, const UT_Vector3T<T>& a) { return a * s; }
typedef UT_Vector3T<fpreal32> UT_Vector3;
typedef UT_Vector3T<fpreal32> UT_Vector3F;
typedef UT_Vector3T<fpreal64> UT_Vector3D;
typedef UT_Vector3T<int32_t> UT_Vector3i;
typedef UT_Vector3T<int64_t> UT_Vector3I;

// ---- UT_Array: the growable array (deque-backed so that UT_Array<bool> hands out real references) ----
template <class T>
class UT_Array {
public:
    typedef typename std::deque<T>::iterator iterator;
    typedef typename std::deque<T>::const_iterator const_iterator;
    UT_Array() {}
    explicit UT_Array(exint capacity, exint entries = 0) { d.resize((size_t)entries); (void)capacity; }
    exint size() const { return (exint)d.size(); }
    exint entries() const { return (exint)d.size(); }
    bool isEmpty() const { return d.empty(); }
    void setSize(exint n) { d.resize((size_t)n); }
    void setSizeNoInit(exint n) { d.resize((size_t)n); }
    void setCapacity(exint) {}
    void bumpCapacity(exint) {}
    exint capacity() const { return (exint)d.size(); }
    void clear() { d.clear(); }
    void constant(const T& v) { std::fill(d.begin(), d.end(), v); }
    exint append(const T& v) { d.push_back(v); return (exint)d.size() - 1; }
    void concat(const UT_Array<T>& o) { d.insert(d.end(), o.d.begin(), o.d.end()); }
    T& operator[](exint i) { assert(i >= 0 && i < size()); return d[(size_t)i]; }
    const T& operator[](exint i) const { assert(i >= 0 && i < size()); return d[(size_t)i]; }
    T& operator()(exint i) { return (*this)[i]; }
    const T& operator()(exint i) const { return (*this)[i]; }
    T& last() { return d.back(); }
    const T& last() const { return d.back(); }
    iterator begin() { return d.begin(); } iterator end() { return d.end(); }
    const_iterator begin() const { return d.begin(); } const_iterator end() const { return d.end(); }
private:
    std::deque<T> d;
};

// ---- threading: HDK_SHIM_THREADS jobs (default 1: everything serial and in order, which is what the bit-exact parity tests use) ----
// The reference splits its voxel sweeps by tile over UT_ThreadedAlgorithm jobs (per-job accumulators indexed by info.job(), merged in job
// order, e.g. S_CMB:324-341, S.cpp:1285-1324), its per-region work over tbb::parallel_for and a few loops over UTparallelFor*.  The stand-in
// runs job j of n on its own std::thread with the SAME contiguous split the single job would walk in order, so list-building stages give
// identical results for any n; floating-point sums that are accumulated per job (centres of mass, Gram matrices) differ in the last bits
// between different n, exactly as they do between machines with different core counts in Houdini.  bench.py sets the job count to the
// number of host cores (reffull_set_threads); tests leave it at 1.
#include "shim_threads.h"
class UT_JobInfo {
public:
    UT_JobInfo() {}
    UT_JobInfo(int j, int n) : myJob(j), myNum(n) {}
    int job() const { return myJob; }
    int numJobs() const { return myNum; }
    void divideWork(exint units, exint& start, exint& end) const { start = units * myJob / myNum; end = units * (myJob + 1) / myNum; }
    void divideWork(int units, int& start, int& end) const { start = (int)((int64_t)units * myJob / myNum); end = (int)((int64_t)units * (myJob + 1) / myNum); }
private:
    int myJob = 0, myNum = 1;
};
class UT_ThreadedAlgorithm {
public:
    template <class F> void run(F f) { const int n = hdk_shim::threads(); hdk_shim::runJobs(n, [&](int j) { UT_JobInfo info(j, n); f(info); }); }
};
class UT_Thread { public: static int getNumProcessors() { return hdk_shim::threads(); } };
class UT_Interrupt { public: bool opInterrupt(int = -1) { return false; } bool opStart(const char* = nullptr) { return true; } void opEnd() {} };
inline UT_Interrupt* UTgetInterrupt() { static UT_Interrupt boss; return &boss; }
class UT_AutoInterrupt { public: explicit UT_AutoInterrupt(const char*) {} bool wasInterrupted() { return false; } };
template <class T>
class UT_BlockedRange { public: UT_BlockedRange(T b, T e, size_t = 1) : b(b), e(e) {} T begin() const { return b; } T end() const { return e; } private: T b, e; };
template <class I, class F> inline void UTparallelForEachNumber(I n, const F& f) { hdk_shim::forRange(I(0), n, [&](I lo, I hi) { f(UT_BlockedRange<I>(lo, hi)); }); }
template <class R, class F> inline void UTparallelFor(const R& r, const F& f, int = 0, int = 0) { hdk_shim::forRange(r.begin(), r.end(), [&](decltype(r.begin()) lo, decltype(r.begin()) hi) { f(R(lo, hi)); }); }
template <class R, class F> inline void UTparallelForLightItems(const R& r, const F& f) { UTparallelFor(r, f); }
template <class R, class F> inline void UTparallelForHeavyItems(const R& r, const F& f) { UTparallelFor(r, f); }
template <class R, class F> inline void UTserialFor(const R& r, const F& f) { f(r); }
template <class It, class C> inline void UTparallelSort(It b, It e, C c) { std::sort(b, e, c); }
template <class It> inline void UTparallelSort(It b, It e) { std::sort(b, e); }
template <class It, class C> inline void UTparallelStableSort(It b, It e, C c) { std::stable_sort(b, e, c); }

// THREADED_METHODn(CLASS, DOMULTI, METHOD, types/names...): METHOD(args) runs METHODPartial(args, info) once per job
#define HDK_SHIM_JOBS(CALL) do { const int n_ = hdk_shim::threads(); hdk_shim::runJobs(n_, [&](int j_) { const UT_JobInfo info(j_, n_); CALL; }); } while (0)
#define THREADED_METHOD(C, DOMULTI, M) void M() { HDK_SHIM_JOBS(M##Partial(info)); }
#define THREADED_METHOD1(C, DOMULTI, M, T1, P1) void M(T1 P1) { HDK_SHIM_JOBS(M##Partial(P1, info)); }
#define THREADED_METHOD2(C, DOMULTI, M, T1, P1, T2, P2) void M(T1 P1, T2 P2) { HDK_SHIM_JOBS(M##Partial(P1, P2, info)); }
#define THREADED_METHOD3(C, DOMULTI, M, T1, P1, T2, P2, T3, P3) void M(T1 P1, T2 P2, T3 P3) { HDK_SHIM_JOBS(M##Partial(P1, P2, P3, info)); }
#define THREADED_METHOD4(C, DOMULTI, M, T1, P1, T2, P2, T3, P3, T4, P4) void M(T1 P1, T2 P2, T3 P3, T4 P4) { HDK_SHIM_JOBS(M##Partial(P1, P2, P3, P4, info)); }

// ---- UT_VoxelArray: dense storage, logical 16^3 tiles (partial at the high edges), tile-linear index tx fastest ----
enum UT_VoxelBorderType { UT_VOXELBORDER_CONSTANT, UT_VOXELBORDER_REPEAT, UT_VOXELBORDER_STREAK, UT_VOXELBORDER_EXTRAP };
const int UT_VOXEL_TILE = 16;

template <class T> class UT_VoxelArray;
template <class T>
class UT_VoxelTile {        // a view of one tile; "constant" is a flag: the data always exists
public:
    UT_VoxelArray<T>* arr = nullptr; int t = 0;
    bool isConstant() const;
    void makeConstant(T v);
    void uncompress();
    bool tryCompress() { return false; }
    int xres() const; int yres() const; int zres() const;
};

template <class T>
class UT_VoxelArray {
public:
    UT_VoxelArray() {}
    void size(int rx, int ry, int rz) {
        r[0] = rx; r[1] = ry; r[2] = rz;
        for (int a = 0; a < 3; ++a) nt[a] = (r[a] + UT_VOXEL_TILE - 1) / UT_VOXEL_TILE;
        d.assign((size_t)rx * ry * rz, T(0));
        isConst.assign((size_t)numTiles(), 1);
        tiles.resize((size_t)numTiles());
        for (int i = 0; i < numTiles(); ++i) { tiles[(size_t)i].arr = this; tiles[(size_t)i].t = i; }
    }
    int getXRes() const { return r[0]; } int getYRes() const { return r[1]; } int getZRes() const { return r[2]; }
    int getRes(int a) const { return r[a]; }
    UT_Vector3I getVoxelRes() const { return UT_Vector3I(r[0], r[1], r[2]); }
    int numTiles() const { return nt[0] * nt[1] * nt[2]; }
    int getTileRes(int a) const { return nt[a]; }
    int indexToLinearTile(int x, int y, int z) const { return (x / UT_VOXEL_TILE) + nt[0] * ((y / UT_VOXEL_TILE) + nt[1] * (z / UT_VOXEL_TILE)); }
    void linearTileToXYZ(int t, int& tx, int& ty, int& tz) const { tx = t % nt[0]; ty = (t / nt[0]) % nt[1]; tz = t / (nt[0] * nt[1]); }
    UT_VoxelTile<T>* getLinearTile(int t) const { return const_cast<UT_VoxelTile<T>*>(&tiles[(size_t)t]); }
    bool isValidIndex(int x, int y, int z) const { return x >= 0 && x < r[0] && y >= 0 && y < r[1] && z >= 0 && z < r[2]; }
    size_t lin(int x, int y, int z) const { return (size_t)x + (size_t)r[0] * ((size_t)y + (size_t)r[1] * (size_t)z); }
    // reads outside the array follow the border mode: constant value, or the nearest voxel (STREAK, the HDK default of float fields)
    T getValue(int x, int y, int z) const {
        if (!isValidIndex(x, y, z)) {
            if (border == UT_VOXELBORDER_CONSTANT) return borderValue;
            x = SYSclamp(x, 0, r[0] - 1); y = SYSclamp(y, 0, r[1] - 1); z = SYSclamp(z, 0, r[2] - 1);
        }
        return d[lin(x, y, z)];
    }
    T operator()(int x, int y, int z) const { return getValue(x, y, z); }
    T operator()(const UT_Vector3I& p) const { return getValue((int)p[0], (int)p[1], (int)p[2]); }
    void setValue(int x, int y, int z, T v) {
        assert(isValidIndex(x, y, z));
        T& s = d[lin(x, y, z)];
        if (!(s == v)) isConst[(size_t)indexToLinearTile(x, y, z)] = 0;      // writing a different value uncompresses the tile
        s = v;
    }
    void setValue(const UT_Vector3I& p, T v) { setValue((int)p[0], (int)p[1], (int)p[2], v); }
    void constant(T v) { std::fill(d.begin(), d.end(), v); std::fill(isConst.begin(), isConst.end(), 1); }
    void setBorder(UT_VoxelBorderType t, T v) { border = t; borderValue = v; }
    void collapseAllTiles() { for (int t = 0; t < numTiles(); ++t) { bool same = true; T v0 = T(0); bool first = true; forTile(t, [&](int x, int y, int z) { T v = d[lin(x, y, z)]; if (first) { v0 = v; first = false; } else if (!(v == v0)) same = false; }); isConst[(size_t)t] = same ? 1 : 0; } }
    void expandAllTiles() { std::fill(isConst.begin(), isConst.end(), 0); }
    template <class F> void forTile(int t, F f) const {
        int tx, ty, tz; linearTileToXYZ(t, tx, ty, tz);
        const int x0 = tx * UT_VOXEL_TILE, y0 = ty * UT_VOXEL_TILE, z0 = tz * UT_VOXEL_TILE;
        const int x1 = std::min(x0 + UT_VOXEL_TILE, r[0]), y1 = std::min(y0 + UT_VOXEL_TILE, r[1]), z1 = std::min(z0 + UT_VOXEL_TILE, r[2]);
        for (int z = z0; z < z1; ++z) for (int y = y0; y < y1; ++y) for (int x = x0; x < x1; ++x) f(x, y, z);
    }
    int r[3] = {0, 0, 0}, nt[3] = {0, 0, 0};
    std::vector<T> d;
    std::vector<char> isConst;
    std::vector<UT_VoxelTile<T> > tiles;
    UT_VoxelBorderType border = UT_VOXELBORDER_STREAK;
    T borderValue = T(0);
};
template <class T> bool UT_VoxelTile<T>::isConstant() const { return arr->isConst[(size_t)t] != 0; }
template <class T> void UT_VoxelTile<T>::makeConstant(T v) { UT_VoxelArray<T>* a = arr; a->forTile(t, [&](int x, int y, int z) { a->d[a->lin(x, y, z)] = v; }); a->isConst[(size_t)t] = 1; }
template <class T> void UT_VoxelTile<T>::uncompress() { arr->isConst[(size_t)t] = 0; }
typedef UT_VoxelArray<fpreal32> UT_VoxelArrayF;
typedef UT_VoxelArray<exint> UT_VoxelArrayI;
typedef UT_VoxelTile<fpreal32> UT_VoxelTileF;

// iterator over the voxels of tiles [myTileStart, myTileEnd): tiles in linear order, voxels x fastest inside a tile
template <class T>
class UT_VoxelArrayIterator {
public:
    UT_VoxelArrayIterator() {}
    explicit UT_VoxelArrayIterator(UT_VoxelArray<T>* a) { setArray(a); }
    void setArray(UT_VoxelArray<T>* a) { arr = a; myTileStart = 0; myTileEnd = a->numTiles(); }
    void setConstArray(const UT_VoxelArray<T>* a) { setArray(const_cast<UT_VoxelArray<T>*>(a)); }
    void setCompressOnExit(bool) {}
    void setPartialRange(int idx, int num) { const int n = arr->numTiles(); myTileStart = (int)((int64_t)n * idx / num); myTileEnd = (int)((int64_t)n * (idx + 1) / num); }
    void splitByTile(const UT_JobInfo& info) { setPartialRange(info.job(), info.numJobs()); }
    void rewind() { tile = myTileStart; enterTile(); }
    bool atEnd() const { return tile >= myTileEnd; }
    void advance() {
        if (++cx < x1) return;
        cx = x0; if (++cy < y1) return;
        cy = y0; if (++cz < z1) return;
        ++tile; enterTile();
    }
    void advanceTile() { ++tile; enterTile(); }
    int x() const { return cx; } int y() const { return cy; } int z() const { return cz; }
    int idx(int a) const { return a == 0 ? cx : (a == 1 ? cy : cz); }
    T getValue() const { return arr->d[arr->lin(cx, cy, cz)]; }
    void setValue(T v) const { arr->setValue(cx, cy, cz, v); }
    bool isTileConstant() const { return arr->isConst[(size_t)tile] != 0; }
    bool isStartOfTile() const { return cx == x0 && cy == y0 && cz == z0; }
    UT_VoxelTile<T>* getTile() const { return arr->getLinearTile(tile); }
    int getLinearTileNum() const { return tile; }
    int myTileStart = 0, myTileEnd = 0;
    UT_VoxelArray<T>* arr = nullptr;
    int tile = 0, x0 = 0, y0 = 0, z0 = 0, x1 = 0, y1 = 0, z1 = 0, cx = 0, cy = 0, cz = 0;
private:
    void enterTile() {
        if (tile >= myTileEnd) return;
        int tx, ty, tz; arr->linearTileToXYZ(tile, tx, ty, tz);
        x0 = tx * UT_VOXEL_TILE; y0 = ty * UT_VOXEL_TILE; z0 = tz * UT_VOXEL_TILE;
        x1 = std::min(x0 + UT_VOXEL_TILE, arr->r[0]); y1 = std::min(y0 + UT_VOXEL_TILE, arr->r[1]); z1 = std::min(z0 + UT_VOXEL_TILE, arr->r[2]);
        cx = x0; cy = y0; cz = z0;
    }
};
// iterator over the voxels of ONE tile (the current tile of an array iterator)
template <class T>
class UT_VoxelTileIterator {
public:
    void setTile(const UT_VoxelArrayIterator<T>& vit) { arr = vit.arr; x0 = vit.x0; y0 = vit.y0; z0 = vit.z0; x1 = vit.x1; y1 = vit.y1; z1 = vit.z1; done = false; cx = x0; cy = y0; cz = z0; }
    void rewind() { cx = x0; cy = y0; cz = z0; done = !(x0 < x1 && y0 < y1 && z0 < z1); }
    bool atEnd() const { return done; }
    void advance() {
        if (++cx < x1) return;
        cx = x0; if (++cy < y1) return;
        cy = y0; if (++cz < z1) return;
        done = true;
    }
    int x() const { return cx; } int y() const { return cy; } int z() const { return cz; }
    T getValue() const { return arr->d[arr->lin(cx, cy, cz)]; }
    void setValue(T v) const { arr->setValue(cx, cy, cz, v); }
private:
    UT_VoxelArray<T>* arr = nullptr;
    int x0 = 0, y0 = 0, z0 = 0, x1 = 0, y1 = 0, z1 = 0, cx = 0, cy = 0, cz = 0;
    bool done = true;
};
typedef UT_VoxelArrayIterator<fpreal32> UT_VoxelArrayIteratorF;
typedef UT_VoxelArrayIterator<exint> UT_VoxelArrayIteratorI;
typedef UT_VoxelTileIterator<fpreal32> UT_VoxelTileIteratorF;
typedef UT_VoxelTileIterator<exint> UT_VoxelTileIteratorI;

// ---- SIM fields ----
enum SIM_FieldSample { SIM_SAMPLE_CENTER = 0, SIM_SAMPLE_EDGEXY = 1, SIM_SAMPLE_EDGEXZ = 2, SIM_SAMPLE_EDGEYZ = 3, SIM_SAMPLE_FACEX = 4, SIM_SAMPLE_FACEY = 5, SIM_SAMPLE_FACEZ = 6, SIM_SAMPLE_CORNER = 7 };
enum { SIM_DATA_ID_PRESERVE = 0 };
#define SIM_NAME_TOLERANCE "tolerance"

namespace hdk_shim {
// voxel resolution of a sample position on an nx x ny x nz cell grid
inline void sampleRes(SIM_FieldSample s, int nx, int ny, int nz, int out[3]) {
    out[0] = nx; out[1] = ny; out[2] = nz;
    switch (s) {
    case SIM_SAMPLE_FACEX: out[0] += 1; break;
    case SIM_SAMPLE_FACEY: out[1] += 1; break;
    case SIM_SAMPLE_FACEZ: out[2] += 1; break;
    case SIM_SAMPLE_EDGEXY: out[0] += 1; out[1] += 1; break;
    case SIM_SAMPLE_EDGEXZ: out[0] += 1; out[2] += 1; break;
    case SIM_SAMPLE_EDGEYZ: out[1] += 1; out[2] += 1; break;
    case SIM_SAMPLE_CORNER: out[0] += 1; out[1] += 1; out[2] += 1; break;
    default: break;
    }
}
}

template <class T>
class SIM_RawFieldT {
public:
    typedef UT_VoxelArray<T> Array;
    SIM_RawFieldT() {}
    void init(SIM_FieldSample s, const UT_Vector3& orig, const UT_Vector3& size, int nx, int ny, int nz) {
        mySample = s; myOrig = orig; mySize = size; cells[0] = nx; cells[1] = ny; cells[2] = nz;
        int r[3]; hdk_shim::sampleRes(s, nx, ny, nz, r);
        arr.size(r[0], r[1], r[2]);
    }
    template <class U> void match(const SIM_RawFieldT<U>& o) { const UT_VoxelBorderType b = arr.border; const T bv = arr.borderValue; init(o.getSample(), o.getOrig(), o.getSize(), o.cells[0], o.cells[1], o.cells[2]); arr.border = b; arr.borderValue = bv; }
    const Array* field() const { return &arr; }
    Array* fieldNC() const { return const_cast<Array*>(&arr); }
    void makeConstant(T v) { arr.constant(v); }
    void setBorder(UT_VoxelBorderType t, T v) { arr.setBorder(t, v); }
    SIM_FieldSample getSample() const { return mySample; }
    const UT_Vector3& getOrig() const { return myOrig; }
    const UT_Vector3& getSize() const { return mySize; }
    void getVoxelRes(int& x, int& y, int& z) const { x = arr.r[0]; y = arr.r[1]; z = arr.r[2]; }
    UT_Vector3I getVoxelRes() const { return arr.getVoxelRes(); }
    T getCellValue(int x, int y, int z) const { return arr.getValue(x, y, z); }
    T operator()(int x, int y, int z) const { return arr.getValue(x, y, z); }       // reads follow the border mode
    bool shouldMultiThread() const { return hdk_shim::threads() > 1; }
    bool isMatching(const SIM_RawFieldT&) const { return true; }
    // position of sample (x, y, z): cell corner + sample offset, in the float precision of the HDK
    bool indexToPos(int x, int y, int z, UT_Vector3& pos) const {
        const float off[3] = {(mySample == SIM_SAMPLE_CENTER || mySample == SIM_SAMPLE_FACEY || mySample == SIM_SAMPLE_FACEZ || mySample == SIM_SAMPLE_EDGEYZ) ? 0.5f : 0.f,
                              (mySample == SIM_SAMPLE_CENTER || mySample == SIM_SAMPLE_FACEX || mySample == SIM_SAMPLE_FACEZ || mySample == SIM_SAMPLE_EDGEXZ) ? 0.5f : 0.f,
                              (mySample == SIM_SAMPLE_CENTER || mySample == SIM_SAMPLE_FACEX || mySample == SIM_SAMPLE_FACEY || mySample == SIM_SAMPLE_EDGEXY) ? 0.5f : 0.f};
        const int idx[3] = {x, y, z};
        for (int a = 0; a < 3; ++a) pos[a] = myOrig[a] + ((float)idx[a] + off[a]) * (mySize[a] / (float)cells[a]);
        return true;
    }
    // SIM_RawField::getValue(pos) (closed source): trilinear sample; every position the reference asks for is
    // a sample position of the same grid, so the index-space coordinates are snapped to multiples of 1/2 (fractions exactly 0 or
    // 1/2); lerp along x, then y, then z as a + t (b - a) in the field's precision; reads outside follow the border mode.
    T getValue(const UT_Vector3& pos) const {
        const float off[3] = {(mySample == SIM_SAMPLE_CENTER || mySample == SIM_SAMPLE_FACEY || mySample == SIM_SAMPLE_FACEZ || mySample == SIM_SAMPLE_EDGEYZ) ? 0.5f : 0.f,
                              (mySample == SIM_SAMPLE_CENTER || mySample == SIM_SAMPLE_FACEX || mySample == SIM_SAMPLE_FACEZ || mySample == SIM_SAMPLE_EDGEXZ) ? 0.5f : 0.f,
                              (mySample == SIM_SAMPLE_CENTER || mySample == SIM_SAMPLE_FACEX || mySample == SIM_SAMPLE_FACEY || mySample == SIM_SAMPLE_EDGEXY) ? 0.5f : 0.f};
        int base[3]; T t[3];
        for (int a = 0; a < 3; ++a) {
            const double u = ((double)pos[a] - (double)myOrig[a]) / ((double)mySize[a] / (double)cells[a]) - (double)off[a];
            const long long twice = std::llround(2. * u);
            base[a] = (int)std::floor((double)twice / 2.);
            t[a] = (T)(0.5 * (double)(twice - 2ll * base[a]));
        }
        T cz[2];
        for (int dz = 0; dz < 2; ++dz) {
            T cy[2];
            for (int dy = 0; dy < 2; ++dy) {
                const T c0 = arr.getValue(base[0], base[1] + dy, base[2] + dz), c1 = arr.getValue(base[0] + 1, base[1] + dy, base[2] + dz);
                cy[dy] = c0 + t[0] * (c1 - c0);
            }
            cz[dz] = cy[0] + t[1] * (cy[1] - cy[0]);
        }
        return cz[0] + t[2] * (cz[1] - cz[0]);
    }
    Array arr;
    SIM_FieldSample mySample = SIM_SAMPLE_CENTER;
    UT_Vector3 myOrig, mySize;
    int cells[3] = {0, 0, 0};
};
class SIM_RawField : public SIM_RawFieldT<fpreal32> {
public:
    // SIM_RawField::computeSDFWeightsSampled (closed source; semantics DEFINED in BASELINE.md section 3): weight of a sample =
    // (1/8) * #{2 x 2 x 2 sub-sample points at +-1/4 voxel around the sample position where the trilinearly interpolated SDF is < 0}.
    // The five-argument form is the reference's "fluid" (non-solid) weight, S.cpp:322-323: 1 in fluid, 0 in solid (S_Cls:106-107),
    // i.e. the test is "outside the collision SDF" (>= 0).  Interpolation fractions are exactly 1/4 or 3/4; the trilinear sum is
    // evaluated in double, where every product is exact.  SDF reads outside the grid clamp to the edge.
    void computeSDFWeightsSampled(const SIM_RawField* sdf, int samplesperaxis, bool invert, fpreal minweight) { sampleWeights(*sdf, invert); (void)samplesperaxis; (void)minweight; }
    void computeSDFWeightsSampled(const SIM_RawField* sdf, int samplesperaxis, bool invert, fpreal minweight, fpreal dilate) { sampleWeights(*sdf, !invert); (void)samplesperaxis; (void)minweight; (void)dilate; }
    void setScaleDivideThreshold(fpreal, const SIM_RawField*, const SIM_RawField*, fpreal) {}
private:
    void sampleWeights(const SIM_RawField& sdf, bool countNonNegative) {
        const int off2[3] = {(mySample == SIM_SAMPLE_CENTER || mySample == SIM_SAMPLE_FACEY || mySample == SIM_SAMPLE_FACEZ || mySample == SIM_SAMPLE_EDGEYZ) ? 1 : 0,
                             (mySample == SIM_SAMPLE_CENTER || mySample == SIM_SAMPLE_FACEX || mySample == SIM_SAMPLE_FACEZ || mySample == SIM_SAMPLE_EDGEXZ) ? 1 : 0,
                             (mySample == SIM_SAMPLE_CENTER || mySample == SIM_SAMPLE_FACEX || mySample == SIM_SAMPLE_FACEY || mySample == SIM_SAMPLE_EDGEXY) ? 1 : 0};   // sample offset in half voxels
        const UT_VoxelArrayF& S = *sdf.field();
        // z-planes split over the jobs (the HDK's own routine is threaded by tile); every voxel is independent
        hdk_shim::forRange(0, arr.r[2], [&](int k0, int k1) {
        for (int k = k0; k < k1; ++k) for (int j = 0; j < arr.r[1]; ++j) for (int i = 0; i < arr.r[0]; ++i) {
            const int idx[3] = {i, j, k};
            int count = 0;
            for (int sub = 0; sub < 8; ++sub) {
                int base[3]; double fr[3];
                for (int a = 0; a < 3; ++a) {
                    const int q = 4 * idx[a] + 2 * off2[a] - 2 + (((sub >> a) & 1) ? 1 : -1);      // quarter-voxel coordinate relative to the cell centres
                    const int fl = (q >= 0) ? q / 4 : -((-q + 3) / 4);
                    base[a] = fl; fr[a] = (q - 4 * fl) * 0.25;
                }
                double acc = 0.;
                if (base[0] >= 0 && base[1] >= 0 && base[2] >= 0 && base[0] + 1 < S.r[0] && base[1] + 1 < S.r[1] && base[2] + 1 < S.r[2]) {
                    // all eight corners inside the grid: same sum, same order, without the border handling of getValue
                    const float* c = S.d.data() + S.lin(base[0], base[1], base[2]);
                    const size_t sy = (size_t)S.r[0], sz = (size_t)S.r[0] * (size_t)S.r[1];
                    for (int dz = 0; dz < 2; ++dz) for (int dy = 0; dy < 2; ++dy) for (int dxx = 0; dxx < 2; ++dxx) {
                        const double w = (dxx ? fr[0] : 1. - fr[0]) * (dy ? fr[1] : 1. - fr[1]) * (dz ? fr[2] : 1. - fr[2]);
                        acc += w * (double)c[(size_t)dxx + sy * (size_t)dy + sz * (size_t)dz];
                    }
                } else
                for (int dz = 0; dz < 2; ++dz) for (int dy = 0; dy < 2; ++dy) for (int dxx = 0; dxx < 2; ++dxx) {
                    const double w = (dxx ? fr[0] : 1. - fr[0]) * (dy ? fr[1] : 1. - fr[1]) * (dz ? fr[2] : 1. - fr[2]);
                    acc += w * (double)S.getValue(base[0] + dxx, base[1] + dy, base[2] + dz);
                }
                if (countNonNegative ? (acc >= 0.) : (acc < 0.)) ++count;
            }
            arr.d[arr.lin(i, j, k)] = (float)count * 0.125f;
        }
        });
        arr.expandAllTiles();
    }
};
class SIM_RawIndexField : public SIM_RawFieldT<exint> {};

class SIM_ScalarField { public: SIM_RawField* getField() const { return const_cast<SIM_RawField*>(&f); } SIM_RawField f; };
class SIM_VectorField {
public:
    SIM_RawField* getField(int a) const { return const_cast<SIM_RawField*>(&f[a]); }
    SIM_RawField* getXField() const { return getField(0); } SIM_RawField* getYField() const { return getField(1); } SIM_RawField* getZField() const { return getField(2); }
    UT_Vector3 getSize() const { return f[0].getSize(); }
    UT_Vector3 getOrig() const { return f[0].getOrig(); }
    UT_Vector3 getTotalVoxelRes() const { return UT_Vector3((float)f[0].cells[0], (float)f[0].cells[1], (float)f[0].cells[2]); }
    bool isFaceSampled() const { return true; }
    void pubHandleModification() {}
    SIM_RawField f[3];
};
class SIM_IndexField { public: SIM_RawIndexField* getField() const { return const_cast<SIM_RawIndexField*>(&f); } SIM_RawIndexField f; };

namespace SIM { namespace FieldUtils {
inline UT_Vector3I cellToFaceMap(UT_Vector3I c, int axis, int dir) { if (dir == 1) c[axis] += 1; return c; }
inline UT_Vector3I faceToCellMap(UT_Vector3I f, int axis, int dir) { if (dir == 0) f[axis] -= 1; return f; }
inline UT_Vector3I cellToCellMap(UT_Vector3I c, int axis, int dir) { c[axis] += (dir == 0) ? -1 : 1; return c; }
inline UT_Vector3I faceToEdgeMap(UT_Vector3I f, int faceAxis, int edgeAxis, int dir) { if (dir == 1) f[3 - faceAxis - edgeAxis] += 1; return f; }
inline UT_Vector3I edgeToFaceMap(UT_Vector3I e, int edgeAxis, int faceAxis, int dir) { if (dir == 0) e[3 - faceAxis - edgeAxis] -= 1; return e; }
template <class F> inline auto getFieldValue(const F& f, const UT_Vector3I& p) -> decltype(f.field()->getValue(0, 0, 0)) { return f.field()->getValue((int)p[0], (int)p[1], (int)p[2]); }
template <class F, class V> inline void setFieldValue(F& f, const UT_Vector3I& p, const V& v) { f.fieldNC()->setValue((int)p[0], (int)p[1], (int)p[2], v); }
} }

// ---- SIM_VolumetricConnectedComponentBuilder (closed source; semantics DEFINED in BASELINE.md section 3) ----
// 6-connected components of the cells whose label satisfies the predicate; two cells are connected only through a face with
// weight > 0; ids 0..R-1 in order of first encounter in the tile iteration order; every other cell gets INACTIVE_REGION.
class SIM_VolumetricConnectedComponentBuilder {
public:
    static const exint INACTIVE_REGION = -2;
    static const exint UNVISITED_REGION = -3;
    SIM_VolumetricConnectedComponentBuilder(SIM_RawIndexField& regions, const SIM_RawIndexField& labels, const SIM_RawField* const* faceWeights)
        : R(regions), L(labels) { for (int a = 0; a < 3; ++a) W[a] = faceWeights ? faceWeights[a] : nullptr; }
    template <class Pred>
    exint buildConnectedComponents(const Pred& pred) {
        using namespace SIM::FieldUtils;
        UT_VoxelArrayI* r = R.fieldNC(); const UT_VoxelArrayI* l = L.field();
        r->constant(UNVISITED_REGION);
        exint count = 0;
        std::vector<UT_Vector3I> stack;
        UT_VoxelArrayIteratorI vit; vit.setConstArray(l);
        for (vit.rewind(); !vit.atEnd(); vit.advance()) {
            const UT_Vector3I start(vit.x(), vit.y(), vit.z());
            if (!pred(vit.getValue())) { r->setValue(start, INACTIVE_REGION); continue; }
            if (r->getValue(vit.x(), vit.y(), vit.z()) != UNVISITED_REGION) continue;
            const exint id = count++;
            r->setValue(start, id); stack.push_back(start);
            while (!stack.empty()) {
                const UT_Vector3I c = stack.back(); stack.pop_back();
                for (int axis = 0; axis < 3; ++axis) for (int dir = 0; dir < 2; ++dir) {
                    const UT_Vector3I adj = cellToCellMap(c, axis, dir);
                    if (!l->isValidIndex((int)adj[0], (int)adj[1], (int)adj[2])) continue;
                    if (!pred(l->getValue((int)adj[0], (int)adj[1], (int)adj[2])) || r->getValue((int)adj[0], (int)adj[1], (int)adj[2]) != UNVISITED_REGION) continue;
                    const UT_Vector3I face = cellToFaceMap(c, axis, dir);
                    if (W[axis] && !(W[axis]->field()->getValue((int)face[0], (int)face[1], (int)face[2]) > 0.f)) continue;
                    r->setValue(adj, id); stack.push_back(adj);
                }
            }
        }
        return count;
    }
private:
    SIM_RawIndexField& R; const SIM_RawIndexField& L; const SIM_RawField* W[3];
};

// ---- node plumbing: only what exec/HDK_PolyStokes.h needs to be a complete class ----
typedef double SIM_Time;
class SIM_Engine {};
class SIM_Object {};
class SIM_Data {};
class SIM_DataFactory {};
class SIM_DopDescription {};
class SIM_Geometry {};
class SIM_GeometryCopy {};
// viewport dumps (printAllData, S.cpp:1030-1270): accepted and dropped
typedef exint GA_Offset;
enum GA_AttributeOwner { GA_ATTRIB_VERTEX, GA_ATTRIB_POINT, GA_ATTRIB_PRIMITIVE, GA_ATTRIB_GLOBAL };
class GA_Defaults { public: GA_Defaults(double) {} };
class GA_AttributeSet { public: void bumpAllDataIds(GA_AttributeOwner) {} };
class GU_Detail {
public:
    void clear() { n = 0; }
    void addFloatTuple(GA_AttributeOwner, const char*, int, const GA_Defaults&) {}
    GA_Offset appendPoint() { return n++; }
    GA_Offset appendPointBlock(exint k) { const GA_Offset o = n; n += k; return o; }
    void setPos3(GA_Offset, const UT_Vector3&) {}
    GA_AttributeSet& getAttributes() { return attrs; }
private:
    exint n = 0; GA_AttributeSet attrs;
};
class GA_RWHandleF {
public:
    GA_RWHandleF() {}
    GA_RWHandleF(GU_Detail*, GA_AttributeOwner, const char*) {}
    bool isValid() const { return true; }
    void bumpDataId() {}
    void set(GA_Offset, fpreal) {}
};
class SIM_GeometryAutoWriteLock { public: SIM_GeometryAutoWriteLock(SIM_GeometryCopy*, int) {} GU_Detail& getGdp() { return gdp; } private: GU_Detail gdp; };
class PRM_Template {};
class GAS_SubSolver {
public:
    explicit GAS_SubSolver(const SIM_DataFactory*) {}
    virtual ~GAS_SubSolver() {}
    SIM_GeometryCopy* getOrCreateGeometry(SIM_Object*, const char*) { static SIM_GeometryCopy g; return &g; }
protected:
    virtual bool solveGasSubclass(SIM_Engine&, SIM_Object*, SIM_Time, SIM_Time) = 0;
};
#define GET_DATA_FUNC_I(NAME, FN) int get##FN() const { return (int)hdk_shim::param(NAME); }
#define GET_DATA_FUNC_B(NAME, FN) bool get##FN() const { return hdk_shim::param(NAME) != 0.; }
#define GET_DATA_FUNC_F(NAME, FN) fpreal get##FN() const { return (fpreal)hdk_shim::param(NAME); }
#define GET_DATA_FUNC_E(NAME, FN, ENUMT) ENUMT get##FN() const { return (ENUMT)(int)hdk_shim::param(NAME); }
#define GET_DATA_FUNC_S(NAME, FN) void get##FN(UT_String& str) const { str = UT_String(""); }
#define SET_DATA_FUNC_S(NAME, FN) void set##FN(const UT_String&) {}
#define SET_DATA_FUNC_I(NAME, FN) void set##FN(int) {}
#define SET_DATA_FUNC_F(NAME, FN) void set##FN(fpreal) {}
#define SET_DATA_FUNC_B(NAME, FN) void set##FN(bool) {}
#define DECLARE_STANDARD_GETCASTTOTYPE()
#define DECLARE_DATAFACTORY(CLASS, PARENT, DESC, DOPDESC)
#define IMPLEMENT_DATAFACTORY(CLASS)
