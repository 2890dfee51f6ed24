// eigen_facade / hdk_shim: stands in for <tbb/tbb.h> (absent offline).  units.h:7 and the exec/ sources include it; the uses compiled
// here are tbb::parallel_for over a tbb::blocked_range, run serially in order (one job, like the rest of the shim).
// The real header also drags in the standard headers pcg.h relies on without including them itself.
#pragma once
#include <chrono>
#include <cmath>
#include <cstddef>
#include <iostream>
namespace tbb {
template <class T>
class blocked_range {
public:
    blocked_range(T b, T e, std::size_t = 1) : b_(b), e_(e) {}
    T begin() const { return b_; }
    T end() const { return e_; }
    bool empty() const { return !(b_ < e_); }
private:
    T b_, e_;
};
template <class R, class F> inline void parallel_for(const R& r, const F& f) { if (!r.empty()) f(r); }
}
#include <vector>
namespace tbb {
// one "thread": local() is the single instance, combine_each visits it once
template <class T>
class enumerable_thread_specific {
public:
    T& local() { return v; }
    template <class F> void combine_each(F f) const { f(v); }
    void clear() { v = T(); }
private:
    T v;
};
}
