// eigen_facade / hdk_shim: stands in for <tbb/tbb.h> (absent offline).  units.h:7 and the exec/ sources include it; the uses compiled
// here are tbb::parallel_for over a tbb::blocked_range, cut into one contiguous piece per job of the shim (one job by default).
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
}
// The jobs of oracle/hdk_shim (HDK_SHIM_THREADS, default 1 = serial and in order): a blocked_range is cut into one contiguous piece per job.
#include <mutex>
#include <list>
#include <thread>
#include <vector>
#include <utility>
#include "../shim_threads.h"
namespace tbb {
template <class R, class F> inline void parallel_for(const R& r, const F& f) {
    if (r.empty()) return;
    const long long len = (long long)(r.end() - r.begin());
    long long n = hdk_shim::threads(); if (n > len) n = len;
    if (n <= 1 || hdk_shim::insideJob()) { f(r); return; }
    std::vector<std::thread> th;
    auto piece = [&](long long j) { const auto lo = r.begin() + (decltype(r.begin()))(len * j / n), hi = r.begin() + (decltype(r.begin()))(len * (j + 1) / n); if (lo < hi) f(R(lo, hi)); };
    for (long long j = 1; j < n; ++j) th.emplace_back([&piece, j] { hdk_shim::insideJob() = true; piece(j); hdk_shim::insideJob() = false; });
    hdk_shim::insideJob() = true; piece(0); hdk_shim::insideJob() = false;
    for (auto& t : th) t.join();
}
// one instance per calling thread, created on first use; combine_each visits them in creation order
template <class T>
class enumerable_thread_specific {
public:
    T& local() {
        const std::thread::id me = std::this_thread::get_id();
        std::lock_guard<std::mutex> g(m);
        for (auto& e : items) if (e.first == me) return e.second;
        items.emplace_back(me, T());
        return items.back().second;
    }
    template <class F> void combine_each(F f) const { for (const auto& e : items) f(e.second); }
    void clear() { items.clear(); }
private:
    std::mutex m;
    std::list<std::pair<std::thread::id, T> > items;
};
}
