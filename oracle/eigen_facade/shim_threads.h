// shim_threads.h -- TEST INFRASTRUCTURE.  The job pool shared by oracle/hdk_shim/hdk_shim.h (UT_ThreadedAlgorithm, UTparallelFor*,
// THREADED_METHOD*) and oracle/eigen_facade/tbb/tbb.h (tbb::parallel_for): HDK_SHIM_THREADS jobs on std::thread, default 1.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <thread>
#include <vector>
namespace hdk_shim {
inline int& threadsRef() {
    static int n = [] { const char* e = getenv("HDK_SHIM_THREADS"); int v = e ? atoi(e) : 1; if (v <= 0) v = (int)std::thread::hardware_concurrency(); return v < 1 ? 1 : v; }();
    return n;
}
inline int threads() { return threadsRef(); }
inline void setThreads(int n) { if (n <= 0) n = (int)std::thread::hardware_concurrency(); threadsRef() = n < 1 ? 1 : n; }
inline bool& insideJob() { static thread_local bool in = false; return in; }
// f(job) for job in [0, n): jobs 1..n-1 on fresh threads, job 0 on the caller; nested regions run serially
template <class F> inline void runJobs(int n, const F& f) {
    if (n <= 1 || insideJob()) { for (int j = 0; j < n; ++j) f(j); return; }
    std::vector<std::thread> th;
    th.reserve((size_t)n - 1);
    for (int j = 1; j < n; ++j) th.emplace_back([&f, j] { insideJob() = true; f(j); insideJob() = false; });
    insideJob() = true; f(0); insideJob() = false;
    for (auto& t : th) t.join();
}
// contiguous split of [b, e) over the jobs
template <class I, class F> inline void forRange(I b, I e, const F& f) {
    if (!(b < e)) return;
    const long long len = (long long)(e - b);
    const int n = (int)std::min<long long>((long long)threads(), len);
    runJobs(n, [&](int j) { const I lo = b + (I)(len * j / n), hi = b + (I)(len * (j + 1) / n); if (lo < hi) f(lo, hi); });
}
}
