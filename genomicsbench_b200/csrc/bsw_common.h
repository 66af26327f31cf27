// bsw_common.h -- shared host-side helpers of the bsw_b200 library (not part of the ABI).
#pragma once
#include <cstdint>
#include <cstddef>
#include <thread>
#include <vector>
#include <algorithm>
#include <chrono>
#include "../../include/bsw.h"

namespace bsw {

inline double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

inline int auto_threads(int requested)
{
    if (requested > 0) return requested;
    unsigned hc = std::thread::hardware_concurrency();
    if (hc == 0) hc = 4;
    return (int)std::min<unsigned>(hc, 32u);
}

// Static block-cyclic parallel loop over [0, n) in chunks of `grain`; fn(begin, end, tid).
template <class F>
void parallel_chunks(int64_t n, int64_t grain, int nthreads, F&& fn)
{
    if (n <= 0) return;
    if (grain < 1) grain = 1;
    const int64_t nchunks = (n + grain - 1) / grain;
    nthreads = (int)std::min<int64_t>(std::max(nthreads, 1), nchunks);
    if (nthreads == 1) { fn((int64_t)0, n, 0); return; }
    std::vector<std::thread> th;
    th.reserve(nthreads);
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([=, &fn]() {
            for (int64_t c = t; c < nchunks; c += nthreads) {
                const int64_t b = c * grain, e = std::min(n, b + grain);
                fn(b, e, t);
            }
        });
    for (auto& x : th) x.join();
}

// splitmix64: the generator's PRNG.
struct SplitMix64 {
    uint64_t s;
    explicit SplitMix64(uint64_t seed) : s(seed) {}
    inline uint64_t next()
    {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    // uniform integer in [lo, hi]
    inline int32_t range(int32_t lo, int32_t hi)
    {
        if (hi <= lo) return lo;
        return lo + (int32_t)(next() % (uint64_t)(hi - lo + 1));
    }
    inline double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

} // namespace bsw
