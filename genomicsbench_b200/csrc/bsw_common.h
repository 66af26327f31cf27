// bsw_common.h -- shared host-side helpers of the bsw_b200 library (not part of the ABI).
#pragma once
#include <cstdint>
#include <cstddef>
#include <thread>
#include <vector>
#include <algorithm>
#include <chrono>
#include <atomic>
#include <mutex>
#include <condition_variable>
#include <functional>
#include <cstdlib>
#include <cstring>
#include <new>
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

// Persistent worker pool: run(nchunks, fn) executes fn(chunk, tid) for every chunk in
// [0, nchunks), chunks handed out dynamically; the calling thread works too (tid 0).
class ThreadPool {
public:
    explicit ThreadPool(int nthreads) : n_(std::max(nthreads, 1))
    {
        for (int t = 1; t < n_; ++t) workers_.emplace_back([this, t] { loop(t); });
    }
    ~ThreadPool()
    {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
            ++epoch_;
        }
        cv_.notify_all();
        for (auto& w : workers_) w.join();
    }
    int size() const { return n_; }

    template <class F>
    void run(int64_t nchunks, F&& fn)
    {
        if (nchunks <= 0) return;
        if (n_ == 1 || nchunks == 1) {
            for (int64_t c = 0; c < nchunks; ++c) fn(c, 0);
            return;
        }
        std::lock_guard<std::mutex> serial(run_m_);       // one parallel region at a time
        std::function<void(int64_t, int)> f = std::ref(fn);
        {
            std::lock_guard<std::mutex> g(m_);
            job_ = &f;
            total_ = nchunks;
            next_.store(0, std::memory_order_relaxed);
            pending_ = n_ - 1;
            ++epoch_;
        }
        cv_.notify_all();
        work(f, 0);
        std::unique_lock<std::mutex> lk(m_);
        done_cv_.wait(lk, [this] { return pending_ == 0; });
        job_ = nullptr;
    }

    // Convenience: contiguous ranges [b, e) of `grain` items; fn(b, e, tid).
    template <class F>
    void for_range(int64_t n, int64_t grain, F&& fn)
    {
        if (n <= 0) return;
        grain = std::max<int64_t>(grain, 1);
        const int64_t nchunks = (n + grain - 1) / grain;
        run(nchunks, [&](int64_t c, int tid) { fn(c * grain, std::min(n, (c + 1) * grain), tid); });
    }

private:
    void work(std::function<void(int64_t, int)>& f, int tid)
    {
        for (;;) {
            const int64_t c = next_.fetch_add(1, std::memory_order_relaxed);
            if (c >= total_) break;
            f(c, tid);
        }
    }
    void loop(int tid)
    {
        uint64_t seen = 0;
        for (;;) {
            std::function<void(int64_t, int)>* f;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return epoch_ != seen; });
                seen = epoch_;
                if (stop_) return;
                f = job_;
            }
            if (f) work(*f, tid);
            {
                std::lock_guard<std::mutex> g(m_);
                if (--pending_ == 0) done_cv_.notify_one();
            }
        }
    }
    int n_;
    std::vector<std::thread> workers_;
    std::mutex m_, run_m_;
    std::condition_variable cv_, done_cv_;
    std::function<void(int64_t, int)>* job_ = nullptr;
    std::atomic<int64_t> next_{0};
    int64_t total_ = 0;
    int pending_ = 0;
    uint64_t epoch_ = 0;
    bool stop_ = false;
};

// Process-wide pool for the host utilities that have no engine (bsw_bucket_order, generator).
ThreadPool& global_pool();

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
    inline int32_t range(int32_t lo, int32_t hi)     // uniform integer in [lo, hi]
    {
        if (hi <= lo) return lo;
        return lo + (int32_t)(next() % (uint64_t)(hi - lo + 1));
    }
    inline double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

// ---- 2-bit packing (host): 16 bases per 32-bit word, base k of a sequence at bits [2k, 2k+2) of word k / 16 -------
// 8 base codes (one per byte, little endian in v) -> 16 bits
inline uint32_t pack8(uint64_t v)
{
    v &= 0x0303030303030303ull;
    v = (v | (v >> 6)) & 0x000F000F000F000Full;
    v = (v | (v >> 12)) & 0x000000FF000000FFull;
    return (uint32_t)((v | (v >> 24)) & 0xFFFFu);
}

// packs len bases into ceil(len / 16) words; returns true when a code above 3 (N) was seen (the words are then unusable).
// Never reads outside [src, src + len): the last, partial word is taken from the sequence's LAST 8 bytes shifted into
// place (a byte loop only for sequences shorter than 8).
inline bool pack_seq(const uint8_t* src, int len, uint32_t* dst)
{
    int k = 0, wi = 0;
    uint64_t high = 0;
    for (; k + 16 <= len; k += 16, ++wi) {
        uint64_t a, b;
        memcpy(&a, src + k, 8); memcpy(&b, src + k + 8, 8);
        high |= a | b;
        dst[wi] = pack8(a) | (pack8(b) << 16);
    }
    int r = len - k;
    if (r > 0) {
        uint32_t wv = 0;
        int sh = 0;
        if (r >= 8) {
            uint64_t a;
            memcpy(&a, src + k, 8);
            high |= a;
            wv = pack8(a);
            k += 8; r -= 8; sh = 16;
        }
        if (r > 0) {
            if (len >= 8) {
                uint64_t a;
                memcpy(&a, src + len - 8, 8);
                a >>= 8 * (8 - r);                      // the last r bytes, in the low bytes
                high |= a;
                wv |= pack8(a) << sh;
            } else {
                for (int j = 0; j < r; ++j) { high |= src[k + j]; wv |= (uint32_t)(src[k + j] & 3u) << (2 * j + sh); }
            }
        }
        dst[wi] = wv;
    }
    return (high & 0xFCFCFCFCFCFCFCFCull) != 0;
}

// ---- bucketing (bsw_host.cpp) ----------------------------------------------------------
// A batch in processing order: position s holds caller index idx[s]; len2/len1/h0 of that
// pair are kept alongside so later passes never touch the caller's array out of order.
// Grow-only uninitialised array: reused across calls so that steady-state batches pay neither
// the zero-fill of std::vector nor fresh page faults.
template <class T>
struct RawBuf {
    T* p = nullptr; size_t cap = 0;
    RawBuf() = default;
    RawBuf(const RawBuf&) = delete;
    RawBuf& operator=(const RawBuf&) = delete;
    ~RawBuf() { free(p); }
    void reserve(size_t n)
    {
        if (n <= cap) return;
        free(p);
        cap = n + n / 8 + 64;
        p = static_cast<T*>(malloc(cap * sizeof(T)));
        if (!p) { cap = 0; throw std::bad_alloc(); }
    }
    void swap(RawBuf& o) { std::swap(p, o.p); std::swap(cap, o.cap); }
    T& operator[](size_t i) { return p[i]; }
    const T& operator[](size_t i) const { return p[i]; }
    T* data() { return p; }
    const T* data() const { return p; }
};

struct SortedBatch {
    int64_t n = 0;
    // caller (input) order: everything later passes need from the 72-byte SeqPair records, so the
    // caller's array is read exactly once, sequentially
    RawBuf<uint64_t> offr, offq;    // idr / idq
    RawBuf<uint16_t> in_len2, in_len1, in_h0;
    // processing order
    RawBuf<uint32_t> idx;           // processing order -> caller index
    RawBuf<uint16_t> len2, len1, h0;
    RawBuf<uint64_t> tmpA, tmpB;    // radix-sort scratch: (key32 << 32) | idx
    int64_t cells_nominal = 0;
    bool domain_ok = true;
};

// Validates the domain, accumulates sum len1*len2 and sorts by (len2, h0, len1).
void build_sorted_batch(ThreadPool& pool, const SeqPair* pairs, int64_t n, int32_t match, SortedBatch& out);

} // namespace bsw
