// bsw_global_plan.h -- host-side plan of one chunk of bsw_global on the second kernel (bsw_global2.cuh): which
// alignments form the chunk, where each one's bytes / direction words / operation list live, the order the threads
// run in and the launches (one per shared-memory class).  Plain C++ -- no CUDA call -- so the very same code runs
// under the engine (bsw_global.inl, on the thread pool) and under the CPU emulation of the kernel
// (tests/emu/g2_emu.cu), which checks plan + kernel + compaction against the reference's goldens without a GPU.
//
// Passes over the chunk's alignments are handed to a `par(n, grain, fn(b, e, tid))` callable; everything a pass
// writes is indexed by alignment or by range, so the result does not depend on how the ranges were scheduled.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../include/bsw.h"
#include "bsw_global2.cuh"

namespace bsw {
namespace g2 {

struct Launch {
    int32_t first, count;            // desc[first .. first + count) of the chunk's sorted descriptors
    int32_t slots, qwords;           // the class: row slots and query words per thread
};

constexpr int64_t PLAN_RANGE = 2048;             // alignments per range of a pass

struct Caps {
    long long z_bytes = 3ll << 29;               // direction words of a chunk
    long long cig_words = 1ll << 27;             // uncompacted operation lists of a chunk
    int64_t m = 1 << 18;                         // alignments per chunk (the sort key keeps 18 bits of index)
    long long seq_bytes = 1ll << 31;             // gathered queries / targets of a chunk (descriptors hold 32-bit offsets)
};

struct ChunkPlan {
    int64_t first = 0, m = 0;
    long long q_bytes = 0, r_bytes = 0, z_bytes = 0, cig_words = 0, cells = 0, cells_nominal = 0;
    std::vector<Launch> launches;
    // per range of PLAN_RANGE alignments: where its first alignment's bytes / words start
    std::vector<long long> rq, rr, rz, rc;
    std::vector<uint64_t> key, tmp;              // work order
    std::vector<GlobalDesc> din;                 // descriptors in input order (permuted into work order by plan_fill)
    std::vector<uint32_t> hist;
};

// classes: a launch's shared memory is sized for the class, so its steps bound the waste (<= 25 % above 64 slots)
inline int slots_class(int s)
{
    const int g = s <= 96 ? 16 : s <= 192 ? 32 : s <= 384 ? 64 : 128;
    return (s + g - 1) / g * g;
}
inline int qwords_class(int qw)
{
    if (qw <= 64) return (qw + 15) / 16 * 16;
    int c = 128;
    while (c < qw) c <<= 1;
    return c;
}
// a band at least as wide as the longer sequence is no band at all: every row spans the whole query, and any such
// w gives the same cells, the same first row and column (ksw.c:521-531) -- the rows need no more slots than that
inline int eff_w(int qlen, int tlen, int w) { const int m = qlen > tlen ? qlen : tlen; return w < m ? w : m; }
inline long long q_span(int qlen) { return ((long long)qlen + 7) & ~7ll; }      // bytes a query occupies in the gathered buffer
inline long long r_span(int tlen) { return ((long long)tlen + 3) & ~3ll; }
inline long long z_span(int qlen, int tlen, int w) { return (((long long)z_pitch(qlen, w) * 4 * tlen) + 7) & ~7ll; }
// DP cells inside the band: sum over rows i < tlen of min(i + w + 1, qlen) - max(i - w, 0), in closed form
// (|tlen - qlen| <= w keeps every row's window non-empty)
inline long long band_cells(long long Q, long long T, long long W)
{
    const long long a = std::min(T, std::max(0ll, Q - W - 1));           // rows whose window ends before qlen
    const long long kb = std::max(0ll, T - W - 1);                        // rows whose window starts after 0
    return a * (a - 1) / 2 + a * (W + 1) + (T - a) * Q - kb * (kb + 1) / 2;
}

// Pass 1: the chunk [first, first + m) -- as many whole ranges as the caps allow (a range that alone exceeds them is
// cut inside) -- and the byte / word totals.  Fills everything of `pl` but the launches.
template <class Par>
void plan_sizes(const SeqPair* pairs, const int32_t* w, int64_t first, int64_t n, const Caps& caps, Par&& par, ChunkPlan& pl)
{
    const int64_t avail = std::min<int64_t>(n - first, caps.m);
    const int64_t nr = (avail + PLAN_RANGE - 1) / PLAN_RANGE;
    pl.rq.assign((size_t)nr + 1, 0); pl.rr.assign((size_t)nr + 1, 0); pl.rz.assign((size_t)nr + 1, 0); pl.rc.assign((size_t)nr + 1, 0);
    std::vector<long long> cells((size_t)nr, 0), nominal((size_t)nr, 0);
    par(nr, 1, [&](int64_t rb, int64_t re, int) {
        for (int64_t rg = rb; rg < re; ++rg) {
            long long q = 0, r = 0, z = 0, c = 0, ce = 0, cn = 0;
            const int64_t b = first + rg * PLAN_RANGE, e = std::min(first + avail, b + PLAN_RANGE);
            for (int64_t i = b; i < e; ++i) {
                const SeqPair& sp = pairs[i];
                const int wv = eff_w(sp.len2, sp.len1, w[i]);
                q += q_span(sp.len2); r += r_span(sp.len1); z += z_span(sp.len2, sp.len1, wv); c += (long long)sp.len1 + sp.len2;
                ce += band_cells(sp.len2, sp.len1, wv); cn += (long long)sp.len1 * sp.len2;
            }
            pl.rq[(size_t)rg + 1] = q; pl.rr[(size_t)rg + 1] = r; pl.rz[(size_t)rg + 1] = z; pl.rc[(size_t)rg + 1] = c;
            cells[(size_t)rg] = ce; nominal[(size_t)rg] = cn;
        }
    });
    // ranges into the chunk while they fit
    int64_t take = 0;
    pl.cells = pl.cells_nominal = 0;
    for (; take < nr; ++take) {
        const long long z = pl.rz[(size_t)take] + pl.rz[(size_t)take + 1], c = pl.rc[(size_t)take] + pl.rc[(size_t)take + 1];
        if (z > caps.z_bytes || c > caps.cig_words || pl.rq[(size_t)take] + pl.rq[(size_t)take + 1] > caps.seq_bytes ||
            pl.rr[(size_t)take] + pl.rr[(size_t)take + 1] > caps.seq_bytes) break;
        pl.rq[(size_t)take + 1] += pl.rq[(size_t)take]; pl.rr[(size_t)take + 1] += pl.rr[(size_t)take];
        pl.rz[(size_t)take + 1] = z; pl.rc[(size_t)take + 1] = c;
        pl.cells += cells[(size_t)take]; pl.cells_nominal += nominal[(size_t)take];
    }
    pl.first = first;
    if (take > 0) {
        pl.m = std::min<int64_t>(avail, take * PLAN_RANGE);
    } else {
        // the first range alone is too much: alignment by alignment (at least one -- a single alignment always runs)
        long long q = 0, r = 0, z = 0, c = 0;
        int64_t m = 0;
        pl.cells = pl.cells_nominal = 0;
        for (; m < std::min<int64_t>(avail, PLAN_RANGE); ++m) {
            const SeqPair& sp = pairs[first + m];
            const int wv = eff_w(sp.len2, sp.len1, w[first + m]);
            const long long zi = z_span(sp.len2, sp.len1, wv), ci = (long long)sp.len1 + sp.len2;
            if (m > 0 && (z + zi > caps.z_bytes || c + ci > caps.cig_words || q + q_span(sp.len2) > caps.seq_bytes ||
                          r + r_span(sp.len1) > caps.seq_bytes)) break;
            q += q_span(sp.len2); r += r_span(sp.len1); z += zi; c += ci;
            pl.cells += band_cells(sp.len2, sp.len1, wv); pl.cells_nominal += (long long)sp.len1 * sp.len2;
        }
        pl.m = m; take = 1;
        pl.rq[1] = q; pl.rr[1] = r; pl.rz[1] = z; pl.rc[1] = c;
    }
    pl.q_bytes = pl.rq[(size_t)take]; pl.r_bytes = pl.rr[(size_t)take]; pl.z_bytes = pl.rz[(size_t)take]; pl.cig_words = pl.rc[(size_t)take];
}

// stable LSD radix sort of keys[0 .. n) on bits [lo, hi): 12-bit digits, `slices` contiguous slices counted and
// scattered side by side
template <class Par>
void radix_sort(std::vector<uint64_t>& keys, std::vector<uint64_t>& tmp, std::vector<uint32_t>& hist, int lo, int hi, int slices, Par&& par)
{
    const int64_t n = (int64_t)keys.size();
    if (n < 2) return;
    constexpr int BITS = 12, BINS = 1 << BITS;
    slices = (int)std::max<int64_t>(1, std::min<int64_t>(slices, n / 4096));
    tmp.resize((size_t)n);
    hist.resize((size_t)slices * BINS);
    const int64_t per = (n + slices - 1) / slices;
    for (int sh = lo; sh < hi; sh += BITS) {
        const uint64_t mask = (uint64_t)(std::min(BITS, hi - sh) == BITS ? BINS - 1 : (1 << (hi - sh)) - 1);
        par(slices, 1, [&](int64_t sb, int64_t se, int) {
            for (int64_t s = sb; s < se; ++s) {
                uint32_t* h = hist.data() + (size_t)s * BINS;
                std::fill(h, h + BINS, 0u);
                const int64_t b = s * per, e = std::min(n, b + per);
                for (int64_t k = b; k < e; ++k) ++h[(keys[(size_t)k] >> sh) & mask];
            }
        });
        uint32_t run = 0;
        for (int d = 0; d < BINS; ++d)
            for (int s = 0; s < slices; ++s) { uint32_t& c = hist[(size_t)s * BINS + d]; const uint32_t t = c; c = run; run += t; }
        par(slices, 1, [&](int64_t sb, int64_t se, int) {
            for (int64_t s = sb; s < se; ++s) {
                uint32_t* h = hist.data() + (size_t)s * BINS;
                const int64_t b = s * per, e = std::min(n, b + per);
                for (int64_t k = b; k < e; ++k) { const uint64_t v = keys[(size_t)k]; tmp[h[(v >> sh) & mask]++] = v; }
            }
        });
        keys.swap(tmp);
    }
}

// Pass 2: descriptors in work order into desc[0 .. m), the two byte strings of every alignment into hq / hr
// (q_bytes / r_bytes of them, every alignment 4-aligned, queries padded to 8), the launches.
// Work order: class (slots, query words) descending, inside a class band width descending, then target length --
// a warp's lanes share the row loop (target length) and the block loop (band), and the longest run first.
template <class Par>
void plan_fill(const SeqPair* pairs, const int32_t* w, const uint8_t* seq_ref, const uint8_t* seq_qer, int slices, Par&& par,
               ChunkPlan& pl, GlobalDesc* desc, uint8_t* hq, uint8_t* hr)
{
    const int64_t m = pl.m, first = pl.first;
    const int64_t nr = (m + PLAN_RANGE - 1) / PLAN_RANGE;
    pl.key.resize((size_t)m);
    pl.din.resize((size_t)m);
    GlobalDesc* const din = pl.din.data();
    par(nr, 1, [&](int64_t rb, int64_t re, int) {
        for (int64_t rg = rb; rg < re; ++rg) {
            long long q = pl.rq[(size_t)rg], r = pl.rr[(size_t)rg], z = pl.rz[(size_t)rg], c = pl.rc[(size_t)rg];
            const int64_t b = rg * PLAN_RANGE, e = std::min(m, b + PLAN_RANGE);
            for (int64_t k = b; k < e; ++k) {
                const SeqPair& sp = pairs[first + k];
                const int wv = eff_w(sp.len2, sp.len1, w[first + k]);
                GlobalDesc& d = din[k];
                d.qoff = (uint32_t)q; d.roff = (uint32_t)r; d.qlen = sp.len2; d.tlen = sp.len1; d.w = wv; d.idx = (int32_t)k;
                d.zoff = z; d.coff = c;
                memcpy(hq + q, seq_qer + sp.idq, (size_t)sp.len2);
                memcpy(hr + r, seq_ref + sp.idr, (size_t)sp.len1);
                q += q_span(sp.len2); r += r_span(sp.len1); z += z_span(sp.len2, sp.len1, wv); c += (long long)sp.len1 + sp.len2;
                const uint64_t sc = (uint64_t)slots_class(row_slots(sp.len2, wv)) / 16;              // <= 2^12 / 16 ... 7 bits inside the kernel's reach
                const uint64_t qc = (uint64_t)qwords_class(query_words(sp.len2)) / 16;               // <= 4096 / 16 + 1: 9 bits
                const uint64_t band = (uint64_t)(sp.len2 < 2 * wv + 1 ? sp.len2 : 2 * wv + 1);       // 15 bits
                const uint64_t key = sc << 39 | qc << 30 | band << 15 | (uint64_t)sp.len1;            // 46 bits
                pl.key[(size_t)k] = (~key & ((1ull << 46) - 1)) << 18 | (uint64_t)k;                 // ascending sort = descending key
            }
        }
    });
    radix_sort(pl.key, pl.tmp, pl.hist, 18, 64, slices, par);
    // descriptors into work order; a launch starts wherever the class changes
    std::vector<std::vector<int32_t>> cuts((size_t)nr);
    par(nr, 1, [&](int64_t rb, int64_t re, int) {
        for (int64_t rg = rb; rg < re; ++rg) {
            const int64_t b = rg * PLAN_RANGE, e = std::min(m, b + PLAN_RANGE);
            for (int64_t k = b; k < e; ++k) {
                desc[k] = din[pl.key[(size_t)k] & 0x3ffff];
                if (k == 0 || (pl.key[(size_t)k] >> 48) != (pl.key[(size_t)k - 1] >> 48)) cuts[(size_t)rg].push_back((int32_t)k);
            }
        }
    });
    pl.launches.clear();
    for (const auto& v : cuts)
        for (int32_t k : v) {
            if (!pl.launches.empty()) pl.launches.back().count = k - pl.launches.back().first;
            const uint64_t key = ~(pl.key[(size_t)k] >> 18) & ((1ull << 46) - 1);
            pl.launches.push_back(Launch{k, 0, (int32_t)(key >> 39) * 16, (int32_t)(key >> 30 & 0x1ff) * 16});
        }
    if (!pl.launches.empty()) pl.launches.back().count = (int32_t)m - pl.launches.back().first;
}

} // namespace g2
} // namespace bsw
