// bsw_host.cpp -- length bucketing and the cost-balanced cut of the multi-GPU partitioner (host side).
//
// Replaces, for the new engine:
//   sortPairsLen / sortPairsId   benchmarks/bsw/bandedSWA.cpp:368-420  (counting sort by len1 in
//                                16384-blocks, then inverse by id)
//   the OpenMP batch loop        benchmarks/bsw/main_banded.cpp:279-291 (dynamic batches -> threads)
// Here the order is a permutation (the caller's array is never moved, pair.id never read),
// sorted by (len2, h0, len1) so that the 32 pairs of a warp walk near-identical DP windows
// (same query length, same initial band, near-equal row count), and the partitioner cuts
// that order into shards of equal estimated DP cost.
#include "bsw_common.h"
#include <cstring>
#include <numeric>

namespace bsw {

ThreadPool& global_pool()
{
    static ThreadPool pool(auto_threads(0));
    return pool;
}

namespace {

inline int bits_for(uint32_t range)      // bits needed to hold values 0..range
{
    int b = 0;
    while (range) { ++b; range >>= 1; }
    return b;
}

} // namespace

void build_sorted_batch(ThreadPool& pool, const SeqPair* pairs, int64_t n, int32_t match, SortedBatch& out)
{
    out.n = n;
    out.cells_nominal = 0;
    out.domain_ok = true;
    const size_t N = (size_t)std::max<int64_t>(n, 0);
    out.idx.reserve(N); out.len2.reserve(N); out.len1.reserve(N); out.h0.reserve(N);
    out.offr.reserve(N); out.offq.reserve(N); out.in_len2.reserve(N); out.in_len1.reserve(N); out.in_h0.reserve(N);
    out.tmpA.reserve(N); out.tmpB.reserve(N);
    if (n <= 0) return;
    const int nt = pool.size();

    // ---- pass A (input order, streaming; the only pass over the caller's SeqPair array):
    //      validate, nominal cells, field ranges, compact copies of the fields used later
    struct Acc { int64_t nominal = 0; int bad = 0; int32_t mn[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, mx[3] = {0, 0, 0}; char pad[64]; };
    std::vector<Acc> acc((size_t)nt);
    const int64_t grain = std::max<int64_t>(8192, (n + 4 * nt - 1) / (4 * nt));
    pool.for_range(n, grain, [&](int64_t b, int64_t e, int t) {
        Acc a = acc[t];
        for (int64_t k = b; k < e; ++k) {
            const SeqPair& sp = pairs[k];
            if (sp.len1 < 1 || sp.len1 > 32767 || sp.len2 < 1 || sp.len2 > 32767 || sp.h0 < 0 ||
                (int64_t)sp.h0 + (int64_t)sp.len2 * match > 32767 || sp.idr < 0 || sp.idq < 0) { a.bad = 1; continue; }
            a.nominal += (int64_t)sp.len1 * sp.len2;
            out.in_len2[k] = (uint16_t)sp.len2; out.in_h0[k] = (uint16_t)sp.h0; out.in_len1[k] = (uint16_t)sp.len1;
            out.offr[k] = (uint64_t)sp.idr; out.offq[k] = (uint64_t)sp.idq;
            const int32_t v[3] = {sp.len2, sp.h0, sp.len1};
            for (int f = 0; f < 3; ++f) { a.mn[f] = std::min(a.mn[f], v[f]); a.mx[f] = std::max(a.mx[f], v[f]); }
        }
        acc[t] = a;
    });
    int32_t mn[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, mx[3] = {0, 0, 0};
    for (const Acc& a : acc) {
        out.cells_nominal += a.nominal;
        if (a.bad) out.domain_ok = false;
        for (int f = 0; f < 3; ++f) { mn[f] = std::min(mn[f], a.mn[f]); mx[f] = std::max(mx[f], a.mx[f]); }
    }
    if (!out.domain_ok) return;

    // ---- compressed key (len2 | h0 | len1, only as many bits as vary; if that exceeds 32 bits
    //      the low bits are dropped -- the order is a performance heuristic, any permutation is valid)
    const int b_l1 = bits_for((uint32_t)(mx[2] - mn[2])), b_h0 = bits_for((uint32_t)(mx[1] - mn[1]));
    const int b_l2 = bits_for((uint32_t)(mx[0] - mn[0]));
    const int full_bits = b_l1 + b_h0 + b_l2;
    const int drop = std::max(0, full_bits - 32);
    const int total_bits = full_bits - drop;
    const int npass = total_bits == 0 ? 0 : (total_bits + 10) / 11;
    const int digit = npass ? (total_bits + npass - 1) / npass : 0;
    const int RAD = 1 << digit;
    uint64_t* const A = out.tmpA.data(); uint64_t* const B = out.tmpB.data();
    pool.for_range(n, grain, [&](int64_t b, int64_t e, int) {
        for (int64_t k = b; k < e; ++k) {
            const uint64_t key = ((uint64_t)(out.in_len2[k] - mn[0]) << (b_h0 + b_l1)) |
                                 ((uint64_t)(out.in_h0[k] - mn[1]) << b_l1) | (uint64_t)(out.in_len1[k] - mn[2]);
            A[k] = ((key >> drop) << 32) | (uint64_t)k;
        }
    });

    // ---- stable LSD radix sort, per-chunk histograms (chunks are contiguous => stable)
    const int64_t sgrain = (n + nt - 1) / nt;
    const int nch = (int)((n + sgrain - 1) / sgrain);
    std::vector<int64_t> hist;
    if (npass) hist.resize((size_t)nch * RAD);
    uint64_t* in = A; uint64_t* outp = B;
    for (int pass = 0; pass < npass; ++pass) {
        const int shift = 32 + pass * digit;
        std::fill(hist.begin(), hist.end(), 0);
        pool.for_range(n, sgrain, [&](int64_t b, int64_t e, int) {
            int64_t* h = hist.data() + (size_t)(b / sgrain) * RAD;
            for (int64_t k = b; k < e; ++k) ++h[(in[k] >> shift) & (RAD - 1)];
        });
        int64_t run = 0;
        for (int d = 0; d < RAD; ++d)
            for (int c = 0; c < nch; ++c) {
                int64_t& slot = hist[(size_t)c * RAD + d];
                const int64_t cnt = slot; slot = run; run += cnt;
            }
        pool.for_range(n, sgrain, [&](int64_t b, int64_t e, int) {
            int64_t* h = hist.data() + (size_t)(b / sgrain) * RAD;
            for (int64_t k = b; k < e; ++k) outp[h[(in[k] >> shift) & (RAD - 1)]++] = in[k];
        });
        std::swap(in, outp);
    }
    // ---- processing-order SoA the later passes stream over
    pool.for_range(n, grain, [&](int64_t b, int64_t e, int) {
        for (int64_t s = b; s < e; ++s) {
            const uint32_t i = (uint32_t)in[s];
            out.idx[s] = i;
            out.len2[s] = out.in_len2[i]; out.len1[s] = out.in_len1[i]; out.h0[s] = out.in_h0[i];
        }
    });
}

// Estimated DP cost of a pair: rows x min(columns, band).
static inline int64_t pair_cost(int32_t len1, int32_t len2, int32_t w)
{
    const int64_t band = 2ll * w + 1;
    return (int64_t)len1 * std::min<int64_t>(len2, band) + 64;   // +64: per-pair fixed overhead
}

} // namespace bsw

using namespace bsw;

extern "C" {

int bsw_bucket_order(const SeqPair* pairs, int64_t n, int64_t* order)
{
    if (n < 0 || n > 0x7fffffff || (n > 0 && (!pairs || !order))) return BSW_ERR_PARAM;
    SortedBatch sb;
    build_sorted_batch(global_pool(), pairs, n, 0, sb);
    if (!sb.domain_ok) return BSW_ERR_DOMAIN;
    for (int64_t s = 0; s < n; ++s) order[s] = sb.idx[s];
    return BSW_OK;
}

// Contiguous cost-balanced cut of the input order (exact: every record is read; the engine's own cut inside
// bsw_extend reads a sample).  shard_begin has n_shards + 1 entries.
int bsw_split_by_cost(const SeqPair* pairs, int64_t n, int32_t w, int32_t n_shards, int64_t* shard_begin)
{
    if (n < 0 || n_shards < 1 || !shard_begin || w < 0 || (n > 0 && !pairs)) return BSW_ERR_PARAM;
    std::vector<int64_t> cum((size_t)n + 1, 0);
    for (int64_t k = 0; k < n; ++k) {
        if (pairs[k].len1 < 1 || pairs[k].len2 < 1 || pairs[k].len1 > 32767 || pairs[k].len2 > 32767) return BSW_ERR_DOMAIN;
        cum[(size_t)k + 1] = cum[(size_t)k] + pair_cost(pairs[k].len1, pairs[k].len2, w);
    }
    shard_begin[0] = 0; shard_begin[n_shards] = n;
    for (int g = 1; g < n_shards; ++g) {
        const int64_t target = cum[(size_t)n] / n_shards * g;
        const int64_t k = std::lower_bound(cum.begin(), cum.end(), target) - cum.begin();
        shard_begin[g] = std::max(shard_begin[g - 1], std::min(n, k));
    }
    return BSW_OK;
}

} // extern "C"
