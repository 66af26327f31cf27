// bsw_host.cpp -- length bucketing and the multi-GPU partitioner (host side).
//
// Replaces, for the new engine:
//   sortPairsLen / sortPairsId   benchmarks/bsw/bandedSWA.cpp:368-420  (counting sort by len1 in
//                                16384-blocks, then inverse by id)
//   the OpenMP batch loop        benchmarks/bsw/main_banded.cpp:279-291 (dynamic batches -> threads)
// Here the order is a permutation (the caller's array is never moved, pair.id never read),
// sorted by (len2, len1, h0) so that the 32 pairs of a warp walk near-identical DP windows,
// and the partitioner cuts that order into shards of equal estimated DP cost.
#include "bsw_common.h"
#include <cstring>
#include <numeric>

using namespace bsw;

namespace bsw {

// 45-bit sort key: len2 | len1 | h0 (clamped to 15 bits each)
static inline uint64_t sort_key(const SeqPair& p)
{
    const uint64_t l2 = (uint64_t)std::min(std::max(p.len2, 0), 32767);
    const uint64_t l1 = (uint64_t)std::min(std::max(p.len1, 0), 32767);
    const uint64_t h0 = (uint64_t)std::min(std::max(p.h0, 0), 32767);
    return (l2 << 30) | (l1 << 15) | h0;
}

// Parallel LSD radix sort of (key, index) by 3 x 15-bit digits.  order[] gets the permutation.
void bucket_order(const SeqPair* pairs, int64_t n, int64_t* order, int nthreads)
{
    if (n <= 0) return;
    nthreads = auto_threads(nthreads);
    const int64_t grain = std::max<int64_t>(4096, (n + nthreads - 1) / nthreads);
    const int nchunks = (int)((n + grain - 1) / grain);
    std::vector<uint64_t> keyA((size_t)n), keyB((size_t)n);
    std::vector<int64_t> idxB((size_t)n);
    int64_t* idxA = order;
    parallel_chunks(n, grain, nthreads, [&](int64_t b, int64_t e, int) {
        for (int64_t k = b; k < e; ++k) { keyA[k] = sort_key(pairs[k]); idxA[k] = k; }
    });
    // skip digits that are constant across the batch (common: fixed read length)
    uint64_t all_or = 0, all_and = ~0ull;
    for (int64_t k = 0; k < n; ++k) { all_or |= keyA[k]; all_and &= keyA[k]; }
    const uint64_t varying = all_or ^ all_and;

    const int RAD = 1 << 15;
    std::vector<int64_t> hist((size_t)nchunks * RAD);
    uint64_t* kin = keyA.data(); uint64_t* kout = keyB.data();
    int64_t* iin = idxA; int64_t* iout = idxB.data();
    for (int pass = 0; pass < 3; ++pass) {
        const int shift = pass * 15;
        if (((varying >> shift) & (RAD - 1)) == 0) continue;
        std::fill(hist.begin(), hist.end(), 0);
        parallel_chunks(n, grain, nthreads, [&](int64_t b, int64_t e, int) {
            int64_t* h = hist.data() + (size_t)(b / grain) * RAD;
            for (int64_t k = b; k < e; ++k) ++h[(kin[k] >> shift) & (RAD - 1)];
        });
        int64_t run = 0;
        for (int d = 0; d < RAD; ++d)
            for (int c = 0; c < nchunks; ++c) {
                int64_t& slot = hist[(size_t)c * RAD + d];
                const int64_t cnt = slot; slot = run; run += cnt;
            }
        parallel_chunks(n, grain, nthreads, [&](int64_t b, int64_t e, int) {
            int64_t* h = hist.data() + (size_t)(b / grain) * RAD;
            for (int64_t k = b; k < e; ++k) {
                const int64_t pos = h[(kin[k] >> shift) & (RAD - 1)]++;
                kout[pos] = kin[k]; iout[pos] = iin[k];
            }
        });
        std::swap(kin, kout); std::swap(iin, iout);
    }
    if (iin != order) memcpy(order, iin, sizeof(int64_t) * (size_t)n);
}

// Estimated DP cost of a pair: rows x min(columns, band).
static inline int64_t pair_cost(const SeqPair& p, int32_t w)
{
    const int64_t band = 2ll * w + 1;
    return (int64_t)p.len1 * std::min<int64_t>(p.len2, band) + 64;   // +64: per-pair fixed overhead
}

} // namespace bsw

extern "C" {

int bsw_bucket_order(const SeqPair* pairs, int64_t n, int64_t* order)
{
    if ((!pairs || !order) && n > 0) return BSW_ERR_PARAM;
    if (n < 0) return BSW_ERR_PARAM;
    bucket_order(pairs, n, order, 0);
    return BSW_OK;
}

// Cuts the bucketed order into n_shards interleaved-by-block shards of near-equal cost.
// Shard g owns order[shard_begin[g] .. shard_begin[g+1]).  To keep every shard's length mix
// (and therefore kernel occupancy classes) similar, the sorted order is dealt out in
// blocks of 1024 pairs round-robin by running cost, and order[] is rewritten shard-major.
int bsw_partition(const SeqPair* pairs, int64_t n, int32_t w, int32_t n_shards,
                  int64_t* order, int64_t* shard_begin)
{
    if (n < 0 || n_shards < 1 || !shard_begin || (n > 0 && (!pairs || !order))) return BSW_ERR_PARAM;
    bucket_order(pairs, n, order, 0);
    if (n_shards == 1) { shard_begin[0] = 0; shard_begin[1] = n; return BSW_OK; }
    const int64_t BLK = 1024;
    const int64_t nblk = (n + BLK - 1) / BLK;
    std::vector<int64_t> cost((size_t)n_shards, 0);
    std::vector<int32_t> owner((size_t)nblk);
    std::vector<int64_t> count((size_t)n_shards, 0);
    // longest blocks first (the order is ascending in length): greedy onto the lightest shard
    for (int64_t b = nblk - 1; b >= 0; --b) {
        int64_t c = 0;
        const int64_t lo = b * BLK, hi = std::min(n, lo + BLK);
        for (int64_t k = lo; k < hi; ++k) c += pair_cost(pairs[order[k]], w);
        int best = 0;
        for (int g = 1; g < n_shards; ++g) if (cost[g] < cost[best]) best = g;
        owner[b] = best; cost[best] += c; count[best] += hi - lo;
    }
    shard_begin[0] = 0;
    for (int g = 0; g < n_shards; ++g) shard_begin[g + 1] = shard_begin[g] + count[g];
    std::vector<int64_t> cursor(shard_begin, shard_begin + n_shards);
    std::vector<int64_t> tmp((size_t)n);
    for (int64_t b = 0; b < nblk; ++b) {      // ascending inside each shard -> stays bucketed
        const int64_t lo = b * BLK, hi = std::min(n, lo + BLK);
        int64_t& cur = cursor[owner[b]];
        for (int64_t k = lo; k < hi; ++k) tmp[cur++] = order[k];
    }
    memcpy(order, tmp.data(), sizeof(int64_t) * (size_t)n);
    return BSW_OK;
}

} // extern "C"
