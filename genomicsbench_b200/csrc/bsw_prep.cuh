// bsw_prep.cuh -- device-side batch preparation of the bsw engine (sm_100a).
//
// Stands in for the host-side batch wrapper of the reference, smithWatermanBatchWrapper16
// (benchmarks/bsw/bandedSWA.cpp:1150-1431: pad, sortPairsLen :368-403, AoS->SoA transpose,
// sortPairsId :405-420), re-thought for a GPU behind PCIe: the host only streams bytes (or
// nothing at all when the caller's buffers are pinned), and these kernels
//   bsw_scan_pairs   read the caller's 72-byte SeqPair records, validate the domain, emit the
//                    16-byte descriptor per pair and the chunk summary the host plans with
//   bsw_bucket_*     counting sort of the chunk by (len2, h0, len1) -> processing order perm[];
//                    lists the pairs too long for the short kernel
//   bsw_pack_pairs   2-bit pack query / reference bytes (16 bases per word) in processing order,
//                    flag and list the pairs with N (and those outside the packed kernel's domain)
//   bsw_writeback    unpack the 16-byte results into the caller's SeqPair records (input order)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/bsw.h"
#include "bsw_kernels.cuh"
#include "bsw_kernel16.cuh"

namespace bsw {

constexpr int LEN_HIST = 1024;        // len2 histogram bins 0..1022, 1023 = everything longer

// Chunk summary: written by bsw_scan_pairs / bsw_pack_pairs (or by the host pass for pageable
// buffers), read by the host to size buffers and to plan the launches.
struct ChunkInfo {
    unsigned long long min_r, max_r, min_q, max_q;   // byte extents [min, max) of the sequences, relative to base0
    unsigned long long nominal;                      // sum len1 * len2
    unsigned long long qbases, tbases;               // sum len2 / sum len1 of the chunk
    int mn[3], mx[3];                                // ranges of len2, h0, len1 (short pairs)
    int bad;                                         // pairs outside the domain
    int n_short;                                     // pairs with len2 <= short_max
    unsigned int n_nlist, n_llist;                   // pairs with N (short) / long pairs, listed for the byte kernels
    unsigned int qcursor, tcursor;                   // packed words handed out
    int qmax_n;                                      // longest query among the N pairs
    int qmax_all;                                    // longest query of the chunk
    unsigned int hist[LEN_HIST];                     // pairs per len2
};

__device__ __forceinline__ unsigned long long bsw_umin64(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
__device__ __forceinline__ unsigned long long bsw_umax64(unsigned long long a, unsigned long long b) { return a > b ? a : b; }

// desc[i] = {query byte offset, reference byte offset (both relative to base0_q / base0_r, as
// signed 32-bit), len2 | len1 << 16, h0}.  info must be zeroed except mn[] = INT_MAX, min_* = ~0.
__global__ void __launch_bounds__(256)
bsw_scan_pairs(const SeqPair* __restrict__ pairs, int n, long long base0_r, long long base0_q,
               int match, int short_max, int4* __restrict__ desc, ChunkInfo* __restrict__ info)
{
    __shared__ unsigned int s_hist[LEN_HIST];
    __shared__ unsigned long long s_u64[7];       // min_r max_r min_q max_q nominal qbases tbases
    __shared__ int s_i[9];                        // mn[3] mx[3] bad n_short qmax_all
    for (int k = threadIdx.x; k < LEN_HIST; k += blockDim.x) s_hist[k] = 0;
    if (threadIdx.x < 7) s_u64[threadIdx.x] = (threadIdx.x == 0 || threadIdx.x == 2) ? ~0ull : 0ull;
    if (threadIdx.x < 9) s_i[threadIdx.x] = threadIdx.x < 3 ? 0x7fffffff : 0;
    __syncthreads();
    unsigned long long min_r = ~0ull, max_r = 0, min_q = ~0ull, max_q = 0, nominal = 0, qb = 0, tb = 0;
    int mn0 = 0x7fffffff, mn1 = 0x7fffffff, mn2 = 0x7fffffff, mx0 = 0, mx1 = 0, mx2 = 0, bad = 0, nshort = 0, qall = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        // 72-byte record: idr@0 idq@8 len1@24 len2@28 h0@32 (bandedSWA.h:91-100)
        const long long* p64 = reinterpret_cast<const long long*>(pairs + i);
        const int* p32 = reinterpret_cast<const int*>(pairs + i);
        const long long idr = p64[0], idq = p64[1];
        const int len1 = p32[6], len2 = p32[7], h0 = p32[8];
        const long long dr = idr - base0_r, dq = idq - base0_q;
        const bool ok = len1 >= 1 && len1 <= 32767 && len2 >= 1 && len2 <= 32767 && h0 >= 1 &&
                        (long long)h0 + (long long)len2 * match <= 32767 && idr >= 0 && idq >= 0 &&
                        dr > -(1ll << 30) && dr < (1ll << 30) && dq > -(1ll << 30) && dq < (1ll << 30);
        if (!ok) { bad = 1; desc[i] = make_int4(0, 0, 1 | (1 << 16), 1); continue; }
        desc[i] = make_int4((int)dq, (int)dr, len2 | (len1 << 16), h0);
        nominal += (unsigned long long)len1 * (unsigned long long)len2;
        const unsigned long long ur = (unsigned long long)(dr + (1ll << 30)), uq = (unsigned long long)(dq + (1ll << 30));
        min_r = bsw_umin64(min_r, ur); max_r = bsw_umax64(max_r, ur + len1);
        min_q = bsw_umin64(min_q, uq); max_q = bsw_umax64(max_q, uq + len2);
        atomicAdd(&s_hist[min(len2, LEN_HIST - 1)], 1u);
        qall = max(qall, len2);
        qb += len2; tb += len1;
        if (len2 <= short_max) {
            ++nshort;
            mn0 = min(mn0, len2); mx0 = max(mx0, len2);
            mn1 = min(mn1, h0);   mx1 = max(mx1, h0);
            mn2 = min(mn2, len1); mx2 = max(mx2, len1);
        }
    }
    atomicMin(&s_u64[0], min_r); atomicMax(&s_u64[1], max_r);
    atomicMin(&s_u64[2], min_q); atomicMax(&s_u64[3], max_q);
    atomicAdd(&s_u64[4], nominal); atomicAdd(&s_u64[5], qb); atomicAdd(&s_u64[6], tb);
    atomicMin(&s_i[0], mn0); atomicMin(&s_i[1], mn1); atomicMin(&s_i[2], mn2);
    atomicMax(&s_i[3], mx0); atomicMax(&s_i[4], mx1); atomicMax(&s_i[5], mx2);
    if (bad) atomicAdd(&s_i[6], 1);
    atomicAdd(&s_i[7], nshort);
    atomicMax(&s_i[8], qall);
    __syncthreads();
    for (int k = threadIdx.x; k < LEN_HIST; k += blockDim.x)
        if (s_hist[k]) atomicAdd(&info->hist[k], s_hist[k]);
    if (threadIdx.x == 0) {
        atomicMin(&info->min_r, s_u64[0]); atomicMax(&info->max_r, s_u64[1]);
        atomicMin(&info->min_q, s_u64[2]); atomicMax(&info->max_q, s_u64[3]);
        atomicAdd(&info->nominal, s_u64[4]); atomicAdd(&info->qbases, s_u64[5]); atomicAdd(&info->tbases, s_u64[6]);
        atomicMin(&info->mn[0], s_i[0]); atomicMin(&info->mn[1], s_i[1]); atomicMin(&info->mn[2], s_i[2]);
        atomicMax(&info->mx[0], s_i[3]); atomicMax(&info->mx[1], s_i[4]); atomicMax(&info->mx[2], s_i[5]);
        if (s_i[6]) atomicAdd(&info->bad, s_i[6]);
        atomicAdd(&info->n_short, s_i[7]);
        atomicMax(&info->qmax_all, s_i[8]);
    }
}

// 16 base codes (one per byte) at any alignment -> one 2-bit word.  *bad collects codes > 3.
// `room` = bytes of the sequence from src on: the word-wise path may touch up to 3 bytes past its
// 16 and is only taken when they still belong to the sequence (the buffer may be host memory
// that ends with it).
__device__ __forceinline__ uint32_t bsw_pack16(const uint8_t* src, int nb, int room, uint32_t& bad)
{
    uint32_t out = 0;
    if (room >= 20) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(src);
        const uint32_t* aw = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
        const unsigned sh = (unsigned)(a & 3) * 8u;
        uint32_t w[5];
#pragma unroll
        for (int k = 0; k < 4; ++k) w[k] = __ldg(aw + k);
        w[4] = sh ? __ldg(aw + 4) : 0u;              // never reads past the 16 bytes when aligned
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t v = __funnelshift_r(w[k], w[k + 1], sh);
            bad |= v & 0xFCFCFCFCu;
            v &= 0x03030303u;
            v = (v | (v >> 6)) & 0x000F000Fu;
            v = (v | (v >> 12)) & 0xFFu;
            out |= v << (8 * k);
        }
    } else {
        for (int k = 0; k < nb; ++k) {
            const uint32_t c = __ldg(src + k);
            bad |= c & 0xFCu;
            out |= (c & 3u) << (2 * k);
        }
    }
    return out;
}

// One warp packs the pairs at 32 consecutive positions of the processing order (perm[]), so that
// the DP kernel later reads descriptors and sequences of neighbouring threads from neighbouring
// addresses.  The word counts are scanned across the warp, one atomicAdd per sequence kind
// reserves the output range, then the lanes sweep the concatenated word list (lane -> word, pair
// found by binary search), so loads stay coalesced for short and long sequences alike.
// meta[s] (processing order) gets word offsets; desc[i] (input order) keeps byte offsets.
__global__ void __launch_bounds__(256)
bsw_pack_pairs(const int4* __restrict__ desc, const uint32_t* __restrict__ perm, int n_sorted,
               const uint8_t* __restrict__ qraw, const uint8_t* __restrict__ rraw, int4* __restrict__ meta,
               uint32_t* __restrict__ qpk, uint32_t* __restrict__ tpk,
               uint32_t* __restrict__ nlist, ChunkInfo* __restrict__ info, int packed16_match)
{
    __shared__ uint32_t s_pre[8][2][33];
    __shared__ uint32_t s_bad[8][32];
    __shared__ int4 s_desc[8][32];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    for (int base = (blockIdx.x * (blockDim.x >> 5) + wib) * 32; base < n_sorted; base += nwarps * 32) {
        const int s = base + lane;
        int4 d = make_int4(0, 0, 0, 0);
        int len2 = 0, len1 = 0, pi = 0;
        if (s < n_sorted) {
            pi = (int)perm[s];
            d = desc[pi];
            len2 = d.z & 0xffff; len1 = (d.z >> 16) & 0xffff;
        }
        s_desc[wib][lane] = d;
        const uint32_t nq = (uint32_t)(len2 + 15) >> 4;
        const uint32_t nt = (uint32_t)(len1 + 15) >> 4;
        uint32_t pq = nq, pt = nt;                  // inclusive warp scans
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t a = __shfl_up_sync(FULL, pq, o), b = __shfl_up_sync(FULL, pt, o);
            if (lane >= o) { pq += a; pt += b; }
        }
        const uint32_t totq = __shfl_sync(FULL, pq, 31), tott = __shfl_sync(FULL, pt, 31);
        uint32_t bq = 0, bt = 0;
        if (lane == 0) { bq = atomicAdd(&info->qcursor, totq); bt = atomicAdd(&info->tcursor, tott); }
        bq = __shfl_sync(FULL, bq, 0); bt = __shfl_sync(FULL, bt, 0);
        s_pre[wib][0][lane + 1] = pq; s_pre[wib][1][lane + 1] = pt;
        if (lane == 0) { s_pre[wib][0][0] = 0; s_pre[wib][1][0] = 0; }
        s_bad[wib][lane] = 0;
        __syncwarp();
#pragma unroll
        for (int kind = 0; kind < 2; ++kind) {
            const uint32_t tot = kind ? tott : totq;
            const uint32_t* pre = s_pre[wib][kind];
            for (uint32_t wv = lane; wv < tot; wv += 32) {
                int lo = 0;                          // largest p with pre[p] <= wv
#pragma unroll
                for (int step = 16; step > 0; step >>= 1)
                    if (pre[lo + step] <= wv) lo += step;
                const int4 dd = s_desc[wib][lo];
                const int len = kind ? (dd.z >> 16) & 0xffff : dd.z & 0xffff;
                const uint32_t wi = wv - pre[lo];
                const uint8_t* src = (kind ? rraw + dd.y : qraw + dd.x) + 16 * (int)wi;
                uint32_t bad = 0;
                const int room = len - 16 * (int)wi;
                const uint32_t word = bsw_pack16(src, min(16, room), room, bad);
                (kind ? tpk + bt : qpk + bq)[wv] = word;
                if (bad) atomicOr(&s_bad[wib][lo], 1u);
            }
        }
        __syncwarp();
        if (s < n_sorted) {
            // pairs for the 32-bit byte kernel: those that contain N and, when the packed 16-bit kernel
            // runs the rest (packed16_match = its match score), those outside its score domain
            const bool has_n = s_bad[wib][lane] != 0 ||
                               (packed16_match > 0 && !k16::eligible(packed16_match, len2, d.w & 0xffff));
            meta[s] = make_int4((int)(bq + pq - nq), (int)(bt + pt - nt), d.z, d.w | (has_n ? BSW_META_NFLAG : 0));
            if (has_n) { nlist[atomicAdd(&info->n_nlist, 1u)] = (uint32_t)pi; atomicMax(&info->qmax_n, len2); }
        }
        __syncwarp();
    }
}

// ---- bucketing: counting sort by the compressed key (len2 | h0 | len1) >> drop --------------
struct BucketKey {
    int mn2, mnh, mn1;        // minima of len2, h0, len1
    int b_h0, b_l1;           // bit widths of the h0 and len1 fields
    int drop;                 // low bits dropped so that the key fits the bin table
    int short_max;
};

__device__ __forceinline__ uint32_t bsw_bucket_of(const BucketKey& K, const int4 d)
{
    const uint32_t len2 = d.z & 0xffff, len1 = (d.z >> 16) & 0xffff, h0 = d.w & 0xffff;
    const uint32_t key = ((len2 - K.mn2) << (K.b_h0 + K.b_l1)) | ((h0 - K.mnh) << K.b_l1) | (len1 - K.mn1);
    return key >> K.drop;
}

// rank[i] = arrival order of pair i inside its bin; bins[] accumulates the bin sizes; pairs whose
// query is too long for the short kernel are listed for the warp-per-pair kernel instead
__global__ void __launch_bounds__(256)
bsw_bucket_count(const int4* __restrict__ desc, int n, const __grid_constant__ BucketKey K,
                 uint32_t* __restrict__ bins, uint32_t* __restrict__ rank,
                 uint32_t* __restrict__ llist, ChunkInfo* __restrict__ info)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int4 d = desc[i];
        if ((d.z & 0xffff) > K.short_max) { llist[atomicAdd(&info->n_llist, 1u)] = (uint32_t)i; continue; }
        rank[i] = atomicAdd(&bins[bsw_bucket_of(K, d)], 1u);
    }
}

// Exclusive prefix sum of the bin table in two levels.  bsw_bucket_scan_tiles: every block scans one
// tile of 1024 bins in place and leaves the tile total; bsw_bucket_scan_totals: one block scans the
// (<= 1024) tile totals; the scatter adds the two.
constexpr int SCAN_TILE = 1024;

__global__ void __launch_bounds__(256)
bsw_bucket_scan_tiles(uint32_t* __restrict__ bins, int nbins, uint32_t* __restrict__ totals)
{
    __shared__ uint32_t s_warp[8];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * 4;
    uint32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = base + k < nbins ? bins[base + k] : 0u;
    const uint32_t mine = v[0] + v[1] + v[2] + v[3];
    uint32_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) s_warp[wib] = inc;
    __syncthreads();
    uint32_t wbase = 0;
    for (int k = 0; k < wib; ++k) wbase += s_warp[k];
    uint32_t run = wbase + inc - mine;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (base + k < nbins) bins[base + k] = run;
        run += v[k];
    }
    if (threadIdx.x == 255) totals[blockIdx.x] = run;
}

__global__ void __launch_bounds__(1024)
bsw_bucket_scan_totals(uint32_t* __restrict__ totals, int ntiles)
{
    __shared__ uint32_t s_part[1024];
    const uint32_t mine = (int)threadIdx.x < ntiles ? totals[threadIdx.x] : 0u;
    s_part[threadIdx.x] = mine;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {            // Hillis-Steele inclusive scan
        const uint32_t v = (int)threadIdx.x >= o ? s_part[threadIdx.x - o] : 0u;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    if ((int)threadIdx.x < ntiles) totals[threadIdx.x] = s_part[threadIdx.x] - mine;
}

__global__ void __launch_bounds__(256)
bsw_bucket_scatter(const int4* __restrict__ desc, int n, const __grid_constant__ BucketKey K,
                   const uint32_t* __restrict__ bins, const uint32_t* __restrict__ totals,
                   const uint32_t* __restrict__ rank, uint32_t* __restrict__ perm)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int4 d = desc[i];
        if ((d.z & 0xffff) > K.short_max) continue;
        const uint32_t b = bsw_bucket_of(K, d);
        perm[bins[b] + totals[b / SCAN_TILE] + rank[i]] = (uint32_t)i;
    }
}

// res[i] (8 x int16) -> the six int32 result fields of the caller's record i (score@44 tle@48
// gtle@52 qle@56 gscore@60 max_off@64, bandedSWA.h:91-100); pairs may be host-mapped memory.
__global__ void __launch_bounds__(256)
bsw_writeback(const int4* __restrict__ res, int n, SeqPair* __restrict__ pairs)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int4 v = res[i];
        int* out = reinterpret_cast<int*>(pairs + i) + 11;
        out[0] = (int)(short)(v.x & 0xffff);        // score
        out[1] = (int)(short)(v.y & 0xffff);        // tle
        out[2] = (int)(short)(v.y >> 16);           // gtle
        out[3] = (int)(short)(v.x >> 16);           // qle
        out[4] = (int)(short)(v.z & 0xffff);        // gscore
        out[5] = (int)(short)(v.z >> 16);           // max_off
    }
}

} // namespace bsw
