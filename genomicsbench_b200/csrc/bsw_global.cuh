// bsw_global.cuh -- banded global alignment with traceback -> CIGAR, sm_100a.
//
// Stands in for ksw_global2 (tools/bwa/ksw.c:502-606; push_cigar :489-500), which bwa_gen_cigar2 calls
// once per alignment region.  One alignment per thread: the row sweep over the fixed band |i - j| <= w
// keeps the reference's int32 arithmetic and its order of comparisons (the direction byte
// f << 4 | e << 2 | h of every cell depends on which side wins a tie), stores the byte matrix in
// HBM (tlen rows of n_col bytes, row pitch rounded up to 8 so that eight cells leave in one 64-bit
// store -- byte stores of 32 lanes into 32 different sectors were the first version's bound), then
// the same thread walks it back from the last cell and writes the merged operations, reversed into
// read order.
// Row state: a row only ever touches the 2w + 2 columns [i - w, i + w + 1], so eh[] is a circular
// buffer of W = 2 wmax + 2 columns per thread in shared memory (column j in slot j mod W, slots
// interleaved over the block's threads: bank = lane).  Chunks whose band does not fit shared memory
// fall back to rows in a global scratch array interleaved over the launch's threads
// (eh[j * stride + thread]); that was the first version, 8 x slower: every cell waited for an L2 round trip.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bsw {

constexpr int G_MINUS_INF = -0x40000000;         // ksw.c:487

struct GlobalDesc {                              // one alignment of a chunk
    uint32_t qoff, roff;                         // byte offsets into the chunk's gathered query / target bytes
    int32_t qlen, tlen, w, idx;                  // idx: position in the chunk's input order (threads run in length order)
    long long zoff;                              // byte offset of its direction matrix
    long long coff;                              // word offset of its (uncompacted) operation list: qlen + tlen entries
};

struct GlobalParams { int o_del, e_del, o_ins, e_ins, match, mismatch_neg, ambig; };

constexpr int GLOBAL_BLOCK = 64;

// SMEM = true: eh points nowhere, the rows are the block's dynamic shared memory, W slots per thread
template <bool SMEM>
__global__ void __launch_bounds__(GLOBAL_BLOCK)
bsw_global_kernel(const GlobalDesc* __restrict__ desc, int n, const uint8_t* __restrict__ qraw,
                  const uint8_t* __restrict__ rraw, int2* __restrict__ eh, int stride, int W, int qstride,
                  uint8_t* __restrict__ z,
                  uint32_t* __restrict__ cigar, int32_t* __restrict__ score, int32_t* __restrict__ n_cigar,
                  const GlobalParams P)
{
    extern __shared__ __align__(16) int2 g_rows[];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const GlobalDesc d = desc[t];
    const int qlen = d.qlen, tlen = d.tlen, w = d.w;
    const uint8_t* q = qraw + d.qoff;
    const uint8_t* r = rraw + d.roff;
    uint8_t* zm = z + d.zoff;
    if (SMEM) { stride = GLOBAL_BLOCK; }
    else W = 0x7fffffff;                         // no wrap: slot == column
    int2* row = SMEM ? g_rows + threadIdx.x : eh + t;      // row[slot * stride] = {h, e} of the column in that slot
    if (SMEM) {
        // the query moves to shared memory once (behind the rows; qstride = 4 x odd bytes per thread: the
        // byte loads of a warp at the same column hit 32 banks): no global load is left in the cell loop
        uint8_t* qs = reinterpret_cast<uint8_t*>(g_rows + (size_t)W * GLOBAL_BLOCK) + (size_t)threadIdx.x * qstride;
        for (int k = 0; k < qlen; ++k) qs[k] = q[k];
        q = qs;
    }
    const int oe_del = P.o_del + P.e_del, oe_ins = P.o_ins + P.e_ins;
    const int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
    const int pitch = (n_col + 7) & ~7;          // bytes per row of the direction matrix (zoff is 8-aligned)
    // first row (ksw.c:521-525).  Only columns 0 .. w are ever read from it (row i reads [i - w, i + w] and
    // leaves column i + w + 1 behind for row i + 1), so the band-sized buffer needs no more.
    row[0] = make_int2(0, G_MINUS_INF);
    int j;
    for (j = 1; j <= qlen && j <= w; ++j) row[(size_t)j * stride] = make_int2(-(P.o_ins + P.e_ins * j), G_MINUS_INF);
    if (!SMEM) for (; j <= qlen; ++j) row[(size_t)j * stride] = make_int2(G_MINUS_INF, G_MINUS_INF);
    int bslot = 0;                               // slot of column beg
    int last_h1 = 0;
    for (int i = 0; i < tlen; ++i) {                                            // ksw.c:527-589
        int f = G_MINUS_INF;
        const int beg = i > w ? i - w : 0;
        const int end = i + w + 1 < qlen ? i + w + 1 : qlen;
        int h1 = beg == 0 ? -(P.o_del + P.e_del * (i + 1)) : G_MINUS_INF;
        const int tb = r[i];
        unsigned long long* zi = reinterpret_cast<unsigned long long*>(zm + (size_t)i * pitch);
        if (i > w) bslot = bslot + 1 == W ? 0 : bslot + 1;                      // beg moved one column right
        int slot = bslot;
        // one cell (ksw.c:544-566): C = {H(i-1,j-1), E(i,j)} of the column, QB its query base; leaves the
        // direction byte in dd, the column's new {H(i,j-1), E(i+1,j)} in OUT and H(i,j) in h1
#define BSW_GLOBAL_CELL(C, QB, OUT)                                                               \
        {                                                                                         \
            int m = (C).x, e = (C).y;                                                             \
            m += (tb >= 4 || (QB) >= 4) ? P.ambig : (tb == (QB) ? P.match : P.mismatch_neg);      \
            dd = m >= e ? 0 : 1;                                                                  \
            int h = m >= e ? m : e;                                                               \
            dd = h >= f ? dd : 2;                                                                 \
            h = h >= f ? h : f;                                                                   \
            int tt = m - oe_del;                                                                  \
            e -= P.e_del;                                                                         \
            dd |= e > tt ? 1 << 2 : 0;                                                            \
            e = e > tt ? e : tt;                                                                  \
            OUT = make_int2(h1, e);                                                               \
            h1 = h;                                                                               \
            tt = m - oe_ins;                                                                      \
            f -= P.e_ins;                                                                         \
            dd |= f > tt ? 2 << 4 : 0;                                                            \
            f = f > tt ? f : tt;                                                                  \
        }
        int dd;
        int col = 0;
        j = beg;
        // eight columns at a time: their row words and query bases are loaded together (eight
        // independent loads in flight instead of one dependent load per cell), their direction bytes
        // leave in one 64-bit store
        for (; j + 8 <= end; j += 8, col += 8) {
            int2 cc[8]; int qq[8]; int ss[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                int sk = slot + k;
                sk = sk >= W ? sk - W : sk;
                ss[k] = sk;
                cc[k] = row[(size_t)sk * stride];
                qq[k] = q[j + k];
            }
            unsigned long long acc = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                int2 out;
                BSW_GLOBAL_CELL(cc[k], qq[k], out)
                row[(size_t)ss[k] * stride] = out;
                acc |= (unsigned long long)dd << (8 * k);
            }
            zi[col >> 3] = acc;
            slot += 8;
            slot = slot >= W ? slot - W : slot;
        }
        {
            unsigned long long acc = 0;
            for (; j < end; ++j, ++col) {
                int2* p = &row[(size_t)slot * stride];
                slot = slot + 1 == W ? 0 : slot + 1;
                const int2 c = *p;
                const int qb = q[j];
                int2 out;
                BSW_GLOBAL_CELL(c, qb, out)
                *p = out;
                acc |= (unsigned long long)dd << ((col & 7) * 8);
            }
            if (col & 7) zi[col >> 3] = acc;
        }
#undef BSW_GLOBAL_CELL
        int2* pe = &row[(size_t)slot * stride];                                 // eh[end] (ksw.c:588)
        *pe = make_int2(h1, G_MINUS_INF);
        last_h1 = h1;
    }
    // eh[qlen].h (ksw.c:590): the last row's end is qlen inside the supported domain (qlen <= tlen + w)
    score[d.idx] = last_h1;
    // backtrack (ksw.c:591-603)
    uint32_t* cg = cigar + d.coff;
    int n_op = 0, which = 0;
    int i = tlen - 1, k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
    auto push = [&](int op, int len) {                                          // push_cigar, ksw.c:489-500
        if (n_op == 0 || op != (int)(cg[n_op - 1] & 0xf)) cg[n_op++] = (uint32_t)len << 4 | (uint32_t)op;
        else cg[n_op - 1] += (uint32_t)len << 4;
    };
    while (i >= 0 && k >= 0) {
        which = zm[(size_t)i * pitch + (k - (i > w ? i - w : 0))] >> (which << 1) & 3;
        if (which == 0) { push(0, 1); --i; --k; }
        else if (which == 1) { push(2, 1); --i; }
        else { push(1, 1); --k; }
    }
    if (i >= 0) push(2, i + 1);
    if (k >= 0) push(1, k + 1);
    for (i = 0; i < n_op >> 1; ++i) { const uint32_t tmp = cg[i]; cg[i] = cg[n_op - 1 - i]; cg[n_op - 1 - i] = tmp; }
    n_cigar[d.idx] = n_op;
}

// operation lists of a chunk, packed one after the other: out[out_off[t] ...] = the n_cigar[t] operations of t
__global__ void __launch_bounds__(128)
bsw_cigar_compact(const GlobalDesc* __restrict__ desc, int n, const uint32_t* __restrict__ cigar,
                  const int32_t* __restrict__ n_cigar, const long long* __restrict__ out_off, uint32_t* __restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint32_t* src = cigar + desc[t].coff;
    uint32_t* dst = out + out_off[desc[t].idx];
    const int m = n_cigar[desc[t].idx];
    for (int k = 0; k < m; ++k) dst[k] = src[k];
}

// out_off[0 .. n] = exclusive prefix sum of n_cigar[0 .. n) (input order): one block, every thread sums a contiguous
// slice, the slice totals are scanned in shared memory.  Keeps the chunk on the device between the alignment kernel
// and the compaction (no host round trip for the offsets).
__global__ void __launch_bounds__(1024)
bsw_cigar_offsets(const int32_t* __restrict__ n_cigar, int n, long long* __restrict__ out_off)
{
    __shared__ long long s_tot[1024];
    const int per = (n + 1023) / 1024;
    const int lo = min(n, (int)threadIdx.x * per), hi = min(n, lo + per);
    long long sum = 0;
    for (int k = lo; k < hi; ++k) sum += n_cigar[k];
    s_tot[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const long long v = threadIdx.x >= o ? s_tot[threadIdx.x - o] : 0;
        __syncthreads();
        s_tot[threadIdx.x] += v;
        __syncthreads();
    }
    long long run = s_tot[threadIdx.x] - sum;
    for (int k = lo; k < hi; ++k) { out_off[k] = run; run += n_cigar[k]; }
    if (threadIdx.x == 1023) out_off[n] = s_tot[1023];
}

} // namespace bsw
