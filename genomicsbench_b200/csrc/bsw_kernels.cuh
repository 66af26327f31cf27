// bsw_kernels.cuh -- hand-written sm_100a kernels of the banded Smith-Waterman extension.
//
// Per-pair semantics: SURVEY.md Appendix A == BandedPairWiseSW::scalarBandedSWA
// (benchmarks/bsw/bandedSWA.cpp:128-249; canonical twin tools/bwa/ksw.c:380-479) with the
// z-drop rule of the vector kernel the benchmark runs (ZSCORE16, bandedSWA.cpp:323-336),
// i.e. exactly what getScores16 (bandedSWA.cpp:1124-1148 -> smithWaterman256_16 :1433-1831)
// returns for a pair run alone in its SIMD group.
//
// Two kernels replace the reference's 8-bit / 16-bit AVX dispatch:
//   bsw_short_kernel  thread-per-pair (inter-sequence), the eh[] row lives in shared memory
//                     as one packed word per cell, DPX (VIADDMNMX / VIMNMX3) recurrence.
//   bsw_long_kernel   warp-per-pair row sweep: lanes own column strips of the band window,
//                     __shfl_sync hands H / F between lanes, a 5-step max-plus scan resolves F.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

// functions shared with the host-side emulation of the row sweeps (tests/k16_emu.cu)
#define BSW_HD __host__ __device__ __forceinline__

namespace bsw {

struct KParams {
    int match;          // +a
    int mismatch_neg;   // -b
    int ambig;          // score of any cell touching an N
    int o_del, e_del, o_ins, e_ins;
    int oe_del, oe_ins;
    int zdrop, end_bonus;
    int zmode;          // BSW_ZDROP_VECTOR / BSW_ZDROP_SCALAR
    int mx;             // max entry of the scoring matrix (band clamp, bandedSWA.cpp:160-168)
    int w;              // caller's band width
    int kone;           // the constant 1, kept opaque to the compiler (argmax-key addends stay in registers)
};

// Shared-memory cell of the short kernel: one 32-bit word per DP column,  [31:16] e  [15:0] h
// (h, e < 32768 is the reference's own 16-bit domain, bandedSWA.h:84 / Q7), so both halves are
// read with 16-bit loads and need no unpacking on the ALU pipe.
// meta.w = h0 | BSW_META_NFLAG when the pair contains an N: the 2-bit variant skips it (the
// byte variant recomputes it from the byte-staged copy)
#define BSW_META_NFLAG (1 << 30)

// band clamp of bandedSWA.cpp:160-168, same double arithmetic
BSW_HD int bsw_clamp_band(const KParams& P, int qlen)
{
    int w = P.w;
    int max_ins = (int)((double)(qlen * P.mx + P.end_bonus - P.o_ins) / P.e_ins + 1.);
    max_ins = max_ins > 1 ? max_ins : 1;
    w = w < max_ins ? w : max_ins;
    int max_del = (int)((double)(qlen * P.mx + P.end_bonus - P.o_del) / P.e_del + 1.);
    max_del = max_del > 1 ? max_del : 1;
    w = w < max_del ? w : max_del;
    return w;
}

// a * b + c on the FMA pipe (IMAD): keeps adds / shifts / re-packs off the ALU pipe, which bounds
// these kernels (VIADDMNMX, VIMNMX3, LOP3, SHF, SEL all issue there at half rate).
__device__ __forceinline__ int bsw_mad(int a, int b, int c)
{
    int r;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// Shared-memory accessors of the short kernel's row sweep: one 128-bit load / store moves a whole
// 4-column group.  "memory" keeps them ordered against the plain C++ accesses to the same array;
// among themselves they stay in program order (volatile).
__device__ __forceinline__ uint4 bsw_lds_u128(uint32_t saddr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ void bsw_sts_u128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t bsw_lds_u8(uint32_t saddr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
    return v;
}
// x >> 16 on the FMA pipe (IMAD.HI.U32 with a register multiplier of 65536; two issue slots there)
__device__ __forceinline__ uint32_t bsw_hi16_fma(uint32_t x, uint32_t k65536)
{
    uint32_t r;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(k65536));
    return r;
}

// Running per-pair state shared by both kernels' row epilogues.
struct PairState {
    int max, max_i, max_j, max_ie, gscore, max_off;
};

// Row epilogue: global max / max_off / z-drop (bandedSWA.cpp:218-228, :323-336).
// Returns true when the row loop must stop.
BSW_HD bool bsw_row_update(const KParams& P, PairState& st, int i, int m, int mj)
{
    if (m == 0) return true;
    if (m > st.max) {
        st.max = m; st.max_i = i; st.max_j = mj;
        int d = mj - i; d = d < 0 ? -d : d;
        st.max_off = st.max_off > d ? st.max_off : d;
        return false;
    }
    const int di = i - st.max_i, dj = mj - st.max_j;
    if (P.zmode == 0) {
        const int gap = di > dj ? di - dj : dj - di;
        return st.max - m - gap > P.zdrop;
    }
    if (P.zdrop > 0) {
        if (di > dj) return st.max - m - (di - dj) * P.e_del > P.zdrop;
        return st.max - m - (dj - di) * P.e_ins > P.zdrop;
    }
    return false;
}

BSW_HD int4 bsw_pack_result(const PairState& st)
{
    // 8 x int16: score, qle, tle, gtle, gscore, max_off, 0, 0
    int4 r;
    r.x = (st.max & 0xffff) | ((st.max_j + 1) << 16);
    r.y = ((st.max_i + 1) & 0xffff) | ((st.max_ie + 1) << 16);
    r.z = (st.gscore & 0xffff) | (st.max_off << 16);
    r.w = 0;
    return r;
}

// ---------------------------------------------------------------------------------------
// Short-pair kernel: one pair per thread.
//   Position p of the processing order handles pair s = perm[first + p] of the chunk; res[s]
//   receives its packed result (input order).  Descriptor {query word/byte offset, target
//   word/byte offset, qlen | tlen << 16, h0 | flags}: meta[first + p] for the 2-bit variant (packed
//   in processing order: neighbouring threads read neighbouring words), meta[s] for the byte variant.
//   BYTESEQ = false: sequences are 2-bit packed, 16 bases per 32-bit word, word-aligned.
//   BYTESEQ = true : one base code per byte (pairs that contain N, code 4).
//   Shared memory of a block (S = qstride = words per thread, S / 4 odd, S >= qlen + 8: the
//   pipelined sweep reads at most one group past the window, i.e. up to column qlen + 3):
//     eh [tid * S + j]                cell j of thread tid: e << 16 | h.  A 128-bit access moves
//                                     columns j..j+3; with S / 4 odd the 8 lanes of a quarter warp
//                                     hit 8 distinct 16-byte bank groups (conflict-free)
//     qpk[(j >> 2) * BLOCK + tid]     byte holding query bases j..j+3, 2 bits each (2-bit variant)
//   Row sweep: columns are processed in groups of 4 aligned to j % 4 == 0: one LDS.128, one LDS.U8
//   (the group's four query bases; the match tests are single LOP3s with immediate masks) and one
//   STS.128 per group.  Columns left of `beg` are dead for the rest of the pair (beg never
//   decreases), so the <= 3 columns between the group boundary and beg are zeroed and swept like
//   live ones: they produce h = e = f = 0, exactly the state the reference enters column beg
//   with when beg > 0.  The <= 3 columns right of the last full group run the scalar tail loop.
//   Instruction budget per cell (scripts/int_pipe_probe.cu: ALU pipe 2 warp-instructions/clk/SM,
//   FMA pipe 2, IMAD.HI 1): ALU = match test, select, M, h, E', F' (+ the e unpack of two cells in
//   four, + half of the dual-issue ops); FMA = cap, h unpack, e unpack of the other two cells, the
//   gap decrement, the cell re-pack and the argmax key.  ~7 ALU + ~6.5 FMA + 0.75 LSU slots.
// ---------------------------------------------------------------------------------------
template <int BLOCK, bool BYTESEQ>
__global__ void __launch_bounds__(BLOCK)
bsw_short_kernel(const int4* __restrict__ meta, const uint32_t* __restrict__ perm,
                 const uint32_t* __restrict__ qseq, const uint32_t* __restrict__ tseq,
                 int4* __restrict__ res, int first, int count, int qstride,
                 const __grid_constant__ KParams P, unsigned long long* __restrict__ cell_counter)
{
    extern __shared__ __align__(16) uint32_t eh_smem[];
    const int tid = threadIdx.x;
    const int local = blockIdx.x * BLOCK + tid;
    long long my_cells = 0;
    int4 md = make_int4(0, 0, 0, 0);
    int s = 0;                        // the pair's index in the chunk (input order)
    bool run = local < count;
    if (run) {
        s = (int)perm[first + local];
        md = BYTESEQ ? meta[s] : meta[first + local];
        if (!BYTESEQ && (md.w & BSW_META_NFLAG)) run = false;
    }
    if (run) {
        const int qlen = md.z & 0xffff, tlen = (md.z >> 16) & 0xffff, h0 = md.w & 0xffff;
        uint32_t* const eh = eh_smem + tid * qstride;
        uint8_t* const qpk = reinterpret_cast<uint8_t*>(eh_smem + qstride * BLOCK) + tid;
        const uint32_t eh_sa = (uint32_t)__cvta_generic_to_shared(eh);
        const uint32_t qpk_sa = (uint32_t)__cvta_generic_to_shared(qpk);
        // byte variant: signed byte offsets (sequences are read in place, before or after offset 0)
        const uint8_t* qb = reinterpret_cast<const uint8_t*>(qseq) + (BYTESEQ ? (ptrdiff_t)md.x : (ptrdiff_t)0);
        const uint8_t* tb = reinterpret_cast<const uint8_t*>(tseq) + (BYTESEQ ? (ptrdiff_t)md.y : (ptrdiff_t)0);
        const uint32_t* qw = qseq + (BYTESEQ ? 0u : (uint32_t)md.x);
        const uint32_t* tw = tseq + (BYTESEQ ? 0u : (uint32_t)md.y);

        // ---- first row (bandedSWA.cpp:155-157), four columns per store; the query goes to its byte plane
        {
            int hv = h0;
            for (int j0 = 0; j0 <= qlen; j0 += 4) {
                uint32_t v[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int j = j0 + c;
                    if (j == 1) hv = h0 > P.oe_ins ? h0 - P.oe_ins : 0;
                    else if (j >= 2) hv = hv > P.e_ins ? hv - P.e_ins : 0;
                    v[c] = j <= qlen ? (uint32_t)hv : 0u;
                }
                bsw_sts_u128(eh_sa + 4u * (uint32_t)j0, v[0], v[1], v[2], v[3]);
            }
            if (!BYTESEQ) {
                const int nw = (qlen + 15) >> 4;
                for (int k = 0; k < nw; ++k) {
                    const uint32_t v = __ldg(qw + k);
                    uint8_t* d = qpk + 4 * k * BLOCK;
                    d[0] = (uint8_t)v; d[BLOCK] = (uint8_t)(v >> 8);
                    d[2 * BLOCK] = (uint8_t)(v >> 16); d[3 * BLOCK] = (uint8_t)(v >> 24);
                }
            }
        }
        const int w = bsw_clamp_band(P, qlen);

        PairState st;
        st.max = h0; st.max_i = -1; st.max_j = -1; st.max_ie = -1; st.gscore = -1; st.max_off = 0;
        int beg = 0, end = qlen;
        uint32_t tword = 0;
        // Loop constants.  `zero` is 0 but opaque to the compiler (bit 31 of a loaded word that never
        // has it set): adding it pins the constants in registers -- otherwise every group iteration
        // re-reads them from the constant bank and the first dependent instruction stalls on it.
        const int zero = (int)((uint32_t)md.z >> 31);
        const int neg_oe_del = zero - P.oe_del, neg_oe_ins = zero - P.oe_ins;
        const int neg_e_del = zero - P.e_del, neg_e_ins = zero - P.e_ins;
        const int c_match = zero + P.match, c_mis = zero + P.mismatch_neg;
        // argmax-key addends kept in registers so that the key is one IMAD (h * 65536 + k)
        const int k1 = zero + P.kone, k2 = zero + 2 * P.kone, k3 = zero + 3 * P.kone;
        const uint32_t k65536 = (uint32_t)(zero + 65536 * P.kone);

        for (int i = 0; i < tlen; ++i) {
            int ti;
            if (BYTESEQ) ti = __ldg(tb + i);
            else {
                if ((i & 15) == 0) tword = __ldg(tw + (i >> 4));
                ti = (tword >> ((i & 15) * 2)) & 3;
            }
            beg = max(beg, i - w);
            end = min(min(end, i + w + 1), qlen);
            int h1 = 0;
            if (beg == 0) h1 = max(h0 - (P.o_del + P.e_del * (i + 1)), 0);
            int f = 0;
            int mkey = 0;                                    // (row max << 16) | argmax column
            // DP recurrence of one cell from its packed word WD = E(i, j) << 16 | H(i-1, j-1) and the
            // match score SC; NW receives the new word H(i, j-1) | E(i+1, j) << 16
            // (bandedSWA.cpp:196-210).  EXPR_E unpacks e on the ALU pipe (shift) or the FMA pipe.
#define BSW_CELL_CORE(NW, WD, EXPR_E, SC, KEYADD)                                                    \
            {                                                                                        \
                const int e = (int)(EXPR_E);                                                         \
                const int Hd = bsw_mad(e, -65536, (int)(WD));            /* low half, on the FMA pipe */ \
                /* M = Hd ? Hd + s : 0, clamped at 0 (a negative M is equivalent to 0 in every use) */ \
                const int M = __viaddmin_s32_relu(Hd, (SC), bsw_mad((int)(WD), 65536, 0));           \
                const int h = __vimax3_s32(M, e, f);                                                 \
                const int en = __viaddmax_s32_relu(M, neg_oe_del, bsw_mad(e, 1, neg_e_del));         \
                f = __viaddmax_s32_relu(M, neg_oe_ins, f + neg_e_ins);                               \
                NW = (uint32_t)bsw_mad(en, 65536, h1);                                               \
                mkey = max(mkey, bsw_mad(h, 65536, (KEYADD)));                                       \
                h1 = h;                                                                              \
            }
            int j;
            if (BYTESEQ) {
                j = beg;
            } else {
                // zero the dead columns between the group boundary and beg, then sweep full groups
                j = beg & ~3;
                if (j + 0 < beg) eh[j + 0] = 0u;
                if (j + 1 < beg) eh[j + 1] = 0u;
                if (j + 2 < beg) eh[j + 2] = 0u;
                const uint32_t trep = (uint32_t)ti * 0x55u;
                uint32_t sa = eh_sa + 4u * (uint32_t)j;
                uint32_t qa = qpk_sa + (uint32_t)(j >> 2) * BLOCK;
                // software pipeline: group g+1 is loaded while group g is computed (the row buffer
                // is padded, so the load past the last full group stays inside this thread's row)
                uint4 cur = bsw_lds_u128(sa);
                uint32_t qv = bsw_lds_u8(qa);
                // one 4-column group: X = query byte ^ target pattern, CUR = the group's four words
#define BSW_GROUP(X, CUR, JBASE)                                                                     \
                {                                                                                    \
                    uint32_t n0, n1, n2, n3;                                                         \
                    int mk4;                                                                         \
                    {                                                                                \
                        int mkey = 0;                                                                \
                        BSW_CELL_CORE(n0, CUR.x, CUR.x >> 16,                 ((X) & 0x03u) ? c_mis : c_match, 0)  \
                        BSW_CELL_CORE(n1, CUR.y, bsw_hi16_fma(CUR.y, k65536), ((X) & 0x0cu) ? c_mis : c_match, k1) \
                        BSW_CELL_CORE(n2, CUR.z, CUR.z >> 16,                 ((X) & 0x30u) ? c_mis : c_match, k2) \
                        BSW_CELL_CORE(n3, CUR.w, bsw_hi16_fma(CUR.w, k65536), ((X) & 0xc0u) ? c_mis : c_match, k3) \
                        mk4 = mkey;                                                                  \
                    }                                                                                \
                    bsw_sts_u128(sa, n0, n1, n2, n3);                                                \
                    mkey = max(mkey, mk4 + (JBASE));          /* same order: j + k < 65536 never carries */ \
                    sa += 16;                                                                        \
                }
                // two groups per iteration while they last (more independent work per warp: these
                // kernels run at 1-3 warps per scheduler because of the shared-memory footprint)
                for (; j + 8 <= end; j += 8) {
                    const uint32_t x0 = qv ^ trep;
                    const uint4 cur2 = bsw_lds_u128(sa + 16);
                    const uint32_t x1 = bsw_lds_u8(qa + BLOCK) ^ trep;
                    const uint4 nxt = bsw_lds_u128(sa + 32);
                    qa += 2 * BLOCK;
                    qv = bsw_lds_u8(qa);
                    BSW_GROUP(x0, cur, j)
                    BSW_GROUP(x1, cur2, j + 4)
                    cur = nxt;
                }
                if (j + 4 <= end) {
                    const uint32_t x = qv ^ trep;
                    BSW_GROUP(x, cur, j)
                    j += 4;
                }
#undef BSW_GROUP
            }
            // scalar tail (2-bit variant: <= 3 columns) / whole window (byte variant)
            for (; j < end; ++j) {
                const uint32_t wd = eh[j];
                int sc;
                if (BYTESEQ) {
                    const int qj = __ldg(qb + j);
                    sc = (qj > 3 || ti > 3) ? P.ambig : (qj == ti ? c_match : c_mis);
                } else {
                    const int qj = ((int)qpk[(j >> 2) * BLOCK] >> ((j & 3) * 2)) & 3;
                    sc = qj == ti ? c_match : c_mis;
                }
                uint32_t nw;
                BSW_CELL_CORE(nw, wd, wd >> 16, sc, j)
                eh[j] = nw;
            }
#undef BSW_CELL_CORE
            if (end > beg) my_cells += end - beg;
            // eh[end] = {h1, 0}  (bandedSWA.cpp:213)
            eh[end] = (uint32_t)h1;
            const int jfin = end > beg ? end : beg;
            if (jfin == qlen) {                               // bandedSWA.cpp:214-217
                if (!(st.gscore > h1)) st.max_ie = i;
                st.gscore = max(st.gscore, h1);
            }
            const int m = mkey >> 16, mj = mkey & 0xffff;
            if (bsw_row_update(P, st, i, m, mj)) break;
            // next row's window (bandedSWA.cpp:230-233)
            {
                int jj = beg;
                while (jj < end && eh[jj] == 0u) ++jj;
                beg = jj;
                jj = end;
                while (jj >= beg && eh[jj] == 0u) --jj;
                end = min(jj + 2, qlen);
            }
        }
        res[s] = bsw_pack_result(st);
    }
    // effective-cell statistic: one atomic per warp
    for (int off = 16; off > 0; off >>= 1) my_cells += __shfl_down_sync(0xffffffffu, my_cells, off);
    if ((tid & 31) == 0 && my_cells) atomicAdd(cell_counter, (unsigned long long)my_cells);
}

// ---------------------------------------------------------------------------------------
// Long-pair kernel: one pair per warp, row sweep.
//   Each DP row's window [beg, end) is cut into chunks of 128 cells; lane l owns 4 consecutive
//   cells of the chunk.  M and E are elementwise in the previous row; F is the max-plus prefix
//   F(j+1) = max(F(j) - e_ins, M(j) - oe_ins, 0): every lane sweeps its 4 cells with F_in = 0,
//   a 5-step __shfl_up_sync scan combines the lanes' outgoing values (decay 4*e_ins per lane),
//   and each cell takes max(local, F_in - k*e_ins).  The diagonal H hand-off between lanes
//   and chunks is one __shfl_up_sync.  Rows stay strictly sequential because the window,
//   the m == 0 exit and z-drop need the complete previous row (SURVEY.md finding 0.6).
//   Sequences are one base per byte (codes 0-4), read in place at any alignment; eh[] (h | e << 16 per cell)
//   lives in a per-warp row: shared memory when the launch's longest query fits (SMEM = true: a row
//   sweep then costs no global round trip, which is what makes this the low-latency kernel for
//   batches too small to fill the machine one pair per thread), else a global scratch row that
//   stays L1/L2 resident.
//   Pairs are pulled from an atomic queue, longest first.
// ---------------------------------------------------------------------------------------
constexpr int LONG_WARPS = 4;

template <bool SMEM>
__global__ void __launch_bounds__(LONG_WARPS * 32)
bsw_long_kernel(const int4* __restrict__ meta, const uint32_t* __restrict__ perm,
                const uint8_t* __restrict__ qbytes, const uint8_t* __restrict__ tbytes,
                int4* __restrict__ res, int count, const __grid_constant__ KParams P,
                uint32_t* __restrict__ scratch, int scratch_stride, unsigned int* __restrict__ queue,
                unsigned long long* __restrict__ cell_counter)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * LONG_WARPS + (threadIdx.x >> 5);
    extern __shared__ __align__(16) uint32_t long_rows[];
    uint32_t* const eh = SMEM ? long_rows + (size_t)(threadIdx.x >> 5) * scratch_stride
                              : scratch + (size_t)gwarp * scratch_stride;
    const int e_ins4 = 4 * P.e_ins;
    long long my_cells = 0, pair_cells0 = 0;

    for (;;) {
        unsigned k = 0;
        if (lane == 0) k = atomicAdd(queue, 1u);
        k = __shfl_sync(FULL, k, 0);
        if (k >= (unsigned)count) break;
        const int s = perm ? (int)perm[k] : (int)k;       // (no list: every pair of the call, in input order -- the latency route)
        const int4 md = meta[s];
        const int qlen = md.z & 0xffff, tlen = (md.z >> 16) & 0xffff, h0 = md.w & 0xffff;
        const uint8_t* qb = qbytes + (ptrdiff_t)md.x;     // signed: sequences are read in place
        const uint8_t* tb = tbytes + (ptrdiff_t)md.y;

        // first row, closed form of bandedSWA.cpp:155-157: eh[j].h = max(h0 - oe_ins - (j-1)*e_ins, 0)
        for (int j = lane; j <= qlen + 4; j += 32) {
            int hv = 0;
            if (j == 0) hv = h0;
            else if (j <= qlen) hv = max(h0 - P.oe_ins - (j - 1) * P.e_ins, 0);
            eh[j] = (uint32_t)hv;
        }
        __syncwarp();
        const int w = bsw_clamp_band(P, qlen);
        PairState st;
        st.max = h0; st.max_i = -1; st.max_j = -1; st.max_ie = -1; st.gscore = -1; st.max_off = 0;
        int beg = 0, end = qlen;

        for (int i = 0; i < tlen; ++i) {
            const int ti = __ldg(tb + i);
            beg = max(beg, i - w);
            end = min(min(end, i + w + 1), qlen);
            const int h1_init = beg == 0 ? max(h0 - (P.o_del + P.e_del * (i + 1)), 0) : 0;
            int f_carry = 0, h_carry = h1_init;
            unsigned mkey = 0;
            int h_end = -1;                                 // H(i, end-1), found by its owner lane
            if (end > beg) {
                for (int cbase = beg & ~3; cbase <= end; cbase += 128) {
                    const int j0 = cbase + 4 * lane;
                    uint4 wd = make_uint4(0, 0, 0, 0);
                    uint32_t qw = 0;
                    if (j0 <= end) {
                        wd = *reinterpret_cast<const uint4*>(eh + j0);
                        if (j0 < qlen) {
                            // 4 query bases at any alignment (sequences are read in place)
                            const uint8_t* pq = qb + j0;
                            if (j0 + 8 <= qlen) {
                                const uintptr_t a = reinterpret_cast<uintptr_t>(pq);
                                const uint32_t* aw = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
                                qw = __funnelshift_r(__ldg(aw), __ldg(aw + 1), (unsigned)(a & 3) * 8u);
                            } else {
#pragma unroll
                                for (int c = 0; c < 4; ++c)
                                    if (j0 + c < qlen) qw |= (uint32_t)__ldg(pq + c) << (8 * c);
                            }
                        }
                    }
                    const uint32_t wds[4] = {wd.x, wd.y, wd.z, wd.w};
                    int M[4], E[4], g[5];
                    g[0] = 0;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int j = j0 + c;
                        const bool in = j >= beg && j < end;
                        const int Hd = in ? (int)(wds[c] & 0xffffu) : 0;
                        E[c] = in ? (int)(wds[c] >> 16) : 0;
                        const int qk = (int)((qw >> (8 * c)) & 0xffu);
                        int sc = qk == ti ? P.match : P.mismatch_neg;
                        if ((qk | ti) > 3) sc = P.ambig;
                        M[c] = __viaddmin_s32_relu(Hd, sc, Hd << 16);
                        g[c + 1] = __viaddmax_s32_relu(M[c], -P.oe_ins, g[c] - P.e_ins);
                    }
                    // inclusive max-plus scan of the lanes' outgoing F (decay 4*e_ins per lane)
                    int x = g[4];
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int y = __shfl_up_sync(FULL, x, d);
                        if (lane >= d) x = max(x, y - e_ins4 * d);
                    }
                    int fin = __shfl_up_sync(FULL, x, 1);
                    fin = lane == 0 ? f_carry : max(fin, f_carry - e_ins4 * lane);
                    int H[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int F = max(g[c], fin - c * P.e_ins);
                        H[c] = __vimax3_s32(M[c], E[c], F);
                    }
                    int hprev = __shfl_up_sync(FULL, H[3], 1);
                    if (lane == 0) hprev = h_carry;
                    // new cells: h = H(i, j-1), e = E(i+1, j); cell `end` gets {H(i,end-1), 0}
                    uint32_t nw[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int j = j0 + c;
                        const int hp = c == 0 ? hprev : H[c - 1];
                        const int en = __viaddmax_s32_relu(M[c], -P.oe_del, E[c] - P.e_del);
                        uint32_t v = wds[c];
                        if (j >= beg && j < end) {
                            v = (uint32_t)hp | ((uint32_t)en << 16);
                            mkey = max(mkey, ((unsigned)H[c] << 16) | (unsigned)j);
                            if (j == end - 1) h_end = H[c];
                        } else if (j == end) {
                            v = (uint32_t)hp;
                        }
                        nw[c] = v;
                    }
                    if (j0 <= end) *reinterpret_cast<uint4*>(eh + j0) = make_uint4(nw[0], nw[1], nw[2], nw[3]);
                    f_carry = max(__shfl_sync(FULL, x, 31), f_carry - e_ins4 * 32);
                    h_carry = __shfl_sync(FULL, H[3], 31);
                }
                if (lane == 0) my_cells += end - beg;
            } else if (lane == 0) {
                eh[end] = (uint32_t)h1_init;               // bandedSWA.cpp:213 with an empty window
            }
            __syncwarp();
            mkey = __reduce_max_sync(FULL, mkey);
            const int h1 = end > beg ? __reduce_max_sync(FULL, h_end) : h1_init;
            const int jfin = end > beg ? end : beg;
            if (jfin == qlen) {
                if (!(st.gscore > h1)) st.max_ie = i;
                st.gscore = max(st.gscore, h1);
            }
            const int m = (int)(mkey >> 16), mj = (int)(mkey & 0xffffu);
            if (bsw_row_update(P, st, i, m, mj)) break;
            // next row's window (bandedSWA.cpp:230-233), 32 cells per probe
            {
                int j = beg;
                while (j < end) {
                    const int idx = j + lane;
                    const uint32_t v = idx < end ? eh[idx] : 1u;
                    const unsigned b = __ballot_sync(FULL, v != 0);
                    if (b) { j += __ffs(b) - 1; break; }
                    j += 32;
                }
                beg = j;
                j = end;
                while (j >= beg) {
                    const int idx = j - lane;
                    const uint32_t v = idx >= beg ? eh[idx] : 1u;
                    const unsigned b = __ballot_sync(FULL, v != 0);
                    if (b) { j -= __ffs(b) - 1; break; }
                    j -= 32;
                }
                end = min(j + 2, qlen);
            }
        }
        if (lane == 0) {
            int4 r = bsw_pack_result(st);
            if (!cell_counter) { r.w = (int)(my_cells - pair_cells0); }   // latency route: the pair's effective cells travel with its result
            res[s] = r;
        }
        pair_cells0 = my_cells;
        __syncwarp();
    }
    if (lane == 0 && my_cells && cell_counter) atomicAdd(cell_counter, (unsigned long long)my_cells);
}

// ---------------------------------------------------------------------------------------
// Dependency-free DPX throughput probe: the roofline denominator P_int.
// Each thread keeps 8 independent VIADDMNMX chains; ops = threads * iters * 8.
// ---------------------------------------------------------------------------------------
__global__ void bsw_int_peak_kernel(int* out, int iters, int seed)
{
    int a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    int a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const int b = seed | 1, c = seed - 7;
#pragma unroll 1
    for (int k = 0; k < iters; ++k) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = __viaddmax_s32(a0, b, c); a1 = __viaddmax_s32(a1, b, c);
            a2 = __viaddmax_s32(a2, b, c); a3 = __viaddmax_s32(a3, b, c);
            a4 = __viaddmax_s32(a4, b, c); a5 = __viaddmax_s32(a5, b, c);
            a6 = __viaddmax_s32(a6, b, c); a7 = __viaddmax_s32(a7, b, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}

} // namespace bsw
