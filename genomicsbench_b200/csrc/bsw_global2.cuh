// bsw_global2.cuh -- banded global alignment with traceback -> CIGAR, second kernel, sm_100a.
//
// Same contract as bsw_global.cuh (ksw_global2, tools/bwa/ksw.c:502-606; push_cigar :489-500): one alignment
// per thread, the reference's int32 recurrence and its order of comparisons.  What changed is everything around
// the recurrence -- the first kernel issued ~32 instructions per cell at 1.5 warps per scheduler:
//
//  * Rows in BAND coordinates.  Row i only touches columns [i - w, i + w + 1]; a cell (i, j) lives in slot
//    d = j - i + w of the thread's row buffer.  Cell (i, j) reads what row i - 1 left for column j -- slot d + 1 --
//    and writes slot d, so the sweep shifts the row down by one slot in place: no circular index, no wrap
//    test, every address of an 8-cell block is the block's base plus a constant.  2 w + 2 slots per thread
//    (min(2 w + 1, qlen + w) + 1 when the query is shorter than the band).
//  * ROWS16: a slot is one 32-bit word, e << 16 | (h & 0xffff) -- half the shared memory, twice the warps.  The
//    arithmetic stays 32 bits in registers.  Exact while every REAL value fits 16 signed bits (host check,
//    rows16_ok below); the reference's -2^30 only ever occurs as a fresh constant (the E of a column the
//    previous row did not reach, F and H(i, beg - 1) at a row's start -- never as the result of a cell inside
//    the band, because m = H(i-1, j-1) + s is real for every cell of the band), so E = -32768 in a slot and
//    -2^30 in registers lose every comparison exactly as in the reference.  Where a value does not fit, the
//    launch uses 64-bit slots {h, e} (ROWS16 = false).
//  * Scores: the query sits in shared memory at 4 bits per base; a block of eight cells cuts its eight bases out
//    of two words with one funnel shift, and ONE PRMT per four cells turns them into four scores: the row's target
//    base selects an 8-byte table (scores against A C G T N as signed bytes, kernel parameter), the four query
//    nibbles are the PRMT's selector.  H(i-1, j-1) + s is then one IDP.4A against a one-hot byte vector (FMA pipe:
//    it widens the byte and adds).  Needs |score| <= 127 (host check).
//  * Directions: four raw comparison bits per cell (M < H, E < H, E extended, F extended), each the SIGN
//    of a difference, shifted into the block's direction word by one funnel shift (SHF.L.W) -- no predicate, no
//    select; the differences run on the FMA pipe, H = max(M, E, F) is one VIMNMX3, E and F one VIADDMNMX each.
//    Eight cells per 32-bit store: the direction matrix is half the first kernel's size.  The backtrack decodes
//    the bits into ksw.c's which-state machine.
//  * Loads run ahead: a block's eight slots and the next query word are requested while the block before it
//    computes, the target's bases four rows before their first use.
//  * Backtrack: the operation being built stays in registers (the first kernel read-modify-wrote the list in
//    HBM at every step), and the words of the next seven rows are requested together with the current one --
//    the walk moves up one row per step and its column drifts by at most one, so their address is known.
//
// align_one is __host__ __device__: tests/emu/g2_emu.cu runs the very same source on the CPU against the
// goldens of the reference's own ksw_global2 (tests/test_g2_emulation.py).
#pragma once
#include <cstdint>
#include <algorithm>
#include <cstring>
#include <cuda_runtime.h>
#include "bsw_global.cuh"

#ifndef BSW_HD
#define BSW_HD __host__ __device__ __forceinline__
#endif

namespace bsw {
namespace g2 {

constexpr int BLOCK = 64;                        // threads (alignments) per block
constexpr int NEG16 = -32768;                    // a slot's E when the reference holds -2^30 there

struct Params {
    int o_del, e_del, o_ins, e_ins;
    // a row's score table = 8 signed bytes indexed by the query base 0 .. 4 (4 = N): for a target base t < 4 the low
    // word is mm4 ^ mx << 8 t (mismatch everywhere, match at byte t), for N it is amb4; byte 4 is always ambig
    uint32_t mm4, amb4, mx, amb1;
};

// ksw.c:505-506 with bwa_fill_scmat's matrix (benchmarks/bsw/main_banded.cpp:73-81): match / -mismatch, ambig against N
inline void fill_table(Params& P, int match, int mismatch_neg, int ambig)
{
    const uint32_t m = (uint8_t)(int8_t)match, x = (uint8_t)(int8_t)mismatch_neg, a = (uint8_t)(int8_t)ambig;
    P.mm4 = x * 0x01010101u; P.amb4 = a * 0x01010101u; P.mx = m ^ x; P.amb1 = a;
}
BSW_HD uint32_t row_table_lo(const Params& P, int tb) { return tb < 4 ? P.mm4 ^ (P.mx << (8 * tb)) : P.amb4; }      // (the high word is amb1)
inline bool scores_ok(int match, int mismatch_neg, int ambig)
{
    auto ok = [](int v) { return v >= -127 && v <= 127; };
    return ok(match) && ok(mismatch_neg) && ok(ambig);
}
// every real H, E, F and every intermediate of a cell within 16 signed bits?  H(i, j) of a cell of the band is at
// least the score of the path "diagonal, then one gap of |i - j| <= w" and at most match * min(i, j); the first row
// and column hold -(o + e * k), k <= w + 1; E and F of the band are at most one gap open below an H, and a cell
// subtracts one more gap open / adds one score before it compares.
inline bool rows16_ok(const Params& P, int match, int mismatch_neg, int ambig, int qlen, int tlen, int w)
{
    const long long worst = std::max(0, -std::min(std::min(match, mismatch_neg), ambig));
    const long long best = std::max(0, std::max(std::max(match, mismatch_neg), ambig));
    const long long len = std::max(qlen, tlen) + 1;
    const long long gap = std::max(P.o_del + P.e_del, P.o_ins + P.e_ins), ext = std::max(P.e_del, P.e_ins);
    const long long lo = worst * len + 3 * gap + ext * ((long long)w + 2);
    const long long hi = best * len + gap;
    return lo <= 32000 && hi <= 32000;
}

// slots a thread's row needs, words per row of its direction matrix
BSW_HD int row_slots(int qlen, int w) { const int a = 2 * w + 1, b = qlen + w; return (a < b ? a : b) + 1; }
BSW_HD int z_pitch(int qlen, int w) { const int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1; return (n_col + 7) >> 3; }

inline thread_local long long emu_wraps = 0;     // host emulation only (tests/emu): values that left 16 bits in a 16-bit slot

// ---- the instructions the cell leans on (device: one instruction each; host: what they compute) -----------
// generic byte permute of {hi, lo}: nibble n of sel picks the byte of result byte n (bit 3: its sign, replicated)
BSW_HD uint32_t prmt(uint32_t lo, uint32_t hi, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
    uint32_t v;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(v) : "r"(lo), "r"(hi), "r"(sel));
    return v;
#else
    const uint64_t src = (uint64_t)hi << 32 | lo;
    uint32_t v = 0;
    for (int n = 0; n < 4; ++n) {
        const uint32_t s = sel >> (4 * n) & 0xf;
        uint32_t b = (uint32_t)(src >> (8 * (s & 7))) & 0xff;
        if (s & 8) b = (b & 0x80) ? 0xff : 0;
        v |= b << (8 * n);
    }
    return v;
#endif
}
// c + (signed byte k of sw), K a compile-time constant: IDP.4A against a one-hot byte vector (FMA pipe)
template <int K> BSW_HD int add_byte(uint32_t sw, int c)
{
#if defined(__CUDA_ARCH__)
    return __dp4a((int)sw, (int)(1u << (8 * K)), c);
#else
    return c + (int)(int8_t)(sw >> (8 * K) & 0xff);
#endif
}
// max(a, b, c): VIMNMX3
BSW_HD int max3(int a, int b, int c)
{
#if defined(__CUDA_ARCH__)
    return __vimax3_s32(a, b, c);
#else
    return std::max(a, std::max(b, c));
#endif
}
// max(a + b, c): VIADDMNMX
BSW_HD int addmax(int a, int b, int c)
{
#if defined(__CUDA_ARCH__)
    return __viaddmax_s32(a, b, c);
#else
    return a + b > c ? a + b : c;
#endif
}
// acc << 1 | (d < 0): the sign of a difference enters the direction word with one funnel shift (SHF.L.W)
BSW_HD uint32_t push_sign(uint32_t acc, int d)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_l((uint32_t)d, acc, 1);
#else
    return acc << 1 | ((uint32_t)d >> 31);
#endif
}
// low 32 bits of {hi, lo} >> sh, sh < 32
BSW_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return (uint32_t)(((uint64_t)hi << 32 | lo) >> (sh & 31));
#endif
}

// ---- row slots --------------------------------------------------------------------------------------------
// R16: word e << 16 | h;  else two words {h, e}.  `rows` points at the thread's slot 0; consecutive slots are
// STRIDE words apart (the block's threads interleaved: bank = lane)
// Accesses are volatile: they stay in source order, so a block's loads are issued a whole block ahead of their use
// (the compiler otherwise sinks each one next to its consumer and the sweep waits for shared memory: 25 % of the
// stall samples of the first build, profiles/r04d_ncu_global2_w20.txt).
template <bool R16> struct Slots;
template <> struct Slots<true> {
    static constexpr int WORDS = 1;
    typedef uint32_t Raw;
    static BSW_HD Raw load_raw(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
    // widening the two halves = two IDP.2A against unit byte vectors: FMA pipe, the ALU pipe is the kernel's bound
    static BSW_HD void unpack(Raw c, int& h, int& e)
    {
#if defined(__CUDA_ARCH__)
        h = __dp2a_lo((int)c, 0x0001, 0); e = __dp2a_lo((int)c, 0x0100, 0);
#else
        h = (int)(int16_t)(c & 0xffffu); e = (int)(int16_t)(c >> 16);
#endif
    }
    static BSW_HD void store(uint32_t* p, int h, int e)
    {
#if !defined(__CUDA_ARCH__)
        // (emulation) a value that does not fit: only H(i, beg - 1) = -2^30 of a row that starts inside the query may, its slot is never read
        if (e < -32768 || e > 32767 || ((h < -32768 || h > 32767) && h != G_MINUS_INF)) ++emu_wraps;
#endif
        *reinterpret_cast<volatile uint32_t*>(p) = prmt((uint32_t)h, (uint32_t)e, 0x5410u);
    }
    static BSW_HD int neg() { return NEG16; }
};
template <> struct Slots<false> {
    static constexpr int WORDS = 2;
    typedef uint2 Raw;
    static BSW_HD Raw load_raw(const uint32_t* p)
    {
#if defined(__CUDA_ARCH__)
        uint2 v;                                 // one LDS.64 that keeps its place (a volatile uint2 access would split in two)
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
        return v;
#else
        return *reinterpret_cast<const uint2*>(p);
#endif
    }
    static BSW_HD void unpack(Raw c, int& h, int& e) { h = (int)c.x; e = (int)c.y; }
    static BSW_HD void store(uint32_t* p, int h, int e)
    {
#if defined(__CUDA_ARCH__)
        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"((uint32_t)__cvta_generic_to_shared(p)), "r"(h), "r"(e) : "memory");
#else
        *reinterpret_cast<uint2*>(p) = make_uint2((uint32_t)h, (uint32_t)e);
#endif
    }
    static BSW_HD int neg() { return G_MINUS_INF; }
};
template <bool R16> BSW_HD void slot_load(const uint32_t* p, int& h, int& e) { Slots<R16>::unpack(Slots<R16>::load_raw(p), h, e); }

// One cell (ksw.c:544-566).  In: m = H(i-1, j-1) + s, E(i, j) in e_in, the running f and h1 = H(i, j-1).
// Out: H(i, j) in h1 (the old h1 and E(i+1, j) go to the slot: out_h, out_e), f = F(i, j+1), and four bits
// shifted into acc, each the sign of a difference (no difference can wrap: |values| <= 2^30 + lengths * scores).
// H = max(M, E, F) is one VIMNMX3; which of the three it came from (ksw.c:551-553: M on M >= E and M >= F, else E
// on E >= F, else F) follows from two signs:
//   M < H              <=>  m - H < 0      (not set: M)
//   E < H              <=>  e - H < 0      (M < H and not set: E; both set: F)
//   E extended         <=>  (m - oe_del) - E' < 0     (:558: e > t, E' = max(e - e_del, t))
//   F extended         <=>  (m - oe_ins) - F' < 0     (:563)
#define BSW_G2_CELL(m_in, e_in, out_h, out_e, acc)                                      \
    {                                                                                   \
        const int m_ = (m_in), e_ = (e_in);                                             \
        const int hn_ = max3(m_, e_, f);                                                \
        acc = push_sign(acc, m_ - hn_);                                                 \
        acc = push_sign(acc, e_ - hn_);                                                 \
        out_h = h1;                                                                     \
        h1 = hn_;                                                                       \
        const int t1_ = m_ - oe_del, t2_ = m_ - oe_ins;                                 \
        out_e = addmax(e_, -P.e_del, t1_);                                              \
        f = addmax(f, -P.e_ins, t2_);                                                   \
        acc = push_sign(acc, t1_ - out_e);                                              \
        acc = push_sign(acc, t2_ - f);                                                  \
    }

// what the backtrack reads: the four bits of cell c (0 .. 7) of a direction word -- cell 0 sits in the top nibble,
// the first comparison in the nibble's top bit
BSW_HD uint32_t cell_bits(uint32_t word, int c) { return word >> (28 - 4 * c) & 0xfu; }

// q: the thread's query, 4 bits per base (codes 0 .. 4), eight bases per word, two words of padding behind;
// consecutive words QSTRIDE words apart.  r: target bases, one per byte (global, 4-aligned, padded).
// rows: slot 0 of the thread's row; z: the alignment's direction words; cg: its operation list (qlen + tlen words)
template <bool R16, int STRIDE, int QSTRIDE>
BSW_HD void align_one(const Params& P, int qlen, int tlen, int w, const uint32_t* q, const uint8_t* r, uint32_t* rows,
                      uint32_t* z, uint32_t* cg, int32_t& score_out, int32_t& ncig_out)
{
    typedef Slots<R16> S;
    constexpr int SW = STRIDE * S::WORDS;                    // words between consecutive slots
    const int oe_del = P.o_del + P.e_del, oe_ins = P.o_ins + P.e_ins;
    const int pitch = z_pitch(qlen, w);
    // first row (ksw.c:521-525): row 0 reads column j from slot j + w + 1; only columns 0 .. w are ever read
    {
        uint32_t* p = rows + (size_t)(w + 1) * SW;
        S::store(p, 0, S::neg());
        for (int j = 1; j < qlen && j <= w; ++j) S::store(p + (size_t)j * SW, -(P.o_ins + P.e_ins * j), S::neg());   // (column min(w + 1, qlen) is written by row 0 itself)
    }
    // byte 4 of every row's score table (N in the query): the same for all rows; kept in a register the optimiser cannot
    // re-derive from the parameter bank (it re-loaded it with an LDC in every block, and the PRMT behind it waited)
    uint32_t thi = P.amb1;
#if defined(__CUDA_ARCH__)
    asm volatile("mov.u32 %0, %0;" : "+r"(thi));
#endif
    int last_h1 = 0;
    // four target bases per load, requested four rows before their first use (r is 4-aligned, readable to its span)
    uint32_t tw = 0, tw_next = *reinterpret_cast<const uint32_t*>(r);
    for (int i = 0; i < tlen; ++i) {                                            // ksw.c:527-589
        const int beg = i > w ? i - w : 0;
        const int end = i + w + 1 < qlen ? i + w + 1 : qlen;
        const int ncell = end - beg;
        int f = G_MINUS_INF;
        int h1 = beg == 0 ? -(P.o_del + P.e_del * (i + 1)) : G_MINUS_INF;
        if ((i & 3) == 0) {
            tw = tw_next;
            if (i + 4 < tlen) tw_next = *reinterpret_cast<const uint32_t*>(r + i + 4);
        }
        int tb = (int)(tw >> (8 * (i & 3)) & 0xff);
        tb = tb > 4 ? 4 : tb;
        const uint32_t tlo = row_table_lo(P, tb);
        uint32_t* rp = rows + (size_t)(w > i ? w - i : 0) * SW;                // slot of the row's first cell
        uint32_t* zi = z + (size_t)i * pitch;
        // the row's query bases: eight per block, cut out of two consecutive words at the row's bit offset
        // (two words are in registers and the third is requested while a block computes: the word behind the query's
        // last two are padding)
        const uint32_t* qw = q + (size_t)(beg >> 3) * QSTRIDE;
        const uint32_t qsh = (uint32_t)(beg & 7) * 4u;
        uint32_t qlo = qw[0], qhi = qw[QSTRIDE];
        int c = 0;
        // a block's eight slots are loaded while the block before it computes (slots c + 9 .. c + 16 are not written before
        // block c + 8 runs: block c stores slots c .. c + 7)
        typename S::Raw cur[8];
        if (ncell >= 8) {
#pragma unroll
            for (int k = 0; k < 8; ++k) cur[k] = S::load_raw(rp + (size_t)(1 + k) * SW);
        }
        for (; c + 8 <= ncell; c += 8) {
            typename S::Raw nxt[8];
            const bool more = c + 16 <= ncell;
            if (more) {
#pragma unroll
                for (int k = 0; k < 8; ++k) nxt[k] = S::load_raw(rp + (size_t)(c + 9 + k) * SW);
            }
            const uint32_t qnext = *reinterpret_cast<const volatile uint32_t*>(qw + 2 * QSTRIDE);
            qw += QSTRIDE;
            const uint32_t q8 = funnel_r(qlo, qhi, qsh);
            qlo = qhi; qhi = qnext;
            const uint32_t s03 = prmt(tlo, thi, q8), s47 = prmt(tlo, thi, q8 >> 16);   // the eight scores, one signed byte each
            int hh[8], ee[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) S::unpack(cur[k], hh[k], ee[k]);
            // (two direction half-words per block: 16 dependent shifts each instead of 32 in a row)
            uint32_t acc = 0, acc2 = 0;
            int oh, oe;
#define BSW_G2_STEP(K, SWORD, ACC)                                                      \
            BSW_G2_CELL(add_byte<(K) & 3>(SWORD, hh[K]), ee[K], oh, oe, ACC)            \
            S::store(rp + (size_t)(c + (K)) * SW, oh, oe);
            BSW_G2_STEP(0, s03, acc) BSW_G2_STEP(1, s03, acc) BSW_G2_STEP(2, s03, acc) BSW_G2_STEP(3, s03, acc)
            BSW_G2_STEP(4, s47, acc2) BSW_G2_STEP(5, s47, acc2) BSW_G2_STEP(6, s47, acc2) BSW_G2_STEP(7, s47, acc2)
            zi[c >> 3] = prmt(acc2, acc, 0x5410u);                               // acc << 16 | acc2
            if (more) {
#pragma unroll
                for (int k = 0; k < 8; ++k) cur[k] = nxt[k];
            }
        }
        if (c < ncell) {
            const uint32_t q8 = funnel_r(qlo, qhi, qsh);
            const uint32_t s03 = prmt(tlo, thi, q8), s47 = prmt(tlo, thi, q8 >> 16);
            uint32_t acc = 0;
            const int c0 = c;
            int k = 0;
            if (c + 4 <= ncell) {                                               // four cells at once, then at most three one by one
                int hh[4], ee[4], oh, oe;
#pragma unroll
                for (int u = 0; u < 4; ++u) slot_load<R16>(rp + (size_t)(c + 1 + u) * SW, hh[u], ee[u]);
                BSW_G2_STEP(0, s03, acc) BSW_G2_STEP(1, s03, acc) BSW_G2_STEP(2, s03, acc) BSW_G2_STEP(3, s03, acc)
                c += 4; k = 4;
            }
#undef BSW_G2_STEP
            for (; c < ncell; ++c, ++k) {
                int hd, e, oh, oe;
                slot_load<R16>(rp + (size_t)(c + 1) * SW, hd, e);
                const int s = (int)(int8_t)((k < 4 ? s03 : s47) >> (8 * (k & 3)) & 0xff);
                BSW_G2_CELL(hd + s, e, oh, oe, acc)
                S::store(rp + (size_t)c * SW, oh, oe);
            }
            zi[c0 >> 3] = acc << (4 * (8 - (ncell - c0)));
        }
        S::store(rp + (size_t)ncell * SW, h1, S::neg());                        // eh[end] (ksw.c:588)
        last_h1 = h1;
    }
    score_out = last_h1;        // eh[qlen].h (ksw.c:590): the last row ends at qlen inside the supported domain (qlen <= tlen + w)

    // backtrack (ksw.c:591-603).  State `which`: 0 = H, 1 = E (deletion), 2 = F (insertion); a cell's four bits
    // decode to the reference's next state: from H 0 unless M < H, then 1 unless E < H, else 2; from E 1 if extended; from F 2 if extended.
    int n_op = 0, which = 0, cur_op = -1, cur_len = 0;
    auto push = [&](int op, int len) {                                          // push_cigar, ksw.c:489-500
        if (op == cur_op) cur_len += len;
        else { if (cur_op >= 0) cg[n_op++] = (uint32_t)cur_len << 4 | (uint32_t)cur_op; cur_op = op; cur_len = len; }
    };
    int i = tlen - 1, k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
    while (i >= 0 && k >= 0) {
        const int wi = (k - (i > w ? i - w : 0)) >> 3;                          // the word the walk is in; rows i .. i-7 of it
        uint32_t b[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) b[t] = z[(size_t)(i - t > 0 ? i - t : 0) * pitch + wi];
        bool reload = false;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            if (reload) break;
            for (;;) {                                                          // steps inside row i (insertions stay in it)
                const int c = k - (i > w ? i - w : 0);
                if ((c >> 3) != wi) { reload = true; break; }
                const uint32_t nib = cell_bits(b[t], c & 7);
                which = which == 0 ? ((nib & 8u) ? ((nib & 4u) ? 2 : 1) : 0) : which == 1 ? (int)(nib >> 1 & 1u) : (int)(nib << 1 & 2u);
                if (which == 2) { push(1, 1); --k; if (k < 0) { reload = true; break; } }
                else { if (which == 0) { push(0, 1); --k; } else push(2, 1); --i; break; }
            }
            if (i < 0 || k < 0) reload = true;
        }
    }
    if (i >= 0) push(2, i + 1);
    if (k >= 0) push(1, k + 1);
    if (cur_op >= 0) cg[n_op++] = (uint32_t)cur_len << 4 | (uint32_t)cur_op;
    for (int a = 0; a < n_op >> 1; ++a) { const uint32_t tmp = cg[a]; cg[a] = cg[n_op - 1 - a]; cg[n_op - 1 - a] = tmp; }
    ncig_out = n_op;
}
#undef BSW_G2_CELL

// eight bases (one per byte, two words) -> one word, 4 bits each, anything above 4 clamped to 4 (N)
BSW_HD uint32_t pack8(uint32_t b03, uint32_t b47)
{
    auto squeeze = [](uint32_t v) {                     // bytes a b c d (values < 16) -> 16 bits d c b a
        v = (v | v >> 4) & 0x00ff00ffu;
        return (v | v >> 8) & 0xffffu;
    };
#if defined(__CUDA_ARCH__)
    b03 = __vminu4(b03, 0x04040404u); b47 = __vminu4(b47, 0x04040404u);
#else
    auto clamp4 = [](uint32_t v) {                      // per byte min(v, 4)
        uint32_t o = 0;
        for (int k = 0; k < 4; ++k) { const uint32_t x = v >> (8 * k) & 0xff; o |= (x > 4 ? 4u : x) << (8 * k); }
        return o;
    };
    b03 = clamp4(b03); b47 = clamp4(b47);
#endif
    return squeeze(b03) | squeeze(b47) << 16;
}
// the thread's packed query: words 0 .. ceil(qlen / 8) + 1 (the last two are the padding the funnel shifts' look-ahead
// reads); src = the query's bytes (4-aligned, readable up to the next multiple of 8 -- the gather pads)
template <int QSTRIDE>
BSW_HD void pack_query(const uint8_t* src, int qlen, uint32_t* q)
{
    const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
    const int nw = (qlen + 7) >> 3;
    for (int k = 0; k < nw; ++k) q[(size_t)k * QSTRIDE] = pack8(s[2 * k], s[2 * k + 1]);
    q[(size_t)nw * QSTRIDE] = 0;
    q[(size_t)(nw + 1) * QSTRIDE] = 0;
}

// 32-bit words of a thread's packed query (with the two padding words); bytes of dynamic shared memory of a launch
// whose threads need `slots` row slots and `qwords` query words each
BSW_HD int query_words(int qlen) { return ((qlen + 7) >> 3) + 2; }
inline size_t smem_bytes(bool r16, int slots, int qwords)
{
    return ((size_t)slots * (r16 ? 4 : 8) + (size_t)qwords * 4) * BLOCK;
}

#if defined(__CUDACC__)
// desc[0 .. n): the launch's alignments (one class: all need at most `slots` row slots and `qwords` query words)
template <bool R16>
__global__ void __launch_bounds__(BLOCK)
bsw_global2_kernel(const GlobalDesc* __restrict__ desc, int n, const uint8_t* __restrict__ qraw, const uint8_t* __restrict__ rraw,
                   int slots, uint8_t* __restrict__ z, uint32_t* __restrict__ cigar, int32_t* __restrict__ score,
                   int32_t* __restrict__ n_cigar, const Params P)
{
    extern __shared__ __align__(16) uint32_t g2_smem[];
    const int t = blockIdx.x * BLOCK + threadIdx.x;
    if (t >= n) return;
    const GlobalDesc d = desc[t];
    uint32_t* rows = g2_smem + threadIdx.x * Slots<R16>::WORDS;
    uint32_t* qs = g2_smem + (size_t)slots * BLOCK * Slots<R16>::WORDS + threadIdx.x;   // word k of the query at qs[k * BLOCK]
    pack_query<BLOCK>(qraw + d.qoff, d.qlen, qs);
    int32_t sc, nc;
    align_one<R16, BLOCK, BLOCK>(P, d.qlen, d.tlen, d.w, qs, rraw + d.roff, rows, reinterpret_cast<uint32_t*>(z + d.zoff),
                                 cigar + d.coff, sc, nc);
    score[d.idx] = sc;
    n_cigar[d.idx] = nc;
}
#endif

} // namespace g2
} // namespace bsw
