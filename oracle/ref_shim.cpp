/*
 * ref_shim.cpp -- thin C exports around the UNMODIFIED reference class, so that the
 * tests and bench.py can execute the reference's own code.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/ksw_oracle.c).  It is compiled together with
 * /root/reference/benchmarks/bsw/bandedSWA.cpp *where that file lies* (oracle/Makefile,
 * -I/root/reference/benchmarks/bsw) into oracle/_ref/libbswref.so; no reference source
 * is copied into this repository.
 *
 * Harness pitfalls honoured here:
 *  - `prof` is extern in bandedSWA.cpp:44 and defined only in main_banded.cpp:71;
 *  - getScores16 writes pad entries up to roundup16(n) (bandedSWA.cpp:1171-1177) and
 *    prefetch-reads two entries beyond (:1261), so calls run on a private, padded copy;
 *  - pair.id must equal the index relative to the pointer passed (sortPairsId, :414).
 */
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <chrono>
#include <omp.h>
#include "bandedSWA.h"

uint64_t prof[10][112];

namespace {

struct RefParams {
    int32_t o_del, e_del, o_ins, e_ins, zdrop, end_bonus, match, mismatch, ambig, zdrop_mode;
};

void fill_scmat(int a, int b, int ambig, int8_t mat[25])
{
    /* same table as bwa_fill_scmat, main_banded.cpp:73-81 */
    int k = 0;
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 4; ++j) mat[k++] = (int8_t)(i == j ? a : -b);
        mat[k++] = (int8_t)ambig;
    }
    for (int j = 0; j < 5; ++j) mat[k++] = (int8_t)ambig;
}

BandedPairWiseSW* make_engine(const RefParams* p, int8_t mat[25])
{
    fill_scmat(p->match, p->mismatch, p->ambig, mat);
    return new BandedPairWiseSW(p->o_del, p->e_del, p->o_ins, p->e_ins, p->zdrop, p->end_bonus,
                                mat, (int8_t)p->match, (int8_t)p->mismatch, 1);
}

inline size_t round16(size_t n) { return (n + SIMD_WIDTH16 - 1) / SIMD_WIDTH16 * SIMD_WIDTH16; }

} // namespace

extern "C" {

int ref_simd_width16(void) { return SIMD_WIDTH16; }
int ref_sizeof_seqpair(void) { return (int)sizeof(SeqPair); }
int ref_max_threads(void) { return omp_get_max_threads(); }

/* getScores16 over n pairs in batches of `batch` (main_banded.cpp:279-291 with
 * schedule(dynamic,1), one engine per thread).  Results land in pairs[] in input order.
 * Returns seconds spent in the getScores16 region only (main_banded.cpp:272-296). */
double ref_getscores16(const RefParams* p, SeqPair* pairs, const uint8_t* seq_ref,
                       const uint8_t* seq_qer, int64_t n, int32_t w, int32_t batch, int32_t nthreads)
{
    if (batch <= 0) batch = 512;
    if (nthreads <= 0) nthreads = 1;
    const int64_t nb = (n + batch - 1) / batch;
    std::vector<BandedPairWiseSW*> eng(nthreads);
    std::vector<std::vector<int8_t>> mats(nthreads, std::vector<int8_t>(25));
    for (int t = 0; t < nthreads; ++t) eng[t] = make_engine(p, mats[t].data());
    /* private padded work buffers, one per batch slot processed by a thread */
    std::vector<SeqPair*> work(nthreads);
    const size_t cap = round16((size_t)batch) + 2;
    for (int t = 0; t < nthreads; ++t) work[t] = (SeqPair*)_mm_malloc(cap * sizeof(SeqPair), 64);

    double seconds = 0.0;
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel num_threads(nthreads)
    {
        const int tid = omp_get_thread_num();
        SeqPair* wk = work[tid];
#pragma omp for schedule(dynamic, 1)
        for (int64_t b = 0; b < nb; ++b) {
            const int64_t first = b * batch;
            const int32_t cnt = (int32_t)((n - first) < batch ? (n - first) : batch);
            memset(wk, 0, cap * sizeof(SeqPair));
            memcpy(wk, pairs + first, (size_t)cnt * sizeof(SeqPair));
            for (int32_t k = 0; k < cnt; ++k) wk[k].id = k;           /* Appendix D.2 */
            eng[tid]->getScores16(wk, const_cast<uint8_t*>(seq_ref), const_cast<uint8_t*>(seq_qer),
                                  cnt, 1, w);
            for (int32_t k = 0; k < cnt; ++k) {
                SeqPair* d = pairs + first + k;
                d->score = wk[k].score; d->tle = wk[k].tle; d->gtle = wk[k].gtle;
                d->qle = wk[k].qle; d->gscore = wk[k].gscore; d->max_off = wk[k].max_off;
            }
        }
    }
    seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (int t = 0; t < nthreads; ++t) { delete eng[t]; _mm_free(work[t]); }
    return seconds;
}

/* In-place timing twin of the stock driver: the caller's array is already padded
 * (capacity roundup16(n)+2, ids batch-relative are set here) and is sorted/unsorted in
 * place exactly as main_banded.cpp:279-291 does.  Used only for the CPU baseline. */
double ref_time_getscores16_inplace(const RefParams* p, SeqPair* pairs, const uint8_t* seq_ref,
                                    const uint8_t* seq_qer, int64_t n, int32_t w, int32_t batch,
                                    int32_t nthreads)
{
    if (batch <= 0) batch = 512;
    if (nthreads <= 0) nthreads = 1;
    std::vector<BandedPairWiseSW*> eng(nthreads);
    std::vector<std::vector<int8_t>> mats(nthreads, std::vector<int8_t>(25));
    for (int t = 0; t < nthreads; ++t) eng[t] = make_engine(p, mats[t].data());
    for (int64_t k = 0; k < n; ++k) pairs[k].id = k % batch;          /* main_banded.cpp:160 */
    const int64_t roundn = (int64_t)round16((size_t)n);
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel num_threads(nthreads)
    {
        const int tid = omp_get_thread_num();
#pragma omp for schedule(dynamic, 1)
        for (int64_t i = 0; i < roundn; i += batch) {
            int32_t cnt = (int32_t)((n - i) >= batch ? batch : n - i);
            if (cnt > 0)
                eng[tid]->getScores16(pairs + i, const_cast<uint8_t*>(seq_ref),
                                      const_cast<uint8_t*>(seq_qer), cnt, 1, w);
        }
    }
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (int t = 0; t < nthreads; ++t) delete eng[t];
    return s;
}

/* One pair alone in a SIMD group (1 real + 15 pads): removes the lane-interaction
 * artifact Q4. */
void ref_getscores16_solo(const RefParams* p, SeqPair* pair, const uint8_t* seq_ref,
                          const uint8_t* seq_qer, int32_t w)
{
    int8_t mat[25];
    BandedPairWiseSW* e = make_engine(p, mat);
    const size_t cap = round16(1) + 2;
    SeqPair* wk = (SeqPair*)_mm_malloc(cap * sizeof(SeqPair), 64);
    memset(wk, 0, cap * sizeof(SeqPair));
    wk[0] = *pair; wk[0].id = 0;
    e->getScores16(wk, const_cast<uint8_t*>(seq_ref), const_cast<uint8_t*>(seq_qer), 1, 1, w);
    /* after the length sort the pads (len1 = 0) come first; find the real pair by id */
    for (size_t k = 0; k < round16(1); ++k)
        if (wk[k].id == 0 && wk[k].len1 == pair->len1 && wk[k].len2 == pair->len2) {
            pair->score = wk[k].score; pair->tle = wk[k].tle; pair->gtle = wk[k].gtle;
            pair->qle = wk[k].qle; pair->gscore = wk[k].gscore; pair->max_off = wk[k].max_off;
            break;
        }
    _mm_free(wk);
    delete e;
}

/* scalarBandedSWAWrapper (bandedSWA.cpp:254-272), single thread. Returns seconds. */
double ref_scalar(const RefParams* p, SeqPair* pairs, const uint8_t* seq_ref,
                  const uint8_t* seq_qer, int64_t n, int32_t w)
{
    int8_t mat[25];
    BandedPairWiseSW* e = make_engine(p, mat);
    auto t0 = std::chrono::steady_clock::now();
    const int64_t step = 1 << 20;
    for (int64_t i = 0; i < n; i += step) {
        int cnt = (int)((n - i) < step ? (n - i) : step);
        e->scalarBandedSWAWrapper(pairs + i, const_cast<uint8_t*>(seq_ref),
                                  const_cast<uint8_t*>(seq_qer), cnt, 1, w);
    }
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    delete e;
    return s;
}

} // extern "C"
