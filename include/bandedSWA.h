/*
 * bandedSWA.h -- header-compatible stand-in for the reference's
 * benchmarks/bsw/bandedSWA.h:114-342 (class BandedPairWiseSW), backed by the B200 engine
 * (include/bsw.h).  The reference driver benchmarks/bsw/main_banded.cpp compiles UNCHANGED
 * against this header:
 *
 *   g++ -O2 -fopenmp -mavx2 -include include/bandedSWA.h \
 *       /root/reference/benchmarks/bsw/main_banded.cpp -Lgenomicsbench_b200/lib -lbsw_b200 -lz
 *
 * (-include makes this file win over the reference header of the same name, which then
 * skips itself through the shared include guard.)  See INTEGRATION.md.
 *
 * Same constructor and method signatures, same SeqPair layout, same in-place result
 * contract.  Differences a maintainer should know:
 *   - pads beyond numPairs are never written and pair.id is never read
 *     (reference: bandedSWA.cpp:1171-1177, :414);
 *   - getScores8 and getScores16 route to the same engine (kernel choice is by query
 *     length, not by score width) and return identical results inside the 8-bit envelope;
 *   - BSW_SHIM_COALESCE=<pairs> in the environment coalesces calls of at most that many pairs from all instances
 *     and threads into shared batches (bsw_extend_async; off by default, see csrc/bsw_shim.cpp);
 *   - there is no CPU path: every method runs on the GPU and aborts, like the reference's
 *     exit(EXIT_FAILURE) (bandedSWA.cpp:94-99), if the device or the library fails.
 */
#ifndef SCALAR_BANDEDSWA_HPP
#define SCALAR_BANDEDSWA_HPP

#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <assert.h>
#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>   /* _mm_malloc / _mm_free / __rdtsc used by main_banded.cpp */
#include <x86intrin.h>
#endif

#include "bsw.h"         /* SeqPair (72 B, bandedSWA.h:91-100) and the C ABI */

/* constants main_banded.cpp and callers rely on (bandedSWA.h:48-87) */
#define MAX_SEQ_LEN_REF 256
#define MAX_SEQ_LEN_QER 128
#define MAX_SEQ_LEN_EXT 256
#define MAX_NUM_PAIRS 10000000
#define MAX_NUM_PAIRS_ALLOC 20000
#define DEFAULT_AMBIG -1
#define SIMD_WIDTH8 32      /* only used by callers to round buffer capacities */
#define SIMD_WIDTH16 16
#define MAX_LINE_LEN 256
#define MAX_SEQ_LEN8 128
#define MAX_SEQ_LEN16 32768
#define SORT_BLOCK_SIZE 16384
#ifndef min_
#define min_(x, y) ((x)>(y)?(y):(x))
#define max_(x, y) ((x)>(y)?(x):(y))
#endif

/* OutScore (bandedSWA.h:103-107) comes with bsw.h: bsw_extend_packed returns it */

typedef struct {
    int32_t h, e;
} eh_t;

class BandedPairWiseSW {
public:
    uint64_t SW_cells;   /* effective DP cells of all calls so far (reference: only with its PROFILE hook) */

    BandedPairWiseSW(const int o_del, const int e_del, const int o_ins,
                     const int e_ins, const int zdrop,
                     const int end_bonus, const int8_t *mat_,
                     const int8_t w_match, const int8_t w_mismatch, int numThreads);
    ~BandedPairWiseSW();

    /* ksw_extend2 semantics for one pair (bandedSWA.cpp:128-249); runs on the GPU */
    int scalarBandedSWA(int qlen, const uint8_t *query, int tlen,
                        const uint8_t *target, int32_t w,
                        int h0, int *_qle, int *_tle,
                        int *_gtle, int *_gscore,
                        int *_max_off);

    /* bandedSWA.cpp:254-272 (scalar z-drop rule) */
    void scalarBandedSWAWrapper(SeqPair *seqPairArray,
                                uint8_t *seqBufRef,
                                uint8_t *seqBufQer,
                                int numPairs,
                                int nthreads,
                                int32_t w);

    /* bandedSWA.cpp:424-446 */
    void getScores8(SeqPair *pairArray,
                    uint8_t *seqBufRef,
                    uint8_t *seqBufQer,
                    int32_t numPairs,
                    uint16_t numThreads,
                    int32_t w);

    /* bandedSWA.cpp:1124-1148 */
    void getScores16(SeqPair *pairArray,
                     uint8_t *seqBufRef,
                     uint8_t *seqBufQer,
                     int32_t numPairs,
                     uint16_t numThreads,
                     int32_t w);

    /* bandedSWA.cpp:108-122: total ticks spent inside the calls above */
    int64_t getTicks();

    /* extra: statistics of the last call (stage timings, effective cells) */
    const bsw_stats* lastStats() const { return &stats_; }

private:
    BandedPairWiseSW(const BandedPairWiseSW&);
    BandedPairWiseSW& operator=(const BandedPairWiseSW&);
    bsw_engine* engine(int zdrop_mode);
    bsw_params checked_params(int zdrop_mode);
    void run(int zdrop_mode, SeqPair*, const uint8_t*, const uint8_t*, int64_t, int32_t);

    bsw_params params_;
    const int8_t *mat;
    bsw_engine *vec_, *scalar_;
    bsw_stats stats_;
    int64_t ticks_;
};

#endif
