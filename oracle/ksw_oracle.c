/*
 * ksw_oracle.c -- CPU restatement of the reference's per-pair extension semantics.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under genomicsbench_b200/ may link, import or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker (or, when oracle/_ref is
 * absent, as the timed CPU "port").
 *
 * What it restates (paths relative to /root/reference):
 *   benchmarks/bsw/bandedSWA.cpp:128-249   BandedPairWiseSW::scalarBandedSWA
 *   tools/bwa/ksw.c:380-479                ksw_extend2 (canonical twin)
 * with the z-drop rule of the vector kernel that the benchmark actually runs
 *   benchmarks/bsw/bandedSWA.cpp:323-336   ZSCORE16 (unconditional, no e_del/e_ins factor)
 * selectable against the scalar rule (bandedSWA.cpp:222-228) via zdrop_mode.
 * Scoring follows bwa_fill_scmat (benchmarks/bsw/main_banded.cpp:73-81) and the
 * vector code's ambiguity handling (bandedSWA.cpp:341-344, :1272).
 *
 * Parity pin: checked against the reference's own AVX2 getScores16 and its
 * scalarBandedSWA, compiled from /root/reference into oracle/_ref/libbswref.so
 * (oracle/Makefile), by tests/test_oracle_vs_reference.py, and against the golden
 * vectors under tests/golden/ that were generated from that library
 * (tests/golden/make_golden.py).  The reference repo holds no bsw test vectors of
 * its own.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int64_t idr, idq, id;
    int32_t len1, len2, h0, seqid, regid;
    int32_t score, tle, gtle, qle, gscore, max_off;
} oracle_seqpair;   /* bandedSWA.h:91-100 */

typedef struct {
    int32_t o_del, e_del, o_ins, e_ins, zdrop, end_bonus;
    int32_t match, mismatch, ambig;
    int32_t zdrop_mode;     /* 0 = vector rule (getScores16), 1 = scalar rule */
} oracle_params;

static inline int sc(const oracle_params *p, int x, int y)
{
    if (x >= 4 || y >= 4) return p->ambig;          /* main_banded.cpp:77-80 */
    return x == y ? p->match : -p->mismatch;        /* main_banded.cpp:76   */
}

/* One pair.  out[6] = score, qle, tle, gtle, gscore, max_off.  Returns the number of
 * inner-loop iterations (the SW_cells++ hook, bandedSWA.cpp:211).  If row_trip != NULL
 * it receives, per target row, the inner-loop trip count (design aid for the tests that
 * model warp divergence); row_trip must hold tlen entries, unused rows are left 0. */
int64_t bsw_oracle_pair(const oracle_params *p, const uint8_t *query, int qlen,
                        const uint8_t *target, int tlen, int w, int h0,
                        int32_t out[6], int32_t *row_trip)
{
    const int o_del = p->o_del, e_del = p->e_del, o_ins = p->o_ins, e_ins = p->e_ins;
    const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    int i, j, beg, end, max, max_i, max_j, max_ie, gscore, max_off, max_ins, max_del, mx;
    int64_t cells = 0;
    int32_t *eh_h = (int32_t *)calloc((size_t)qlen + 1, sizeof(int32_t));
    int32_t *eh_e = (int32_t *)calloc((size_t)qlen + 1, sizeof(int32_t));

    /* first row: bandedSWA.cpp:155-157 */
    eh_h[0] = h0;
    if (qlen >= 1) eh_h[1] = h0 > oe_ins ? h0 - oe_ins : 0;
    for (j = 2; j <= qlen && eh_h[j - 1] > e_ins; ++j)
        eh_h[j] = eh_h[j - 1] - e_ins;

    /* band clamp: bandedSWA.cpp:160-168 (max over the scoring matrix) */
    mx = p->match;
    if (-p->mismatch > mx) mx = -p->mismatch;
    if (p->ambig > mx) mx = p->ambig;
    max_ins = (int)((double)(qlen * mx + p->end_bonus - o_ins) / e_ins + 1.);
    max_ins = max_ins > 1 ? max_ins : 1;
    w = w < max_ins ? w : max_ins;
    max_del = (int)((double)(qlen * mx + p->end_bonus - o_del) / e_del + 1.);
    max_del = max_del > 1 ? max_del : 1;
    w = w < max_del ? w : max_del;

    /* DP loop: bandedSWA.cpp:171-235 */
    max = h0; max_i = max_j = -1; max_ie = -1; gscore = -1; max_off = 0;
    beg = 0; end = qlen;
    for (i = 0; i < tlen; ++i) {
        int f = 0, h1, m = 0, mj = -1;
        const int ti = target[i];
        if (beg < i - w) beg = i - w;
        if (end > i + w + 1) end = i + w + 1;
        if (end > qlen) end = qlen;
        if (beg == 0) {
            h1 = h0 - (o_del + e_del * (i + 1));
            if (h1 < 0) h1 = 0;
        } else h1 = 0;
        for (j = beg; j < end; ++j) {
            int h, M = eh_h[j], e = eh_e[j], t;
            eh_h[j] = h1;
            M = M ? M + sc(p, ti, query[j]) : 0;
            h = M > e ? M : e;
            h = h > f ? h : f;
            h1 = h;
            mj = m > h ? mj : j;
            m = m > h ? m : h;
            t = M - oe_del; t = t > 0 ? t : 0;
            e -= e_del;     e = e > t ? e : t;
            eh_e[j] = e;
            t = M - oe_ins; t = t > 0 ? t : 0;
            f -= e_ins;     f = f > t ? f : t;
        }
        if (end > beg) cells += end - beg;
        if (row_trip) row_trip[i] = end > beg ? end - beg : 0;
        /* j after the loop: end if it ran, beg otherwise (bandedSWA.cpp:213-217) */
        j = end > beg ? end : beg;
        eh_h[end] = h1; eh_e[end] = 0;
        if (j == qlen) {
            max_ie = gscore > h1 ? max_ie : i;
            gscore = gscore > h1 ? gscore : h1;
        }
        if (m == 0) break;
        if (m > max) {
            int d = mj - i; if (d < 0) d = -d;
            max = m; max_i = i; max_j = mj;
            max_off = max_off > d ? max_off : d;
        } else if (p->zdrop_mode == 0) {
            /* vector rule, bandedSWA.cpp:323-336: |di - dj| without the gap-extend factor,
             * evaluated unconditionally */
            int di = i - max_i, dj = mj - max_j;
            int gap = di > dj ? di - dj : dj - di;
            if (max - m - gap > p->zdrop) break;
        } else if (p->zdrop > 0) {
            /* scalar rule, bandedSWA.cpp:222-228 */
            if (i - max_i > mj - max_j) {
                if (max - m - ((i - max_i) - (mj - max_j)) * e_del > p->zdrop) break;
            } else {
                if (max - m - ((mj - max_j) - (i - max_i)) * e_ins > p->zdrop) break;
            }
        }
        /* window for the next row: bandedSWA.cpp:230-233 */
        for (j = beg; j < end && eh_h[j] == 0 && eh_e[j] == 0; ++j);
        beg = j;
        for (j = end; j >= beg && eh_h[j] == 0 && eh_e[j] == 0; --j);
        end = j + 2 < qlen ? j + 2 : qlen;
    }
    free(eh_h); free(eh_e);
    out[0] = max; out[1] = max_j + 1; out[2] = max_i + 1;
    out[3] = max_ie + 1; out[4] = gscore; out[5] = max_off;
    return cells;
}

/* Batch form with the SeqPair boundary of getScores16 (bandedSWA.cpp:1124-1148):
 * results written in place, input order.  Returns total effective cells. */
int64_t bsw_oracle_batch(const oracle_params *p, oracle_seqpair *pairs,
                         const uint8_t *seq_ref, const uint8_t *seq_qer,
                         int64_t n, int w, int nthreads)
{
    int64_t total = 0;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 512) num_threads(nthreads) reduction(+:total)
    for (int64_t k = 0; k < n; ++k) {
        oracle_seqpair *sp = pairs + k;
        int32_t out[6];
        total += bsw_oracle_pair(p, seq_qer + sp->idq, sp->len2, seq_ref + sp->idr, sp->len1,
                                 w, sp->h0, out, NULL);
        sp->score = out[0]; sp->qle = out[1]; sp->tle = out[2];
        sp->gtle = out[3]; sp->gscore = out[4]; sp->max_off = out[5];
    }
    return total;
}

/* Per-row trip counts of one pair (tests use it to model lane divergence). */
int64_t bsw_oracle_row_trips(const oracle_params *p, const uint8_t *query, int qlen,
                             const uint8_t *target, int tlen, int w, int h0, int32_t *row_trip)
{
    int32_t out[6];
    memset(row_trip, 0, sizeof(int32_t) * (size_t)tlen);
    return bsw_oracle_pair(p, query, qlen, target, tlen, w, h0, out, row_trip);
}

int bsw_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
