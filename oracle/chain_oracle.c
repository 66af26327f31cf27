/*
 * chain_oracle.c -- CPU restatement of the aligner's seed -> pair construction around the
 * extension kernel.
 *
 * TEST INFRASTRUCTURE ONLY (same rule as ksw_oracle.c: nothing under genomicsbench_b200/ may
 * link, import or call this file).
 *
 * What it restates (paths relative to /root/reference/tools/bwa):
 *   bwamem.c:620-628   cal_max_gap
 *   bwamem.c:643-659   the reference window [rmax0, rmax1) of a chain, incl. the strand-boundary rule
 *   bwamem.c:661-665   seed order: by (score, index), descending
 *   bwamem.c:667-700   containment test against the alignments already made from this chain
 *   bwamem.c:709-763   left extension: reversed query prefix and reference flank, h0 = len * a,
 *                      MAX_BAND_TRY band doubling, local vs to-end decision against pen_clip5
 *   bwamem.c:765-810   right extension: h0 = the left score, decision against pen_clip3
 *   bwamem.c:812-822   seedcov, w, seedlen0
 * The extension itself is ksw_oracle.c's bsw_oracle_pair (== ksw_extend2, ksw.c:380-479; end_bonus =
 * pen_clip5 / pen_clip3 as the reference passes them).
 *
 * Parity pin: the .npz files under tests/golden/chain/ hold the mem_alnreg_t lists the reference's own
 * mem_chain2aln produced (oracle/_ref/libbwamemref.so = the unmodified tools/bwa sources,
 * tests/golden/make_golden_chain.py); tests/test_chain.py checks this file against them.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t o_del, e_del, o_ins, e_ins, zdrop, end_bonus;
    int32_t match, mismatch, ambig;
    int32_t zdrop_mode;
} oracle_params;                                   /* ksw_oracle.c */

int64_t bsw_oracle_pair(const oracle_params *p, const uint8_t *query, int qlen,
                        const uint8_t *target, int tlen, int w, int h0,
                        int32_t out[6], int32_t *row_trip);

typedef struct { int64_t rbeg; int32_t qbeg, len, score; } oracle_seed;                           /* mem_seed_t */
typedef struct { int64_t rb, re; int32_t qb, qe, score, truesc, w, seedcov, seedlen0, pad; } oracle_alnreg;

static int cal_max_gap(const oracle_params *p, int w, int qlen)                                   /* :620-628 */
{
    int l_del = (int)((double)(qlen * p->match - p->o_del) / p->e_del + 1.);
    int l_ins = (int)((double)(qlen * p->match - p->o_ins) / p->e_ins + 1.);
    int l = l_del > l_ins ? l_del : l_ins;
    l = l > 1 ? l : 1;
    return l < w << 1 ? l : w << 1;
}

void bsw_oracle_chain_window(const oracle_params *p, int w, int64_t l_pac, const oracle_seed *seeds, int n,
                             int l_query, int64_t *rmax0, int64_t *rmax1)                          /* :643-659 */
{
    int64_t r0 = l_pac << 1, r1 = 0;
    for (int i = 0; i < n; ++i) {
        const oracle_seed *t = &seeds[i];
        int64_t b = t->rbeg - (t->qbeg + cal_max_gap(p, w, t->qbeg));
        int64_t e = t->rbeg + t->len + ((l_query - t->qbeg - t->len) + cal_max_gap(p, w, l_query - t->qbeg - t->len));
        r0 = r0 < b ? r0 : b;
        r1 = r1 > e ? r1 : e;
    }
    r0 = r0 > 0 ? r0 : 0;
    r1 = r1 < l_pac << 1 ? r1 : l_pac << 1;
    if (r0 < l_pac && l_pac < r1) {
        if (seeds[0].rbeg < l_pac) r1 = l_pac;
        else r0 = l_pac;
    }
    *rmax0 = r0; *rmax1 = r1;
}

static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return (x > y) - (x < y);
}

/* One chain.  rseq = the window bytes [rmax0, rmax1).  prior[0..n_prior) = the regions the earlier chains
 * of the same read pushed into the shared mem_alnreg_v (mem_align1_core calls mem_chain2aln once per chain with
 * ONE vector, bwamem.c:1105-1112): the containment test sees them first.  out must hold n entries; returns how
 * many were made. */
int bsw_oracle_chain(const oracle_params *p0, int w, int pen_clip5, int pen_clip3, int max_band_try,
                     int l_query, const uint8_t *query, const oracle_seed *seeds, int n,
                     int64_t rmax0, int64_t rmax1, const uint8_t *rseq,
                     const oracle_alnreg *prior, int n_prior, oracle_alnreg *out)
{
    int n_out = 0;
    if (n == 0) return 0;
    uint64_t *srt = (uint64_t *)malloc((size_t)n * 8);
    for (int i = 0; i < n; ++i) srt[i] = (uint64_t)seeds[i].score << 32 | (uint32_t)i;
    qsort(srt, (size_t)n, 8, cmp_u64);
    for (int k = n - 1; k >= 0; --k) {
        const oracle_seed *s = &seeds[(uint32_t)srt[k]];
        int i;
        for (i = 0; i < n_prior + n_out; ++i) {                                                    /* :667-683 */
            const oracle_alnreg *q = i < n_prior ? &prior[i] : &out[i - n_prior];
            int64_t rd; int qd, ww, max_gap;
            if (s->rbeg < q->rb || s->rbeg + s->len > q->re || s->qbeg < q->qb || s->qbeg + s->len > q->qe) continue;
            if (s->len - q->seedlen0 > .1 * l_query) continue;
            qd = s->qbeg - q->qb; rd = s->rbeg - q->rb;
            max_gap = cal_max_gap(p0, w, qd < rd ? qd : (int)rd);
            ww = max_gap < q->w ? max_gap : q->w;
            if (qd - rd < ww && rd - qd < ww) break;
            qd = q->qe - (s->qbeg + s->len); rd = q->re - (s->rbeg + s->len);
            max_gap = cal_max_gap(p0, w, qd < rd ? qd : (int)rd);
            ww = max_gap < q->w ? max_gap : q->w;
            if (qd - rd < ww && rd - qd < ww) break;
        }
        if (i < n_prior + n_out) {                                                                 /* :684-700 */
            for (i = k + 1; i < n; ++i) {
                const oracle_seed *t;
                if (srt[i] == 0) continue;
                t = &seeds[(uint32_t)srt[i]];
                if (t->len < s->len * .95) continue;
                if (s->qbeg <= t->qbeg && s->qbeg + s->len - t->qbeg >= s->len >> 2 && t->qbeg - s->qbeg != t->rbeg - s->rbeg) break;
                if (t->qbeg <= s->qbeg && t->qbeg + t->len - s->qbeg >= s->len >> 2 && s->qbeg - t->qbeg != s->rbeg - t->rbeg) break;
            }
            if (i == n) { srt[k] = 0; continue; }
        }
        oracle_alnreg *a = &out[n_out++];
        memset(a, 0, sizeof(*a));
        int aw[2] = {w, w};
        a->w = w; a->score = a->truesc = -1;
        oracle_params p = *p0;
        int32_t r[6];
        if (s->qbeg) {                                                                             /* :709-763 */
            int64_t tmp = s->rbeg - rmax0;
            uint8_t *qs = (uint8_t *)malloc((size_t)s->qbeg), *rs = (uint8_t *)malloc((size_t)(tmp > 0 ? tmp : 1));
            for (i = 0; i < s->qbeg; ++i) qs[i] = query[s->qbeg - 1 - i];
            for (i = 0; i < tmp; ++i) rs[i] = rseq[tmp - 1 - i];
            p.end_bonus = pen_clip5;
            for (i = 0; i < max_band_try; ++i) {
                int prev = a->score;
                aw[0] = w << i;
                bsw_oracle_pair(&p, qs, s->qbeg, rs, (int)tmp, aw[0], s->len * p.match, r, NULL);
                a->score = r[0];
                if (a->score == prev || r[5] < (aw[0] >> 1) + (aw[0] >> 2)) break;
            }
            if (r[4] <= 0 || r[4] <= a->score - pen_clip5) {             /* gscore: local extension */
                a->qb = s->qbeg - r[1]; a->rb = s->rbeg - r[2];
                a->truesc = a->score;
            } else {                                                     /* to-end extension */
                a->qb = 0; a->rb = s->rbeg - r[3];
                a->truesc = r[4];
            }
            free(qs); free(rs);
        } else { a->score = a->truesc = s->len * p.match; a->qb = 0; a->rb = s->rbeg; }
        if (s->qbeg + s->len != l_query) {                                                         /* :765-810 */
            int qe = s->qbeg + s->len, sc0 = a->score;
            int64_t re = s->rbeg + s->len - rmax0;
            p.end_bonus = pen_clip3;
            for (i = 0; i < max_band_try; ++i) {
                int prev = a->score;
                aw[1] = w << i;
                bsw_oracle_pair(&p, query + qe, l_query - qe, rseq + re, (int)(rmax1 - rmax0 - re), aw[1], sc0, r, NULL);
                a->score = r[0];
                if (a->score == prev || r[5] < (aw[1] >> 1) + (aw[1] >> 2)) break;
            }
            if (r[4] <= 0 || r[4] <= a->score - pen_clip3) {
                a->qe = qe + r[1]; a->re = rmax0 + re + r[2];
                a->truesc += a->score - sc0;
            } else {
                a->qe = l_query; a->re = rmax0 + re + r[3];
                a->truesc += r[4] - sc0;
            }
        } else { a->qe = l_query; a->re = s->rbeg + s->len; }
        a->seedcov = 0;                                                                            /* :812-818 */
        for (i = 0; i < n; ++i) {
            const oracle_seed *t = &seeds[i];
            if (t->qbeg >= a->qb && t->qbeg + t->len <= a->qe && t->rbeg >= a->rb && t->rbeg + t->len <= a->re)
                a->seedcov += t->len;
        }
        a->w = aw[0] > aw[1] ? aw[0] : aw[1];
        a->seedlen0 = s->len;
    }
    free(srt);
    return n_out;
}
