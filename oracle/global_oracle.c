/*
 * global_oracle.c -- CPU restatement of the banded global alignment with traceback that turns an
 * alignment region into a CIGAR.
 *
 * TEST INFRASTRUCTURE ONLY (same rule as ksw_oracle.c: nothing under genomicsbench_b200/ may link,
 * import or call this file).
 *
 * What it restates (paths relative to /root/reference/tools/bwa):
 *   ksw.c:489-500   push_cigar (runs of one operation are merged; op 0 = M, 1 = I, 2 = D; len << 4 | op)
 *   ksw.c:502-606   ksw_global2: fixed band |i - j| <= w, first row / first column gap costs, the
 *                   direction byte f << 4 | e << 2 | h of every cell, the backtrack from the last cell
 *                   and the reversal of the operation list
 * Scoring: bwa_fill_scmat (bwa.c; benchmarks/bsw/main_banded.cpp:73-81): match / -mismatch, -1 against N.
 *
 * Parity pin: the reference's own ksw_global2, compiled unmodified into oracle/_ref/libkswref.so,
 * produced the golden scores and CIGARs under tests/golden/global/ (tests/golden/make_golden_global.py);
 * tests/test_global.py checks this file against them and, where the reference tree is mounted, against
 * live calls.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define G_MINUS_INF (-0x40000000)

typedef struct {
    int32_t o_del, e_del, o_ins, e_ins, zdrop, end_bonus;
    int32_t match, mismatch, ambig;
    int32_t zdrop_mode;
} oracle_params;                                   /* ksw_oracle.c */

static inline int gsc(const oracle_params *p, int x, int y)
{
    if (x >= 4 || y >= 4) return p->ambig;
    return x == y ? p->match : -p->mismatch;
}

static int g_push(uint32_t *cigar, int n, int op, int len)                                   /* ksw.c:489-500 */
{
    if (n == 0 || op != (int)(cigar[n - 1] & 0xf)) cigar[n++] = (uint32_t)len << 4 | (uint32_t)op;
    else cigar[n - 1] += (uint32_t)len << 4;
    return n;
}

/* cigar must hold qlen + tlen entries.  Returns the score; *n_cigar receives the number of operations. */
int bsw_oracle_global(const oracle_params *p, const uint8_t *query, int qlen, const uint8_t *target, int tlen,
                      int w, uint32_t *cigar, int *n_cigar)
{
    const int o_del = p->o_del, e_del = p->e_del, o_ins = p->o_ins, e_ins = p->e_ins;
    const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    const int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
    int32_t *eh_h = (int32_t *)malloc(((size_t)qlen + 1) * 4), *eh_e = (int32_t *)malloc(((size_t)qlen + 1) * 4);
    uint8_t *z = (uint8_t *)malloc((size_t)n_col * (size_t)tlen + 1);
    int i, j, k, score;
    eh_h[0] = 0; eh_e[0] = G_MINUS_INF;                                                         /* ksw.c:521-525 */
    for (j = 1; j <= qlen && j <= w; ++j) { eh_h[j] = -(o_ins + e_ins * j); eh_e[j] = G_MINUS_INF; }
    for (; j <= qlen; ++j) eh_h[j] = eh_e[j] = G_MINUS_INF;
    for (i = 0; i < tlen; ++i) {                                                                /* :527-589 */
        int32_t f = G_MINUS_INF, h1, t;
        const int beg = i > w ? i - w : 0;
        const int end = i + w + 1 < qlen ? i + w + 1 : qlen;
        uint8_t *zi = z + (size_t)i * n_col;
        h1 = beg == 0 ? -(o_del + e_del * (i + 1)) : G_MINUS_INF;
        for (j = beg; j < end; ++j) {
            int32_t h, m = eh_h[j], e = eh_e[j];
            uint8_t d;
            eh_h[j] = h1;
            m += gsc(p, target[i], query[j]);
            d = m >= e ? 0 : 1;
            h = m >= e ? m : e;
            d = h >= f ? d : 2;
            h = h >= f ? h : f;
            h1 = h;
            t = m - oe_del;
            e -= e_del;
            d |= e > t ? 1 << 2 : 0;
            e = e > t ? e : t;
            eh_e[j] = e;
            t = m - oe_ins;
            f -= e_ins;
            d |= f > t ? 2 << 4 : 0;
            f = f > t ? f : t;
            zi[j - beg] = d;
        }
        eh_h[end] = h1; eh_e[end] = G_MINUS_INF;
    }
    score = eh_h[qlen];
    {                                                                                           /* :591-603 */
        int n = 0, which = 0;
        i = tlen - 1; k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
        while (i >= 0 && k >= 0) {
            which = z[(size_t)i * n_col + (k - (i > w ? i - w : 0))] >> (which << 1) & 3;
            if (which == 0) { n = g_push(cigar, n, 0, 1); --i; --k; }
            else if (which == 1) { n = g_push(cigar, n, 2, 1); --i; }
            else { n = g_push(cigar, n, 1, 1); --k; }
        }
        if (i >= 0) n = g_push(cigar, n, 2, i + 1);
        if (k >= 0) n = g_push(cigar, n, 1, k + 1);
        for (i = 0; i < n >> 1; ++i) { uint32_t tmp = cigar[i]; cigar[i] = cigar[n - 1 - i]; cigar[n - 1 - i] = tmp; }
        *n_cigar = n;
    }
    free(eh_h); free(eh_e); free(z);
    return score;
}
