// bsw_gen.cpp -- synthetic read/reference-window generator and the 3-line text format.
//
// The reference ships no generator: its inputs were dumped from bwa-mem
// (tools/bwa/bwamem.c:741-745,788-792) into the text format that
// benchmarks/bsw/main_banded.cpp:131-185 (loadPairs) parses.  This file provides
//   * the seeded synthetic configs of SURVEY.md 8(d) / BASELINE.json `configs`, and
//   * a reader / writer of that text format, so published inputs can be run when present.
#include "bsw_common.h"
#include <cstdio>
#include <cstring>
#include <string>

using namespace bsw;

namespace {

struct PairDraw {
    int32_t qlen, tlen, h0;
};

// Generates one pair.  When q / t are null only the lengths are produced (same RNG path).
PairDraw draw_pair(const bsw_gen_config& c, int64_t idx, uint8_t* q, uint8_t* t)
{
    SplitMix64 rng(c.seed + 0xD1B54A32D192ED03ull * (uint64_t)(idx + 1));
    PairDraw d;
    d.qlen = rng.range(c.qlen_min, c.qlen_max);
    int32_t h0_hi = c.h0_max;
    if (c.max_score8 > 0) h0_hi = std::min(h0_hi, c.max_score8 - d.qlen * std::max(c.match, 1));
    d.h0 = rng.range(std::min(c.h0_min, std::max(h0_hi, 1)), std::max(h0_hi, 1));
    const int32_t cap = c.max_len1 > 0 ? c.max_len1 : 0x7fffffff;
    const double r3 = c.error_rate / 3.0;
    int32_t n = 0;
    for (int32_t j = 0; j < d.qlen; ++j) {
        uint8_t b = (uint8_t)(rng.next() & 3);
        if (c.n_rate > 0 && rng.unit() < c.n_rate) b = 4;
        if (q) q[j] = b;
        const double u = rng.unit();
        if (u < r3) {                                   // substitution
            uint8_t s = (uint8_t)((b + 1 + (rng.next() % 3)) & 3);
            if (n < cap) { if (t) t[n] = s; ++n; }
        } else if (u < 2 * r3) {                        // 1-base insertion in the reference
            uint8_t x = (uint8_t)(rng.next() & 3);
            if (n < cap) { if (t) t[n] = x; ++n; }
            if (n < cap) { if (t) t[n] = b; ++n; }
        } else if (u < c.error_rate) {                  // deletion from the reference
        } else {
            if (n < cap) { if (t) t[n] = b; ++n; }
        }
    }
    const int32_t tail = rng.range(c.tail_min, c.tail_max);
    for (int32_t k = 0; k < tail; ++k) {
        uint8_t x = (uint8_t)(rng.next() & 3);
        if (n < cap) { if (t) t[n] = x; ++n; }
    }
    if (n == 0) { if (t) t[0] = 0; n = 1; }
    d.tlen = n;
    return d;
}

} // namespace

extern "C" {

int bsw_gen_named_config(int32_t which, bsw_gen_config* o)
{
    if (!o) return BSW_ERR_PARAM;
    memset(o, 0, sizeof(*o));
    o->seed = 0xB5B20000ull + (uint64_t)which;
    o->match = 1;
    o->h0_min = 19; o->h0_max = 60;
    switch (which) {
    case 0:  // bsw small: 10k x 151bp, ~251bp window
        o->n_pairs = 10000; o->qlen_min = o->qlen_max = 151; o->error_rate = 0.02;
        o->tail_min = o->tail_max = 100; break;
    case 1:  // 8-bit path: 1M short pairs inside the int8 envelope
        o->n_pairs = 1000000; o->qlen_min = 16; o->qlen_max = 96; o->error_rate = 0.02;
        o->tail_min = 0; o->tail_max = 20; o->max_len1 = 127; o->max_score8 = 127; break;
    case 2:  // 16-bit path: 1M x 250bp, long windows, high scores
        o->n_pairs = 1000000; o->qlen_min = o->qlen_max = 250; o->error_rate = 0.02;
        o->tail_min = 300; o->tail_max = 500; o->h0_min = 100; o->h0_max = 250; break;
    case 3:  // large: 50M mixed 50-300bp
        o->n_pairs = 50000000; o->qlen_min = 50; o->qlen_max = 300; o->error_rate = 0.05;
        o->tail_min = 50; o->tail_max = 150; break;
    case 4:  // band / zdrop sweep: 8M divergent pairs
        o->n_pairs = 8000000; o->qlen_min = 50; o->qlen_max = 300; o->error_rate = 0.10;
        o->tail_min = 50; o->tail_max = 150; break;
    default: return BSW_ERR_PARAM;
    }
    return BSW_OK;
}

int bsw_gen_bounds(const bsw_gen_config* c, int64_t* ref_bytes, int64_t* qer_bytes)
{
    if (!c || c->n_pairs < 0 || c->qlen_min < 1 || c->qlen_max < c->qlen_min) return BSW_ERR_PARAM;
    int64_t maxt = 2ll * c->qlen_max + c->tail_max + 1;
    if (c->max_len1 > 0) maxt = std::min<int64_t>(maxt, c->max_len1);
    if (ref_bytes) *ref_bytes = maxt * c->n_pairs + 64;
    if (qer_bytes) *qer_bytes = (int64_t)c->qlen_max * c->n_pairs + 64;
    return BSW_OK;
}

int bsw_gen_pairs(const bsw_gen_config* c, int64_t first, int64_t n, SeqPair* pairs,
                  uint8_t* seq_ref, uint8_t* seq_qer, int64_t* ref_used, int64_t* qer_used)
{
    if (!c || !pairs || !seq_ref || !seq_qer || n < 0 || first < 0) return BSW_ERR_PARAM;
    if (c->qlen_min < 1 || c->qlen_max < c->qlen_min || c->qlen_max > 32767) return BSW_ERR_PARAM;
    ThreadPool& pool = global_pool();
    // pass 1: lengths (same RNG path as pass 2)
    pool.for_range(n, 4096, [&](int64_t b, int64_t e, int) {
        for (int64_t k = b; k < e; ++k) {
            PairDraw d = draw_pair(*c, first + k, nullptr, nullptr);
            SeqPair& sp = pairs[k];
            memset(&sp, 0, sizeof(sp));
            sp.id = k; sp.len1 = d.tlen; sp.len2 = d.qlen; sp.h0 = d.h0;
            sp.seqid = sp.regid = -1;
            sp.score = sp.tle = sp.gtle = sp.qle = sp.gscore = sp.max_off = -1;   // main_banded.cpp:179-180
        }
    });
    int64_t ro = 0, qo = 0;
    for (int64_t k = 0; k < n; ++k) {
        pairs[k].idr = ro; pairs[k].idq = qo;
        ro += pairs[k].len1; qo += pairs[k].len2;
    }
    // pass 2: bases
    pool.for_range(n, 4096, [&](int64_t b, int64_t e, int) {
        for (int64_t k = b; k < e; ++k)
            draw_pair(*c, first + k, seq_qer + pairs[k].idq, seq_ref + pairs[k].idr);
    });
    if (ref_used) *ref_used = ro;
    if (qer_used) *qer_used = qo;
    return BSW_OK;
}

// ---------------------------------------------------------------------------------------
// Text format (main_banded.cpp:131-141): line 1 h0, line 2 reference digits, line 3 query
// digits, every line '\n'-terminated; the loader subtracts '0' (main_banded.cpp:173-178).
// ---------------------------------------------------------------------------------------
int64_t bsw_count_pairs_file(const char* path)
{
    FILE* f = fopen(path, "rb");
    if (!f) return BSW_ERR_IO;
    std::vector<char> buf(1 << 20);
    int64_t lines = 0; size_t got;
    while ((got = fread(buf.data(), 1, buf.size(), f)) > 0)
        for (size_t i = 0; i < got; ++i) lines += buf[i] == '\n';
    fclose(f);
    return lines / 3;                                   // main_banded.cpp:235
}

static bool read_line(FILE* f, std::string& s)
{
    s.clear();
    int ch;
    while ((ch = fgetc(f)) != EOF) {
        if (ch == '\n') return true;
        if (ch != '\r') s.push_back((char)ch);
    }
    return !s.empty();
}

int bsw_read_pairs_file(const char* path, int64_t max_pairs, SeqPair* pairs, uint8_t* seq_ref,
                        int64_t ref_cap, uint8_t* seq_qer, int64_t qer_cap, int64_t* n_read)
{
    if (!path || !pairs || !seq_ref || !seq_qer) return BSW_ERR_PARAM;
    FILE* f = fopen(path, "rb");
    if (!f) return BSW_ERR_IO;
    std::string l1, l2, l3;
    int64_t n = 0, ro = 0, qo = 0;
    int rc = BSW_OK;
    while (n < max_pairs && read_line(f, l1)) {
        if (!read_line(f, l2) || !read_line(f, l3)) break;     // odd trailing lines: stop like :155-158
        if (l2.empty() || l3.empty()) { rc = BSW_ERR_DOMAIN; break; }   // assert(len > 0), :166-167
        if (ro + (int64_t)l2.size() > ref_cap || qo + (int64_t)l3.size() > qer_cap) { rc = BSW_ERR_NOMEM; break; }
        SeqPair& sp = pairs[n];
        memset(&sp, 0, sizeof(sp));
        sp.id = n; sp.h0 = atoi(l1.c_str());
        sp.len1 = (int32_t)l2.size(); sp.len2 = (int32_t)l3.size();
        sp.idr = ro; sp.idq = qo;
        for (char ch : l2) seq_ref[ro++] = (uint8_t)(ch - '0');
        for (char ch : l3) seq_qer[qo++] = (uint8_t)(ch - '0');
        sp.seqid = sp.regid = -1;
        sp.score = sp.tle = sp.gtle = sp.qle = sp.gscore = sp.max_off = -1;
        ++n;
    }
    fclose(f);
    if (n_read) *n_read = n;
    return rc;
}

int bsw_write_pairs_file(const char* path, const SeqPair* pairs, int64_t n, const uint8_t* seq_ref,
                         const uint8_t* seq_qer)
{
    if (!path || !pairs || !seq_ref || !seq_qer) return BSW_ERR_PARAM;
    FILE* f = fopen(path, "wb");
    if (!f) return BSW_ERR_IO;
    std::string line;
    for (int64_t k = 0; k < n; ++k) {
        const SeqPair& sp = pairs[k];
        fprintf(f, "%d\n", sp.h0);
        line.assign((size_t)sp.len1, '0');
        for (int32_t i = 0; i < sp.len1; ++i) line[i] = (char)('0' + seq_ref[sp.idr + i]);
        fwrite(line.data(), 1, line.size(), f); fputc('\n', f);
        line.assign((size_t)sp.len2, '0');
        for (int32_t i = 0; i < sp.len2; ++i) line[i] = (char)('0' + seq_qer[sp.idq + i]);
        fwrite(line.data(), 1, line.size(), f); fputc('\n', f);
    }
    fclose(f);
    return BSW_OK;
}

} // extern "C"
