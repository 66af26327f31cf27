// bsw_pack.cpp -- builders of the packed host format (include/bsw.h: bsw_packed_batch).  Host only: no CUDA call.
//
// The reference's loader (benchmarks/bsw/main_banded.cpp:131-185 loadPairs, outside its timed region :262-270,306)
// fills SeqPair[] and one byte per base; the conversion to the kernel's layout (AoS -> SoA, bandedSWA.cpp:1266-1326)
// happens inside every getScores16 call.  Here the loader itself emits what the GPU kernels read: 2 bits per base,
// 16 bases per word, a 16-byte descriptor per pair, so that a call moves ~4x fewer bytes over PCIe.  Pairs that
// contain N (code 4) or whose query is too long for the thread-per-pair kernel keep one byte per base (RAW).
#include "bsw_common.h"
#include <cstdio>
#include <cstring>
#include <string>

using namespace bsw;

namespace {

inline void unpack_seq(const uint32_t* src, int len, uint8_t* dst)
{
    for (int k = 0; k < len; ++k) dst[k] = (uint8_t)((src[k >> 4] >> ((k & 15) * 2)) & 3u);
}

// highest base code of a sequence (> 3: N or invalid)
inline uint32_t max_code(const uint8_t* s, int len)
{
    uint64_t acc = 0;
    int k = 0;
    for (; k + 8 <= len; k += 8) { uint64_t v; memcpy(&v, s + k, 8); acc |= v; }
    uint32_t m = 0;
    for (; k < len; ++k) m |= s[k];
    if (acc & 0xF8F8F8F8F8F8F8F8ull) return 255;            // some code > 7
    for (int b = 0; b < 8; ++b) m |= (uint32_t)((acc >> (8 * b)) & 0xff);
    // m is an OR of codes: values 0..7; 4 is the only legal code above 3, and OR-ing codes <= 4 gives <= 7
    return m;
}

void* default_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void default_release(void* p) { free(p); }

struct Alloc {
    bsw_alloc_fn a; bsw_release_fn r;
    Alloc(bsw_alloc_fn a_, bsw_release_fn r_) : a(a_ ? a_ : default_alloc), r(r_ ? r_ : default_release) {}
};

int alloc_batch(const Alloc& A, bsw_packed_batch* b)
{
    // slack behind every buffer: the device side reads whole words / 16-byte groups
    b->desc = static_cast<bsw_pair_desc*>(A.a(((size_t)b->n_pairs + 4) * sizeof(bsw_pair_desc)));
    b->q2 = static_cast<uint32_t*>(A.a(((size_t)b->q2_words + 16) * 4));
    b->r2 = static_cast<uint32_t*>(A.a(((size_t)b->r2_words + 16) * 4));
    b->raw_q = static_cast<uint8_t*>(A.a((size_t)b->raw_q_bytes + 64));
    b->raw_r = static_cast<uint8_t*>(A.a((size_t)b->raw_r_bytes + 64));
    if (!b->desc || !b->q2 || !b->r2 || !b->raw_q || !b->raw_r) { bsw_batch_release(b, A.r); return BSW_ERR_NOMEM; }
    memset(b->q2 + b->q2_words, 0, 64); memset(b->r2 + b->r2_words, 0, 64);
    memset(b->raw_q + b->raw_q_bytes, 0, 64); memset(b->raw_r + b->raw_r_bytes, 0, 64);
    memset(b->desc + b->n_pairs, 0, 4 * sizeof(bsw_pair_desc));
    return BSW_OK;
}

struct BlockSum { int64_t qw = 0, rw = 0, rq = 0, rr = 0; };
constexpr int64_t BLK = 4096;

} // namespace

extern "C" {

void bsw_batch_release(bsw_packed_batch* b, bsw_release_fn release)
{
    if (!b) return;
    bsw_release_fn r = release ? release : default_release;
    if (b->desc) r(b->desc);
    if (b->q2) r(b->q2);
    if (b->r2) r(b->r2);
    if (b->raw_q) r(b->raw_q);
    if (b->raw_r) r(b->raw_r);
    memset(b, 0, sizeof(*b));
}

int bsw_batch_from_pairs(const SeqPair* pairs, const uint8_t* seq_ref, const uint8_t* seq_qer, int64_t n,
                         int32_t raw_min_qlen, bsw_alloc_fn alloc, bsw_release_fn release, bsw_packed_batch* out)
{
    if (!out || n < 0 || n > 0x7fffffff || (n > 0 && (!pairs || !seq_ref || !seq_qer)) || raw_min_qlen < 0) return BSW_ERR_PARAM;
    memset(out, 0, sizeof(*out));
    const int raw_min = raw_min_qlen > 0 ? std::min(raw_min_qlen, BSW_PACKED_MAX_QLEN + 1) : BSW_PACKED_MAX_QLEN + 1;
    const Alloc A(alloc, release);
    ThreadPool& pool = global_pool();
    const int64_t nblk = (n + BLK - 1) / BLK;
    std::vector<BlockSum> sums((size_t)nblk + 1);
    std::vector<uint8_t> raw((size_t)n);
    std::atomic<int> bad{0};
    // pass 1: validate, find the RAW pairs, per-block totals
    pool.run(nblk, [&](int64_t b, int) {
        BlockSum s;
        const int64_t lo = b * BLK, hi = std::min(n, lo + BLK);
        for (int64_t k = lo; k < hi; ++k) {
            const SeqPair& sp = pairs[k];
            if (sp.len1 < 1 || sp.len1 > 32767 || sp.len2 < 1 || sp.len2 > 32767 || sp.h0 < 0 || sp.h0 > 32767 ||
                sp.idr < 0 || sp.idq < 0) { bad.store(1, std::memory_order_relaxed); raw[(size_t)k] = 0; continue; }
            const uint32_t mq = max_code(seq_qer + sp.idq, sp.len2), mr = max_code(seq_ref + sp.idr, sp.len1);
            const uint32_t m = mq | mr;
            bool is_raw = sp.len2 >= raw_min;
            if (m > 3) {
                // OR of legal codes is <= 7 whenever a 4 is present; make sure no single code exceeds 4
                for (int j = 0; j < sp.len2; ++j) if (seq_qer[sp.idq + j] > 4) bad.store(1, std::memory_order_relaxed);
                for (int j = 0; j < sp.len1; ++j) if (seq_ref[sp.idr + j] > 4) bad.store(1, std::memory_order_relaxed);
                is_raw = true;
            }
            raw[(size_t)k] = is_raw ? 1 : 0;
            if (is_raw) { s.rq += sp.len2; s.rr += sp.len1; }
            else { s.qw += (sp.len2 + 15) >> 4; s.rw += (sp.len1 + 15) >> 4; }
        }
        sums[(size_t)b + 1] = s;
    });
    if (bad.load()) return BSW_ERR_DOMAIN;
    for (int64_t b = 0; b < nblk; ++b) {
        sums[(size_t)b + 1].qw += sums[(size_t)b].qw; sums[(size_t)b + 1].rw += sums[(size_t)b].rw;
        sums[(size_t)b + 1].rq += sums[(size_t)b].rq; sums[(size_t)b + 1].rr += sums[(size_t)b].rr;
    }
    const BlockSum& T = sums[(size_t)nblk];
    if (T.qw > 0xffffffffll || T.rw > 0xffffffffll || T.rq > 0xffffffffll || T.rr > 0xffffffffll) return BSW_ERR_PARAM;
    out->n_pairs = n; out->q2_words = T.qw; out->r2_words = T.rw; out->raw_q_bytes = T.rq; out->raw_r_bytes = T.rr;
    out->ordered = 1;
    if (int rc = alloc_batch(A, out)) return rc;
    // pass 2: descriptors + sequences
    pool.run(nblk, [&](int64_t b, int) {
        BlockSum s = sums[(size_t)b];
        const int64_t lo = b * BLK, hi = std::min(n, lo + BLK);
        for (int64_t k = lo; k < hi; ++k) {
            const SeqPair& sp = pairs[k];
            bsw_pair_desc& d = out->desc[k];
            d.len2 = (uint16_t)sp.len2; d.len1 = (uint16_t)sp.len1; d.h0 = (uint16_t)sp.h0;
            if (raw[(size_t)k]) {
                d.flags = BSW_PAIR_RAW; d.q_off = (uint32_t)s.rq; d.r_off = (uint32_t)s.rr;
                memcpy(out->raw_q + s.rq, seq_qer + sp.idq, (size_t)sp.len2);
                memcpy(out->raw_r + s.rr, seq_ref + sp.idr, (size_t)sp.len1);
                s.rq += sp.len2; s.rr += sp.len1;
            } else {
                d.flags = 0; d.q_off = (uint32_t)s.qw; d.r_off = (uint32_t)s.rw;
                pack_seq(seq_qer + sp.idq, sp.len2, out->q2 + s.qw);
                pack_seq(seq_ref + sp.idr, sp.len1, out->r2 + s.rw);
                s.qw += (sp.len2 + 15) >> 4; s.rw += (sp.len1 + 15) >> 4;
            }
        }
    });
    return BSW_OK;
}

int bsw_batch_to_pairs(const bsw_packed_batch* b, SeqPair* pairs, uint8_t* seq_ref, int64_t ref_cap,
                       uint8_t* seq_qer, int64_t qer_cap)
{
    if (!b || b->n_pairs < 0 || (b->n_pairs > 0 && (!pairs || !seq_ref || !seq_qer || !b->desc))) return BSW_ERR_PARAM;
    int64_t ro = 0, qo = 0;
    for (int64_t k = 0; k < b->n_pairs; ++k) {
        const bsw_pair_desc& d = b->desc[k];
        if (d.len1 < 1 || d.len2 < 1) return BSW_ERR_DOMAIN;
        if (ro + d.len1 > ref_cap || qo + d.len2 > qer_cap) return BSW_ERR_NOMEM;
        SeqPair& sp = pairs[k];
        memset(&sp, 0, sizeof(sp));
        sp.id = k; sp.idr = ro; sp.idq = qo; sp.len1 = d.len1; sp.len2 = d.len2; sp.h0 = d.h0;
        sp.seqid = sp.regid = -1;
        sp.score = sp.tle = sp.gtle = sp.qle = sp.gscore = sp.max_off = -1;
        if (d.flags & BSW_PAIR_RAW) {
            if ((int64_t)d.q_off + d.len2 > b->raw_q_bytes || (int64_t)d.r_off + d.len1 > b->raw_r_bytes) return BSW_ERR_DOMAIN;
            memcpy(seq_qer + qo, b->raw_q + d.q_off, d.len2);
            memcpy(seq_ref + ro, b->raw_r + d.r_off, d.len1);
        } else {
            if ((int64_t)d.q_off + ((d.len2 + 15) >> 4) > b->q2_words || (int64_t)d.r_off + ((d.len1 + 15) >> 4) > b->r2_words)
                return BSW_ERR_DOMAIN;
            unpack_seq(b->q2 + d.q_off, d.len2, seq_qer + qo);
            unpack_seq(b->r2 + d.r_off, d.len1, seq_ref + ro);
        }
        ro += d.len1; qo += d.len2;
    }
    return BSW_OK;
}

int bsw_batch_from_file(const char* path, int64_t max_pairs, int32_t raw_min_qlen, bsw_alloc_fn alloc,
                        bsw_release_fn release, bsw_packed_batch* out)
{
    if (!path || !out) return BSW_ERR_PARAM;
    int64_t n = bsw_count_pairs_file(path);
    if (n < 0) return (int)n;
    if (max_pairs >= 0) n = std::min(n, max_pairs);
    FILE* f = fopen(path, "rb");
    if (!f) return BSW_ERR_IO;
    fseek(f, 0, SEEK_END);
    const int64_t cap = (int64_t)ftell(f) + 64;          // the file holds one character per base
    fclose(f);
    std::vector<SeqPair> pairs((size_t)n + 1);
    std::vector<uint8_t> ref((size_t)cap), qer((size_t)cap);
    int64_t got = 0;
    if (int rc = bsw_read_pairs_file(path, n, pairs.data(), ref.data(), cap, qer.data(), cap, &got)) return rc;
    return bsw_batch_from_pairs(pairs.data(), ref.data(), qer.data(), got, raw_min_qlen, alloc, release, out);
}

int bsw_batch_gen(const bsw_gen_config* cfg, int64_t first, int64_t n, int32_t raw_min_qlen,
                  bsw_alloc_fn alloc, bsw_release_fn release, bsw_packed_batch* out)
{
    if (!cfg || !out || n < 0 || first < 0) return BSW_ERR_PARAM;
    // generated in slices through the byte layout (bsw_gen_pairs is the one definition of the stream), each slice
    // packed and appended; slices are sized so the temporary bytes stay small
    const int64_t SLICE = 1 << 18;
    bsw_gen_config sub = *cfg;
    sub.n_pairs = std::min(n, SLICE);
    int64_t rb = 0, qb = 0;
    if (int rc = bsw_gen_bounds(&sub, &rb, &qb)) return rc;
    std::vector<SeqPair> pairs((size_t)sub.n_pairs + 1);
    std::vector<uint8_t> ref((size_t)rb), qer((size_t)qb);
    std::vector<bsw_packed_batch> parts;
    const Alloc A(alloc, release);
    int rc = BSW_OK;
    bsw_packed_batch total;
    memset(&total, 0, sizeof(total));
    for (int64_t a = 0; a < n && rc == BSW_OK; a += SLICE) {
        const int64_t m = std::min(SLICE, n - a);
        int64_t ru = 0, qu = 0;
        rc = bsw_gen_pairs(cfg, first + a, m, pairs.data(), ref.data(), qer.data(), &ru, &qu);
        if (rc) break;
        bsw_packed_batch part;
        rc = bsw_batch_from_pairs(pairs.data(), ref.data(), qer.data(), m, raw_min_qlen, nullptr, nullptr, &part);
        if (rc) break;
        parts.push_back(part);
        total.n_pairs += part.n_pairs; total.q2_words += part.q2_words; total.r2_words += part.r2_words;
        total.raw_q_bytes += part.raw_q_bytes; total.raw_r_bytes += part.raw_r_bytes;
    }
    if (rc == BSW_OK && (total.q2_words > 0xffffffffll || total.r2_words > 0xffffffffll ||
                         total.raw_q_bytes > 0xffffffffll || total.raw_r_bytes > 0xffffffffll)) rc = BSW_ERR_PARAM;
    if (rc == BSW_OK) { total.ordered = 1; rc = alloc_batch(A, &total); }
    if (rc == BSW_OK) {
        int64_t np = 0, qw = 0, rw = 0, rq = 0, rr = 0;
        for (bsw_packed_batch& p : parts) {
            for (int64_t k = 0; k < p.n_pairs; ++k) {
                bsw_pair_desc d = p.desc[k];
                if (d.flags & BSW_PAIR_RAW) { d.q_off += (uint32_t)rq; d.r_off += (uint32_t)rr; }
                else { d.q_off += (uint32_t)qw; d.r_off += (uint32_t)rw; }
                total.desc[np + k] = d;
            }
            memcpy(total.q2 + qw, p.q2, (size_t)p.q2_words * 4); memcpy(total.r2 + rw, p.r2, (size_t)p.r2_words * 4);
            memcpy(total.raw_q + rq, p.raw_q, (size_t)p.raw_q_bytes); memcpy(total.raw_r + rr, p.raw_r, (size_t)p.raw_r_bytes);
            np += p.n_pairs; qw += p.q2_words; rw += p.r2_words; rq += p.raw_q_bytes; rr += p.raw_r_bytes;
        }
        *out = total;
    }
    for (bsw_packed_batch& p : parts) bsw_batch_release(&p, nullptr);
    return rc;
}

} // extern "C"
