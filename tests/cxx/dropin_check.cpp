// dropin_check.cpp -- test harness (not product code): drives the C++ drop-in class BandedPairWiseSW of
// include/bandedSWA.h (csrc/bsw_shim.cpp) exactly as a C++ caller of the reference would
// (benchmarks/bsw/main_banded.cpp:253-258,286,346-349) and dumps the six result fields of every method, so that
// pytest can compare them with the golden vectors of the reference's own getScores16 / scalarBandedSWA.
//
//   dropin_check <in.bin> <out.bin>
//   in.bin : int32 n, w, o_del, e_del, o_ins, e_ins, zdrop, end_bonus, match, mismatch, ambig; int64 ref_bytes, qer_bytes;
//            SeqPair[n]; ref bytes; qer bytes
//   out.bin: five blocks of n x 6 int32 (score qle tle gtle gscore max_off):
//            0 getScores16 (one call)          1 getScores8 (one call)       2 scalarBandedSWAWrapper
//            3 scalarBandedSWA, one pair at a time (first min(n, 48) pairs; the rest of the block is -1)
//            4 getScores16 in batches of 512 from 4 OpenMP threads, one instance per thread created inside the
//              parallel region (the driver's -t 4 -b 512 shape, main_banded.cpp:279-291)
#include "bandedSWA.h"
#include <cstdio>
#include <cstring>
#include <vector>
#include <omp.h>

static void fill_scmat(int a, int b, int ambig, int8_t mat[25])      // bwa_fill_scmat, main_banded.cpp:73-81
{
    int k = 0;
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 4; ++j) mat[k++] = (int8_t)(i == j ? a : -b);
        mat[k++] = (int8_t)ambig;
    }
    for (int j = 0; j < 5; ++j) mat[k++] = (int8_t)ambig;
}

static void put(std::vector<int32_t>& out, int block, int64_t n, int64_t i, const SeqPair& p)
{
    int32_t* o = out.data() + ((size_t)block * n + i) * 6;
    o[0] = p.score; o[1] = p.qle; o[2] = p.tle; o[3] = p.gtle; o[4] = p.gscore; o[5] = p.max_off;
}

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: dropin_check in.bin out.bin\n"); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    int32_t h[11];
    int64_t nb[2];
    if (fread(h, 4, 11, f) != 11 || fread(nb, 8, 2, f) != 2) return 2;
    const int64_t n = h[0];
    const int w = h[1];
    std::vector<SeqPair> pairs((size_t)n + 34);                       // capacity roundup(n, SIMD_WIDTH) + 2, like any caller
    std::vector<uint8_t> ref((size_t)nb[0] + 64), qer((size_t)nb[1] + 64);
    if (fread(pairs.data(), sizeof(SeqPair), n, f) != (size_t)n || fread(ref.data(), 1, nb[0], f) != (size_t)nb[0] ||
        fread(qer.data(), 1, nb[1], f) != (size_t)nb[1]) return 2;
    fclose(f);
    int8_t mat[25];
    fill_scmat(h[8], h[9], h[10], mat);
    std::vector<int32_t> out((size_t)5 * n * 6, -1);
    {
        BandedPairWiseSW sw(h[2], h[3], h[4], h[5], h[6], h[7], mat, (int8_t)h[8], (int8_t)h[9], 1);
        std::vector<SeqPair> p = pairs;
        sw.getScores16(p.data(), ref.data(), qer.data(), (int32_t)n, 1, w);
        for (int64_t i = 0; i < n; ++i) put(out, 0, n, i, p[i]);
        p = pairs;
        sw.getScores8(p.data(), ref.data(), qer.data(), (int32_t)n, 1, w);
        for (int64_t i = 0; i < n; ++i) put(out, 1, n, i, p[i]);
        p = pairs;
        sw.scalarBandedSWAWrapper(p.data(), ref.data(), qer.data(), (int)n, 1, w);
        for (int64_t i = 0; i < n; ++i) put(out, 2, n, i, p[i]);
        for (int64_t i = 0; i < n && i < 48; ++i) {
            SeqPair r = pairs[i];
            r.score = sw.scalarBandedSWA(r.len2, qer.data() + r.idq, r.len1, ref.data() + r.idr, w, r.h0,
                                         &r.qle, &r.tle, &r.gtle, &r.gscore, &r.max_off);
            put(out, 3, n, i, r);
        }
        if (sw.SW_cells == 0 || sw.getTicks() <= 0) { fprintf(stderr, "no cells / ticks counted\n"); return 3; }
    }
    {
        const int T = 4, B = 512;
        BandedPairWiseSW* sw[T] = {nullptr, nullptr, nullptr, nullptr};
        std::vector<SeqPair> p = pairs;
#pragma omp parallel num_threads(T)
        {
            const int tid = omp_get_thread_num();
            sw[tid] = new BandedPairWiseSW(h[2], h[3], h[4], h[5], h[6], h[7], mat, (int8_t)h[8], (int8_t)h[9], 1);
#pragma omp for schedule(dynamic, 1)
            for (int64_t i = 0; i < n; i += B) {
                const int32_t cnt = (int32_t)(n - i >= B ? B : n - i);
                sw[tid]->getScores16(p.data() + i, ref.data(), qer.data(), cnt, 1, w);
            }
        }
        for (int64_t i = 0; i < n; ++i) put(out, 4, n, i, p[i]);
        for (int t = 0; t < T; ++t) delete sw[t];
    }
    f = fopen(argv[2], "wb");
    if (!f) return 2;
    fwrite(out.data(), 4, out.size(), f);
    fclose(f);
    return 0;
}
