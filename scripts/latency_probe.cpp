// latency_probe.cpp -- small-call throughput of the drop-in, measured from C++ (no interpreter in the way):
//   blocking   T OpenMP threads, each calling BandedPairWiseSW::getScores16 with 512 pairs per call -- the reference
//              driver's loop (main_banded.cpp:279-291 with scripts/run-cpu.sh:30), steady state (engines warm)
//   async      T threads, each keeping D bsw_extend_async calls of 512 pairs in flight
// usage: latency_probe <pairs.bin> ; pairs.bin = int64 n, ref_bytes, qer_bytes; SeqPair[n]; ref; qer  (scripts/latency_probe.py writes it)
#include "bandedSWA.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <omp.h>

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    int64_t h[3];
    if (fread(h, 8, 3, f) != 3) return 2;
    const int64_t n = h[0];
    std::vector<SeqPair> pairs((size_t)n + 34), want;
    std::vector<uint8_t> ref((size_t)h[1] + 64), qer((size_t)h[2] + 64);
    if (fread(pairs.data(), sizeof(SeqPair), n, f) != (size_t)n || fread(ref.data(), 1, h[1], f) != (size_t)h[1] ||
        fread(qer.data(), 1, h[2], f) != (size_t)h[2]) return 2;
    fclose(f);
    int8_t mat[25];
    { int k = 0; for (int i = 0; i < 4; ++i) { for (int j = 0; j < 4; ++j) mat[k++] = i == j ? 1 : -4; mat[k++] = -1; } for (int j = 0; j < 5; ++j) mat[k++] = -1; }
    const int B = 512, w = 100;
    {   // expected results: one big call
        BandedPairWiseSW sw(6, 1, 6, 1, 100, 5, mat, 1, 4, 1);
        want = pairs;
        sw.getScores16(want.data(), ref.data(), qer.data(), (int32_t)n, 1, w);
    }
    auto same = [&](const std::vector<SeqPair>& got) {
        for (int64_t i = 0; i < n; ++i)
            if (got[i].score != want[i].score || got[i].qle != want[i].qle || got[i].tle != want[i].tle || got[i].gtle != want[i].gtle ||
                got[i].gscore != want[i].gscore || got[i].max_off != want[i].max_off) return false;
        return true;
    };
    printf("blocking getScores16, %d pairs per call, T threads (one BandedPairWiseSW per thread), %lld pairs per pass:\n", B, (long long)n);
    for (int T : {1, 2, 4, 8, 16, 32, 64}) {
        std::vector<BandedPairWiseSW*> sw((size_t)T);
        for (int t = 0; t < T; ++t) sw[(size_t)t] = new BandedPairWiseSW(6, 1, 6, 1, 100, 5, mat, 1, 4, 1);
        std::vector<SeqPair> got = pairs;
        double best = 1e30;
        for (int pass = 0; pass < 3; ++pass) {
            const double t0 = now_s();
#pragma omp parallel num_threads(T)
            {
                const int tid = omp_get_thread_num();
#pragma omp for schedule(dynamic, 1)
                for (int64_t i = 0; i < n; i += B)
                    sw[(size_t)tid]->getScores16(got.data() + i, ref.data(), qer.data(), (int32_t)(n - i >= B ? B : n - i), 1, w);
            }
            const double dt = now_s() - t0;
            if (pass > 0 && dt < best) best = dt;
        }
        printf("  T=%2d  %8.2f M pairs/s   results %s\n", T, n / best / 1e6, same(got) ? "ok" : "DIFFER");
        fflush(stdout);
        for (auto* p : sw) delete p;
    }
    if (getenv("LAT_BLOCKING_ONLY")) return 0;
    printf("bsw_extend_async, %d pairs per call, T threads x D calls in flight:\n", B);
    bsw_params P;
    bsw_default_params(&P);
    P.tiny_batch = 1536;
    int err = 0;
    bsw_engine* eng = bsw_create(&P, &err);
    if (!eng) return 3;
    const int combos[][2] = {{1, 8}, {1, 32}, {8, 2}, {8, 4}, {8, 8}, {16, 4}, {16, 8}};
    for (auto& c : combos) {
        const int T = c[0], D = c[1];
        std::vector<SeqPair> got = pairs;
        double best = 1e30;
        const int64_t ncall = (n + B - 1) / B;
        for (int pass = 0; pass < 3; ++pass) {
            const double t0 = now_s();
#pragma omp parallel num_threads(T)
            {
                const int tid = omp_get_thread_num();
                std::vector<int64_t> tk;
                for (int64_t c0 = (int64_t)tid * D; c0 < ncall; c0 += (int64_t)T * D) {
                    tk.clear();
                    for (int64_t cc = c0; cc < c0 + D && cc < ncall; ++cc) {
                        int64_t t = 0;
                        const int64_t i = cc * B;
                        bsw_extend_async(eng, got.data() + i, ref.data(), qer.data(), n - i >= B ? B : n - i, w, &t);
                        tk.push_back(t);
                    }
                    for (int64_t t : tk) bsw_wait(eng, t, nullptr);
                }
            }
            const double dt = now_s() - t0;
            if (pass > 0 && dt < best) best = dt;
        }
        int64_t calls = 0, batches = 0;
        bsw_async_stats(eng, &calls, &batches);
        printf("  T=%2d D=%2d  %8.2f M pairs/s   results %s   (calls per batch so far %.1f)\n", T, D, n / best / 1e6,
               same(got) ? "ok" : "DIFFER", batches ? (double)calls / batches : 0.0);
        fflush(stdout);
    }
    bsw_destroy(eng);
    return 0;
}
