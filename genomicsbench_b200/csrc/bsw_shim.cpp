// bsw_shim.cpp -- BandedPairWiseSW (include/bandedSWA.h) on top of the C ABI.
// Mirrors the constructor / method behaviour of benchmarks/bsw/bandedSWA.cpp:51-122,
// :254-272, :424-446, :1124-1148.  No DP arithmetic lives here: every call is a bsw_extend().
#include "../../include/bandedSWA.h"
#include <cstring>
#include <chrono>
#include <mutex>
#include <vector>
#include <utility>

namespace {
inline int64_t ticks_now()
{
#if defined(__x86_64__) || defined(__i386__)
    return (int64_t)__rdtsc();
#else
    return (int64_t)std::chrono::steady_clock::now().time_since_epoch().count();
#endif
}
// BSW_SHIM_DUMP=<file> in the environment: every getScores* / scalar call appends one record
//   { uint64 address of pairArray, int64 n, n x 6 int32 (score qle tle gtle gscore max_off) }
// The reference driver never emits results (main_banded.cpp prints timings only, SURVEY finding 0.3); this is how a
// run of the UNMODIFIED driver on this engine is checked pair by pair (tests/test_gpu_driver.py).
void dump_results(const SeqPair* pairs, int64_t n)
{
    static const char* path = getenv("BSW_SHIM_DUMP");
    if (!path || !*path) return;
    static std::mutex m;
    std::lock_guard<std::mutex> g(m);
    FILE* f = fopen(path, "ab");
    if (!f) return;
    const uint64_t addr = (uint64_t)(uintptr_t)pairs;
    fwrite(&addr, 8, 1, f); fwrite(&n, 8, 1, f);
    for (int64_t i = 0; i < n; ++i) {
        const int32_t v[6] = {pairs[i].score, pairs[i].qle, pairs[i].tle, pairs[i].gtle, pairs[i].gscore, pairs[i].max_off};
        fwrite(v, 4, 6, f);
    }
    fclose(f);
}
// BSW_SHIM_COALESCE=<pairs>: getScores* calls of at most that many pairs, from ALL instances and threads, go through
// one engine per parameter set and its coalescing queue (bsw_extend_async), so that the driver's T threads x 512-pair
// calls (main_banded.cpp:279-291, scripts/run-cpu.sh:30) reach the GPU as batches of up to T x 512 pairs.  OFF by
// default (0): measured from C++ (scripts/latency_probe.cpp, profiles/r02n_latency*.txt) a blocking caller keeps only
// 512 T pairs in flight, and at T <= 32 a coalesced batch's latency (0.55 - 0.75 ms through the bucketing pipeline)
// costs more than T private engines running their 0.3-ms latency route side by side (T = 8: 8.3 M pairs/s private,
// 4.5 M coalesced).  The queue pays for callers written against bsw_extend_async that keep >= 16 k pairs in flight.
// The shared engines live until the process exits.
int64_t coalesce_max()
{
    static const int64_t v = getenv("BSW_SHIM_COALESCE") ? atoll(getenv("BSW_SHIM_COALESCE")) : 0;
    return v;
}
bsw_engine* shared_engine(const bsw_params& p, const char** why)
{
    static std::mutex m;
    static std::vector<std::pair<bsw_params, bsw_engine*>> engines;
    std::lock_guard<std::mutex> g(m);
    for (auto& e : engines)
        if (memcmp(&e.first, &p, sizeof(p)) == 0) return e.second;
    int err = 0;
    bsw_engine* eng = bsw_create(&p, &err);
    if (!eng) { *why = bsw_last_error(nullptr); return nullptr; }
    engines.emplace_back(p, eng);
    return eng;
}
[[noreturn]] void die(const char* what, const char* detail)
{
    fprintf(stderr, "bsw_b200: %s: %s\n", what, detail ? detail : "");
    exit(EXIT_FAILURE);                    // reference behaviour on failure, bandedSWA.cpp:94-99
}
} // namespace

BandedPairWiseSW::BandedPairWiseSW(const int o_del, const int e_del, const int o_ins,
                                   const int e_ins, const int zdrop,
                                   const int end_bonus, const int8_t *mat_,
                                   const int8_t w_match, const int8_t w_mismatch, int numThreads)
    : SW_cells(0), mat(mat_), vec_(nullptr), scalar_(nullptr), ticks_(0)
{
    (void)numThreads;                      // scratch sizing in the reference (:85-92); nothing to size here
    bsw_default_params(&params_);
    params_.o_del = o_del; params_.e_del = e_del;
    params_.o_ins = o_ins; params_.e_ins = e_ins;
    params_.zdrop = zdrop; params_.end_bonus = end_bonus;
    params_.match = w_match;
    params_.mismatch = w_mismatch;         // positive penalty; the reference negates it at :66
    params_.ambig = DEFAULT_AMBIG;         // vector code hard-wires -1 (:69) and ignores `mat`
    params_.tiny_batch = 1536;             // the driver feeds -b 512 pairs per call (scripts/run-cpu.sh:30): latency route
    // experiments (scripts/latency_probe.py): BSW_SHIM_WARP_MAX = bsw_params.warp_max_pairs
    if (const char* e = getenv("BSW_SHIM_WARP_MAX")) params_.warp_max_pairs = atoi(e);
    memset(&stats_, 0, sizeof(stats_));
    // The reference constructs its objects before its timed region (main_banded.cpp:253-258, timing starts at :272):
    // bring the shared engine and its coalescing queue up here, not inside the first getScores16 call.  Best effort:
    // without a device the first call reports the failure, as before.
    if (coalesce_max() > 0) {
        const char* why = "";
        const bsw_params p = checked_params(BSW_ZDROP_VECTOR);
        if (bsw_engine* e = shared_engine(p, &why)) {
            int64_t ticket = 0;
            if (bsw_extend_async(e, nullptr, nullptr, nullptr, 0, 0, &ticket) == BSW_OK) bsw_wait(e, ticket, nullptr);
        }
    }
}

BandedPairWiseSW::~BandedPairWiseSW()
{
    bsw_destroy(vec_);
    bsw_destroy(scalar_);
}

bsw_params BandedPairWiseSW::checked_params(int zdrop_mode)
{
    bsw_params p = params_;
    p.zdrop_mode = zdrop_mode;
    if (zdrop_mode == BSW_ZDROP_SCALAR && mat) {
        // the scalar code scores through `mat` (:148-152): accept the bwa_fill_scmat shape only
        p.ambig = mat[4];
        for (int i = 0; i < 5; ++i)
            for (int j = 0; j < 5; ++j) {
                const int expect = (i == 4 || j == 4) ? p.ambig : (i == j ? p.match : -p.mismatch);
                if (mat[i * 5 + j] != expect)
                    die("scoring matrix", "only matrices of the bwa_fill_scmat form (match / -mismatch / ambig) are supported");
            }
    }
    return p;
}

bsw_engine* BandedPairWiseSW::engine(int zdrop_mode)
{
    bsw_engine*& slot = zdrop_mode == BSW_ZDROP_SCALAR ? scalar_ : vec_;
    if (slot) return slot;
    const bsw_params p = checked_params(zdrop_mode);
    int err = 0;
    slot = bsw_create(&p, &err);
    if (!slot) die("engine creation failed", bsw_last_error(nullptr));
    return slot;
}

void BandedPairWiseSW::run(int zdrop_mode, SeqPair* pairs, const uint8_t* ref, const uint8_t* qer,
                           int64_t n, int32_t w)
{
    const int64_t t0 = ticks_now();
    if (n > 0 && n <= coalesce_max()) {
        bsw_params p = checked_params(zdrop_mode);
        const char* why = "";
        bsw_engine* e = shared_engine(p, &why);
        if (!e) die("engine creation failed", why);
        int64_t ticket = 0, cells = 0;
        if (bsw_extend_async(e, pairs, ref, qer, n, w, &ticket) != BSW_OK) die("bsw_extend_async failed", bsw_last_error(e));
        if (bsw_wait(e, ticket, &cells) != BSW_OK) die("bsw_extend failed", bsw_last_error(e));
        memset(&stats_, 0, sizeof(stats_));
        stats_.pairs = n; stats_.cells_effective = cells;       // (the call's share of its batch)
    } else {
        bsw_engine* e = engine(zdrop_mode);
        if (bsw_extend(e, pairs, ref, qer, n, w) != BSW_OK) die("bsw_extend failed", bsw_last_error(e));
        bsw_get_stats(e, &stats_);
    }
    dump_results(pairs, n);
    SW_cells += (uint64_t)stats_.cells_effective;
    ticks_ += ticks_now() - t0;
}

void BandedPairWiseSW::getScores16(SeqPair *pairArray, uint8_t *seqBufRef, uint8_t *seqBufQer,
                                   int32_t numPairs, uint16_t numThreads, int32_t w)
{
    (void)numThreads;
    run(BSW_ZDROP_VECTOR, pairArray, seqBufRef, seqBufQer, numPairs, w);
}

void BandedPairWiseSW::getScores8(SeqPair *pairArray, uint8_t *seqBufRef, uint8_t *seqBufQer,
                                  int32_t numPairs, uint16_t numThreads, int32_t w)
{
    (void)numThreads;
    run(BSW_ZDROP_VECTOR, pairArray, seqBufRef, seqBufQer, numPairs, w);
}

void BandedPairWiseSW::scalarBandedSWAWrapper(SeqPair *seqPairArray, uint8_t *seqBufRef,
                                              uint8_t *seqBufQer, int numPairs, int nthreads, int32_t w)
{
    (void)nthreads;
    run(BSW_ZDROP_SCALAR, seqPairArray, seqBufRef, seqBufQer, numPairs, w);
}

int BandedPairWiseSW::scalarBandedSWA(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                                      int32_t w, int h0, int *_qle, int *_tle, int *_gtle,
                                      int *_gscore, int *_max_off)
{
    SeqPair sp;
    memset(&sp, 0, sizeof(sp));
    sp.len1 = tlen; sp.len2 = qlen; sp.h0 = h0;
    run(BSW_ZDROP_SCALAR, &sp, target, query, 1, w);
    if (_qle) *_qle = sp.qle;
    if (_tle) *_tle = sp.tle;
    if (_gtle) *_gtle = sp.gtle;
    if (_gscore) *_gscore = sp.gscore;
    if (_max_off) *_max_off = sp.max_off;
    return sp.score;
}

int64_t BandedPairWiseSW::getTicks() { return ticks_; }
