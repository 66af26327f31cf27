// g2_emu.cu -- TEST INFRASTRUCTURE: bsw_global's second kernel on the CPU.  The chunk planner
// (csrc/bsw_global_plan.h) and the per-thread alignment (g2::align_one, csrc/bsw_global2.cuh) are the
// very sources the engine and the GPU run; what is emulated is the device around them -- shared memory as
// a garbage-filled host buffer with the block's interleaving, the handful of instructions the cell uses
// (PRMT, IDP, VIADDMNMX, funnel shifts: the #else branches in bsw_global2.cuh), the launch loop and the
// compaction of the operation lists.  tests/test_g2_emulation.py compares the outcome with the goldens
// the reference's own ksw_global2 produced.  Built by tests/emu/build.py with nvcc as host code.
#include <vector>
#include <cstdio>
#include <algorithm>
#include "../../include/bsw.h"
#include "../../genomicsbench_b200/csrc/bsw_global_plan.h"

using namespace bsw;

namespace {
struct SerialPar {
    template <class F> void operator()(int64_t n, int64_t grain, F&& fn) const
    {
        // ranges in reverse order: nothing a pass writes may depend on the schedule
        const int64_t nch = (n + grain - 1) / grain;
        for (int64_t c = nch - 1; c >= 0; --c) fn(c * grain, std::min(n, (c + 1) * grain), 0);
    }
};

template <bool R16>
void run_thread(const g2::Params& P, const GlobalDesc& d, const uint8_t* hq, const uint8_t* hr, int slots, int qwords, int tid,
                std::vector<uint32_t>& smem, uint8_t* z, uint32_t* cig, int32_t* score, int32_t* ncig)
{
    constexpr int WORDS = g2::Slots<R16>::WORDS;
    smem.assign(g2::smem_bytes(R16, slots, qwords) / 4, 0xA5A5A5A5u);
    uint32_t* rows = smem.data() + tid * WORDS;
    uint32_t* qs = smem.data() + (size_t)slots * g2::BLOCK * WORDS + tid;
    g2::pack_query<g2::BLOCK>(hq + d.qoff, d.qlen, qs);
    int32_t sc, nc;
    g2::align_one<R16, g2::BLOCK, g2::BLOCK>(P, d.qlen, d.tlen, d.w, qs, hr + d.roff, rows, reinterpret_cast<uint32_t*>(z + d.zoff),
                                             cig + d.coff, sc, nc);
    score[d.idx] = sc; ncig[d.idx] = nc;
    // the thread stayed inside its class's slots and query words: everything of OTHER threads is still garbage
    for (size_t k = 0; k < smem.size(); ++k) {
        const size_t lane = k < (size_t)slots * g2::BLOCK * WORDS ? (k / WORDS) % g2::BLOCK : k % g2::BLOCK;
        if ((int)lane != tid && smem[k] != 0xA5A5A5A5u) { fprintf(stderr, "g2_emu: thread %d wrote word %zu of lane %zu\n", tid, k, lane); abort(); }
    }
}
} // namespace

extern "C" {

// prm: o_del e_del o_ins e_ins match mismatch(+) ambig.  rows: 16 / 32 = force that slot width, 0 = what the engine would pick.
// caps_m / caps_z: chunk caps (0 = defaults) so that tests can force several chunks.  Returns 0; -1 = outside the second
// kernel's domain (the engine would run the first kernel); -3 = cigar_cap too small.  info[0] = chunks, [1] = launches,
// [2] = 1 when 16-bit slots ran, [3] = effective cells, [4] = values that left 16 bits in a 16-bit slot (must be 0)
int g2_emu_global(const int* prm, const SeqPair* pairs, const uint8_t* seq_ref, const uint8_t* seq_qer, long long n, const int* w,
                  int rows, long long caps_m, long long caps_z, int* score, int* n_cigar, uint32_t* cigar, long long cigar_cap,
                  long long* cigar_off, long long* info)
{
    g2::Params P{};
    P.o_del = prm[0]; P.e_del = prm[1]; P.o_ins = prm[2]; P.e_ins = prm[3];
    const int match = prm[4], mm = -prm[5], ambig = prm[6];
    if (!g2::scores_ok(match, mm, ambig)) return -1;
    g2::fill_table(P, match, mm, ambig);
    bool all16 = true;
    for (long long i = 0; i < n; ++i) all16 = all16 && g2::rows16_ok(P, match, mm, ambig, pairs[i].len2, pairs[i].len1, w[i]);
    const bool r16 = rows == 16 ? true : rows == 32 ? false : all16;
    if (rows == 16 && !all16) return -1;
    for (long long i = 0; i < n; ++i)
        if (g2::smem_bytes(r16, g2::slots_class(g2::row_slots(pairs[i].len2, g2::eff_w(pairs[i].len2, pairs[i].len1, w[i]))), g2::qwords_class(g2::query_words(pairs[i].len2))) > 200 * 1024) return -1;
    g2::Caps caps;
    if (caps_m) caps.m = caps_m;
    if (caps_z) caps.z_bytes = caps_z;
    SerialPar par;
    g2::ChunkPlan pl;
    long long out_pos = 0, chunks = 0, launches = 0, cells = 0;
    g2::emu_wraps = 0;
    cigar_off[0] = 0;
    std::vector<uint32_t> smem;
    for (long long first = 0; first < n;) {
        g2::plan_sizes(pairs, w, first, n, caps, par, pl);
        std::vector<uint8_t> hq((size_t)pl.q_bytes + 16, 0xEE), hr((size_t)pl.r_bytes + 16, 0xEE), z((size_t)pl.z_bytes + 16, 0xEE);
        std::vector<GlobalDesc> desc((size_t)pl.m);
        std::vector<uint32_t> cig((size_t)pl.cig_words + 16, 0xEEEEEEEEu);
        std::vector<int32_t> sc((size_t)pl.m, -12345), nc((size_t)pl.m, -1);
        g2::plan_fill(pairs, w, seq_ref, seq_qer, 3, par, pl, desc.data(), hq.data(), hr.data());
        long long covered = 0;
        for (const g2::Launch& L : pl.launches) {
            if (L.first != covered || L.count <= 0) return -10;
            covered += L.count;
            for (int t = 0; t < L.count; ++t) {
                const GlobalDesc& d = desc[(size_t)L.first + t];
                if (g2::row_slots(d.qlen, d.w) > L.slots || g2::query_words(d.qlen) > L.qwords) return -11;
                if (r16) run_thread<true>(P, d, hq.data(), hr.data(), L.slots, L.qwords, t % g2::BLOCK, smem, z.data(), cig.data(), sc.data(), nc.data());
                else run_thread<false>(P, d, hq.data(), hr.data(), L.slots, L.qwords, t % g2::BLOCK, smem, z.data(), cig.data(), sc.data(), nc.data());
            }
            ++launches;
        }
        if (covered != pl.m) return -12;
        // compaction: input order
        std::vector<long long> off((size_t)pl.m + 1, 0);
        for (long long k = 0; k < pl.m; ++k) off[(size_t)k + 1] = off[(size_t)k] + nc[(size_t)k];
        if (out_pos + off[(size_t)pl.m] > cigar_cap) return -3;
        for (long long t = 0; t < pl.m; ++t) {
            const GlobalDesc& d = desc[(size_t)t];
            for (int k = 0; k < nc[(size_t)d.idx]; ++k) cigar[out_pos + off[(size_t)d.idx] + k] = cig[(size_t)d.coff + k];
        }
        for (long long k = 0; k < pl.m; ++k) {
            score[first + k] = sc[(size_t)k]; n_cigar[first + k] = nc[(size_t)k];
            cigar_off[first + k + 1] = out_pos + off[(size_t)k + 1];
        }
        out_pos += off[(size_t)pl.m];
        cells += pl.cells;
        first += pl.m; ++chunks;
    }
    if (info) { info[0] = chunks; info[1] = launches; info[2] = r16; info[3] = cells; info[4] = g2::emu_wraps; }
    return 0;
}

// the planner's radix sort on n seeded keys with `slices` slices against std::stable_sort on the same bits; 0 = equal
int g2_emu_sort_check(long long n, int slices, unsigned long long seed)
{
    std::vector<uint64_t> keys((size_t)n), tmp, ref;
    std::vector<uint32_t> hist;
    uint64_t x = seed | 1;
    for (auto& k : keys) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; k = (x & ~0x3ffffull) | (uint64_t)((&k - keys.data()) & 0x3ffff); }
    for (size_t k = 0; k < keys.size(); k += 3) keys[k] = (keys[k / 2] & ~0x3ffffull) | (keys[k] & 0x3ffff);      // many equal keys: stability shows
    ref = keys;
    std::stable_sort(ref.begin(), ref.end(), [](uint64_t a, uint64_t b) { return (a >> 18) < (b >> 18); });
    SerialPar par;
    g2::radix_sort(keys, tmp, hist, 18, 64, slices, par);
    return keys == ref ? 0 : 1;
}

} // extern "C"
