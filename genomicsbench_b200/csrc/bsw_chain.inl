// bsw_chain.inl -- seed -> pair construction and the local / to-end decision of the aligner around the
// extension kernels; included by bsw_engine.cu inside extern "C".
//
// Stands in for the body of mem_chain2aln (tools/bwa/bwamem.c:632-822).  The reference walks the
// seeds of one chain sequentially and calls ksw_extend2 twice per seed; here every round takes the
// next surviving seed of EVERY chain, builds all left-flank pairs (reversed copies) and runs them as
// one bsw_extend_retry batch on the GPU, applies the local / to-end decision, then does the same for
// the right flanks (which read the caller's read and window bytes in place: no copy).  The seeds of
// one chain stay sequential because the containment test (:667-700) looks at the alignments already
// made from that chain.
namespace {

inline int chain_max_gap(const bsw_params& p, int w, int qlen)                  // cal_max_gap, bwamem.c:620-628
{
    const int l_del = (int)((double)(qlen * p.match - p.o_del) / p.e_del + 1.);
    const int l_ins = (int)((double)(qlen * p.match - p.o_ins) / p.e_ins + 1.);
    int l = l_del > l_ins ? l_del : l_ins;
    l = l > 1 ? l : 1;
    return l < (w << 1) ? l : (w << 1);
}

// ksw_extend2 over an empty target (tlen == 0: the row loop never runs, ksw.c:413-473) inside the
// MAX_BAND_TRY loop: score = h0, every end = 0, gscore = -1, max_off = 0
inline void chain_empty_target(int h0, int w, int max_try, int prev0, SeqPair& r, int& band)
{
    int score = prev0;
    for (int t = 0; t < max_try; ++t) {
        const int prev = score;
        band = w << t;
        score = h0;
        if (score == prev || 0 < (band >> 1) + (band >> 2)) break;
    }
    r.score = score; r.qle = 0; r.tle = 0; r.gtle = 0; r.gscore = -1; r.max_off = 0;
}

struct ChainRun {
    uint64_t* srt = nullptr;        // score << 32 | index, ascending (bwamem.c:661-665); a slice of one flat array
    int k = -1;                     // next position in srt, walking down
};

struct ChainBufs {                  // page-locked, grow-only staging of the left / right flanks; owned by the engine
    void* lp = nullptr; void* lq = nullptr; void* lr = nullptr;
    size_t lp_cap = 0, lq_cap = 0, lr_cap = 0;
    void* rpp = nullptr; void* rq = nullptr; void* rr = nullptr;
    size_t rpp_cap = 0, rq_cap = 0, rr_cap = 0;
    static bool grow(void*& p, size_t& cap, size_t need)
    {
        if (need <= cap) return true;
        if (p) bsw_host_free(p);
        cap = need + need / 2;
        p = bsw_host_alloc(cap);
        if (!p) cap = 0;
        return p != nullptr;
    }
    ~ChainBufs() { bsw_host_free(lp); bsw_host_free(lq); bsw_host_free(lr); bsw_host_free(rpp); bsw_host_free(rq); bsw_host_free(rr); }
};

struct ChainCand {                  // one seed being extended in the current round
    int64_t chain; int seed;        // seed index inside the chain
    int lpair = -1, rpair = -1;     // positions in the round's left / right batch, -1 = none
    int64_t lq_off = 0, lr_off = 0; // byte offsets of the reversed left flanks in the staging buffers
    int64_t rq_off = 0, rr_off = 0; // byte offsets of the right flanks' copies (caller's buffers pageable)
    int aw0, aw1;
};

} // namespace

struct ChainLanes;
static void bsw_chain_lanes_release(bsw_engine* eng);

static void bsw_chain_release(bsw_engine* eng)
{
    delete static_cast<ChainBufs*>(eng->cbufs);
    eng->cbufs = nullptr;
    bsw_chain_lanes_release(eng);
}

int bsw_chain_window(const bsw_params* p, int32_t w, int64_t l_pac, const bsw_seed* seeds, int32_t n,
                     int32_t l_query, int64_t* rmax0, int64_t* rmax1)
{
    if (!p || !seeds || n < 1 || !rmax0 || !rmax1 || p->e_del < 1 || p->e_ins < 1) return BSW_ERR_PARAM;
    int64_t r0 = l_pac << 1, r1 = 0;
    for (int i = 0; i < n; ++i) {                                               // bwamem.c:643-652
        const bsw_seed& t = seeds[i];
        const int64_t b = t.rbeg - (t.qbeg + chain_max_gap(*p, w, t.qbeg));
        const int64_t e = t.rbeg + t.len + ((l_query - t.qbeg - t.len) + chain_max_gap(*p, w, l_query - t.qbeg - t.len));
        r0 = r0 < b ? r0 : b;
        r1 = r1 > e ? r1 : e;
    }
    r0 = r0 > 0 ? r0 : 0;
    r1 = r1 < (l_pac << 1) ? r1 : (l_pac << 1);
    if (r0 < l_pac && l_pac < r1) {                                             // :655-658: stay on the seeds' strand
        if (seeds[0].rbeg < l_pac) r1 = l_pac;
        else r0 = l_pac;
    }
    *rmax0 = r0; *rmax1 = r1;
    return BSW_OK;
}

static int extend_chains_range(bsw_engine* eng, const bsw_chain* chains, int64_t n_chains, const bsw_seed* seeds,
                               const uint8_t* query, const uint8_t* ref, const bsw_chain_opt* opt,
                               bsw_alnreg* out, int32_t* out_count);

// Reads are independent, so a large batch is cut at read boundaries into a few lanes that run side by side, each on
// its own host thread and child engine: while one lane's extensions are on the GPU another lane picks seeds and builds
// flanks (a round's picks depend on the previous round's results of the SAME read only).  One lane alone alternates
// between host and device: 52 ms per 200 k chains, half of it with the GPU idle (BSW_TIMELINE).
struct ChainLanes {
    std::vector<bsw_engine*> child;
    ~ChainLanes() { for (bsw_engine* e : child) bsw_destroy(e); }
};

int bsw_extend_chains(bsw_engine* eng, const bsw_chain* chains, int64_t n_chains, const bsw_seed* seeds,
                      const uint8_t* query, const uint8_t* ref, const bsw_chain_opt* opt,
                      bsw_alnreg* out, int32_t* out_count)
{
    if (!eng) return BSW_ERR_PARAM;
    static const int env_lanes = getenv("BSW_CHAIN_LANES") ? atoi(getenv("BSW_CHAIN_LANES")) : 0;
    // (measured, 200 k chains: 1 lane 3.6 M chains/s, 2 lanes 5.4 M, 4 lanes 7.5 M, 8 lanes 8.0 M)
    int lanes = env_lanes > 0 ? env_lanes : (n_chains >= 120000 ? 8 : n_chains >= 60000 ? 4 : n_chains >= 20000 ? 2 : 1);
    lanes = std::min(lanes, 8);
    if (lanes <= 1 || !chains) return extend_chains_range(eng, chains, n_chains, seeds, query, ref, opt, out, out_count);
    // cut points at read boundaries (a chain with same_read set belongs to its predecessor's read)
    std::vector<int64_t> cut{0};
    for (int g = 1; g < lanes; ++g) {
        int64_t c = n_chains * g / lanes;
        while (c < n_chains && c > 0 && chains[c].same_read) ++c;
        if (c > cut.back() && c < n_chains) cut.push_back(c);
    }
    cut.push_back(n_chains);
    lanes = (int)cut.size() - 1;
    if (lanes <= 1) return extend_chains_range(eng, chains, n_chains, seeds, query, ref, opt, out, out_count);
    if (!eng->clanes) eng->clanes = new ChainLanes();
    ChainLanes& L = *static_cast<ChainLanes*>(eng->clanes);
    while ((int)L.child.size() < lanes) {
        bsw_params p = eng->p;
        p.host_threads = std::max(2, eng->pool->size() / lanes);
        int err = 0;
        bsw_engine* e = bsw_create(&p, &err);
        if (!e) { eng->err = "bsw_extend_chains: cannot create a lane engine"; return BSW_ERR_CUDA; }
        L.child.push_back(e);
    }
    std::vector<int> rcs((size_t)lanes, BSW_OK);
    std::vector<std::thread> th;
    auto work = [&](int g) {
        const int64_t a = cut[(size_t)g], m = cut[(size_t)g + 1] - a;
        rcs[(size_t)g] = extend_chains_range(L.child[(size_t)g], chains + a, m, seeds, query, ref, opt, out, out_count + a);
    };
    for (int g = 1; g < lanes; ++g) th.emplace_back(work, g);
    work(0);
    for (std::thread& t : th) t.join();
    bsw_stats total;
    memset(&total, 0, sizeof(total));
    int rc = BSW_OK;
    for (int g = 0; g < lanes; ++g) {
        const bsw_engine* e = L.child[(size_t)g];
        if (rcs[(size_t)g] != BSW_OK && rc == BSW_OK) { rc = rcs[(size_t)g]; eng->err = e->err; }
        const bsw_stats& s = e->stats;
        total.pairs += s.pairs; total.cells_nominal += s.cells_nominal; total.cells_effective += s.cells_effective;
        total.kernel_launches += s.kernel_launches; total.h2d_bytes += s.h2d_bytes; total.d2h_bytes += s.d2h_bytes;
        total.ms_kernel = std::max(total.ms_kernel, s.ms_kernel); total.ms_pack += s.ms_pack; total.ms_scatter += s.ms_scatter;
        total.n_short += s.n_short; total.n_long += s.n_long;
    }
    total.shards = lanes;
    eng->stats = total;
    return rc;
}

static int extend_chains_range(bsw_engine* eng, const bsw_chain* chains, int64_t n_chains, const bsw_seed* seeds,
                               const uint8_t* query, const uint8_t* ref, const bsw_chain_opt* opt,
                               bsw_alnreg* out, int32_t* out_count)
{
    eng->err.clear();
    if (n_chains < 0 || !opt || (n_chains > 0 && (!chains || !seeds || !query || !ref || !out || !out_count))) {
        eng->err = "bsw_extend_chains: bad arguments";
        return BSW_ERR_PARAM;
    }
    const int w = opt->w, max_try = opt->max_band_try;
    if (w < 0 || max_try < 1 || max_try > 8) { eng->err = "bsw_extend_chains: w >= 0 and max_band_try in 1..8"; return BSW_ERR_PARAM; }
    if (opt->pen_clip5 != eng->p.end_bonus || opt->pen_clip3 != eng->p.end_bonus) {
        eng->err = "bsw_extend_chains: pen_clip5 and pen_clip3 must equal the engine's end_bonus "
                   "(ksw_extend2 receives them as end_bonus, bwamem.c:746,793)";
        return BSW_ERR_PARAM;
    }
    const bsw_params& P = eng->p;
    std::vector<ChainRun> run((size_t)n_chains);
    std::vector<int64_t> groups;                                // first chain of every read (chains of a read are adjacent)
    int64_t n_seeds_total = 0, seed_lo = INT64_MAX;
    for (int64_t c = 0; c < n_chains; ++c) {
        const bsw_chain& ch = chains[c];
        out_count[c] = 0;
        if (ch.n_seeds < 0 || ch.l_query < 1 || ch.rmax1 < ch.rmax0 || ch.seed_first < 0) { eng->err = "bsw_extend_chains: malformed chain"; return BSW_ERR_PARAM; }
        if (!(c > 0 && ch.same_read)) groups.push_back(c);
        if (ch.same_read && (c == 0 || chains[c - 1].l_query != ch.l_query || chains[c - 1].query_off != ch.query_off)) {
            eng->err = "bsw_extend_chains: same_read set on a chain whose predecessor is another read";
            return BSW_ERR_PARAM;
        }
        n_seeds_total = std::max(n_seeds_total, ch.seed_first + ch.n_seeds);
        seed_lo = std::min(seed_lo, ch.seed_first);
    }
    // every chain sorts its own slice [seed_first, + n_seeds); only the span this range of chains uses is allocated
    // (a lane of a large batch sees a fraction of the seeds)
    if (n_chains == 0) seed_lo = 0;
    std::unique_ptr<uint64_t[]> srt_all(new uint64_t[(size_t)std::max<int64_t>(n_seeds_total - seed_lo, 1)]);
    std::atomic<int> bad_seed{0};
    eng->pool->for_range(n_chains, 1024, [&](int64_t cb, int64_t ce, int) {
        for (int64_t c = cb; c < ce; ++c) {
            const bsw_chain& ch = chains[c];
            ChainRun& R = run[(size_t)c];
            R.srt = srt_all.get() + (ch.seed_first - seed_lo);
            for (int i = 0; i < ch.n_seeds; ++i) {
                const bsw_seed& s = seeds[ch.seed_first + i];
                if (s.len < 1 || s.qbeg < 0 || s.qbeg + s.len > ch.l_query || s.score < 1 || s.rbeg < ch.rmax0 ||
                    s.rbeg + s.len > ch.rmax1) bad_seed.store(1, std::memory_order_relaxed);
                R.srt[i] = (uint64_t)(uint32_t)s.score << 32 | (uint32_t)i;
            }
            std::sort(R.srt, R.srt + ch.n_seeds);
            R.k = ch.n_seeds - 1;
        }
    });
    if (bad_seed.load()) { eng->err = "bsw_extend_chains: seed outside its read or reference window"; return BSW_ERR_PARAM; }

    bsw_stats total;
    memset(&total, 0, sizeof(total));
    std::vector<ChainCand> cand;
    if (!eng->cbufs) eng->cbufs = new ChainBufs();
    ChainBufs& CB = *static_cast<ChainBufs*>(eng->cbufs);
    std::vector<int32_t> band, prev, pick;
    auto add_stats = [&]() {
        const bsw_stats& s = eng->stats;
        total.pairs += s.pairs; total.cells_nominal += s.cells_nominal; total.cells_effective += s.cells_effective;
        total.kernel_launches += s.kernel_launches; total.h2d_bytes += s.h2d_bytes; total.d2h_bytes += s.d2h_bytes;
        total.ms_kernel += s.ms_kernel; total.ms_pack += s.ms_pack; total.ms_scatter += s.ms_scatter;
        total.n_short += s.n_short; total.n_long += s.n_long;
    };

    double tl[6] = {0, 0, 0, 0, 0, 0};          // BSW_TIMELINE: pick / left build / left GPU / right build / right GPU / finish
    int rounds = 0;
    double t_mark = now_ms();
    auto lap = [&](int k) { const double t = now_ms(); tl[k] += t - t_mark; t_mark = t; };
    for (;;) {
        ++rounds;
        // ---- next surviving seed of every read (containment test, bwamem.c:667-700) -------------
        // The chains of a read run one after the other (mem_align1_core pushes their regions into one vector,
        // bwamem.c:1105-1112): per read, the first chain that still has seeds is searched; if the search only
        // skips contained seeds and drains the chain, the read's next chain is searched in the same round, so
        // a round without a candidate means every chain of the batch is done (the result never depends on
        // which other reads share the batch).
        cand.clear();
        pick.assign((size_t)n_chains, -1);
        eng->pool->for_range((int64_t)groups.size(), 256, [&](int64_t gb, int64_t ge, int) {
        for (int64_t gi = gb; gi < ge; ++gi) {
        const int64_t g0 = groups[(size_t)gi], g1 = gi + 1 < (int64_t)groups.size() ? groups[(size_t)gi + 1] : n_chains;
        for (int64_t c = g0; c < g1; ++c) {
            const bsw_chain& ch = chains[c];
            ChainRun& R = run[(size_t)c];
            if (R.k < 0) continue;                                              // done: the read's next chain acts
            const bsw_seed* S = seeds + ch.seed_first;
            // regions the containment test looks at: those of the read's earlier chains, then this chain's
            int n_av = 0;
            for (int64_t g = g0; g <= c; ++g) n_av += out_count[g];
            auto reg_at = [&](int i) -> const bsw_alnreg& {
                for (int64_t g = g0;; ++g) {
                    if (i < out_count[g]) return out[chains[g].seed_first + i];
                    i -= out_count[g];
                }
            };
            while (R.k >= 0) {
                const int k = R.k;
                const bsw_seed& s = S[(uint32_t)R.srt[k]];
                int i;
                for (i = 0; i < n_av; ++i) {
                    const bsw_alnreg& p = reg_at(i);
                    if (s.rbeg < p.rb || s.rbeg + s.len > p.re || s.qbeg < p.qb || s.qbeg + s.len > p.qe) continue;
                    if (s.len - p.seedlen0 > .1 * ch.l_query) continue;
                    int qd = s.qbeg - p.qb; int64_t rd = s.rbeg - p.rb;
                    int max_gap = chain_max_gap(P, w, qd < rd ? qd : (int)rd);
                    int ww = max_gap < p.w ? max_gap : p.w;
                    if (qd - rd < ww && rd - qd < ww) break;
                    qd = p.qe - (s.qbeg + s.len); rd = p.re - (s.rbeg + s.len);
                    max_gap = chain_max_gap(P, w, qd < rd ? qd : (int)rd);
                    ww = max_gap < p.w ? max_gap : p.w;
                    if (qd - rd < ww && rd - qd < ww) break;
                }
                if (i < n_av) {
                    for (i = k + 1; i < ch.n_seeds; ++i) {
                        if (R.srt[i] == 0) continue;
                        const bsw_seed& t = S[(uint32_t)R.srt[i]];
                        if (t.len < s.len * .95) continue;
                        if (s.qbeg <= t.qbeg && s.qbeg + s.len - t.qbeg >= s.len >> 2 && t.qbeg - s.qbeg != t.rbeg - s.rbeg) break;
                        if (t.qbeg <= s.qbeg && t.qbeg + t.len - s.qbeg >= s.len >> 2 && s.qbeg - t.qbeg != s.rbeg - t.rbeg) break;
                    }
                    if (i == ch.n_seeds) { R.srt[k] = 0; --R.k; continue; }     // contained: no extension
                }
                pick[(size_t)c] = (int)(uint32_t)R.srt[k];
                break;
            }
            if (pick[(size_t)c] >= 0) break;                                    // the read's later chains wait for this one
        }
        }
        });
        for (int64_t c = 0; c < n_chains; ++c) {
            if (pick[(size_t)c] < 0) continue;
            ChainCand cd;
            cd.chain = c; cd.seed = pick[(size_t)c]; cd.aw0 = cd.aw1 = w;
            cand.push_back(cd);
        }
        if (cand.empty()) break;
        lap(0);

        // ---- left flanks: reversed query prefix and reference flank, h0 = len * a (:709-753) -----
        // Offsets first (a running sum over the candidates), then pairs, region headers and the reversed
        // copies on the thread pool, straight into page-locked buffers: the extension call then takes the
        // engine's direct route (DMA of records and bytes as they are, no second host pass).
        size_t qbytes = 0, rbytes = 0, n_left = 0;
        for (ChainCand& cd : cand) {
            const bsw_chain& ch = chains[cd.chain];
            const bsw_seed& s = seeds[ch.seed_first + cd.seed];
            const int64_t tmp = s.rbeg - ch.rmax0;
            cd.lpair = -1;
            if (s.qbeg > 0 && tmp > 0) {
                if (tmp > 32767) { eng->err = "bsw_extend_chains: left reference flank longer than 32767"; return BSW_ERR_DOMAIN; }
                cd.lpair = (int)n_left++;
                cd.lq_off = (int64_t)qbytes; cd.lr_off = (int64_t)rbytes;
                qbytes += (size_t)s.qbeg; rbytes += (size_t)tmp;
            }
        }
        if (!CB.grow(CB.lp, CB.lp_cap, (n_left + 16) * sizeof(SeqPair)) || !CB.grow(CB.lq, CB.lq_cap, qbytes + 64) ||
            !CB.grow(CB.lr, CB.lr_cap, rbytes + 64)) {
            eng->err = "bsw_extend_chains: page-locked staging allocation failed";
            return BSW_ERR_NOMEM;
        }
        SeqPair* lp = static_cast<SeqPair*>(CB.lp);
        uint8_t* lq = static_cast<uint8_t*>(CB.lq); uint8_t* lr = static_cast<uint8_t*>(CB.lr);
        eng->pool->for_range((int64_t)cand.size(), 256, [&](int64_t b, int64_t e, int) {
            for (int64_t x = b; x < e; ++x) {
                const ChainCand& cd = cand[(size_t)x];
                const bsw_chain& ch = chains[cd.chain];
                const bsw_seed& s = seeds[ch.seed_first + cd.seed];
                bsw_alnreg& a = out[ch.seed_first + out_count[cd.chain]];
                memset(&a, 0, sizeof(a));
                a.w = w; a.score = a.truesc = -1;
                if (cd.lpair < 0) continue;
                const int64_t tmp = s.rbeg - ch.rmax0;
                SeqPair& sp = lp[cd.lpair];
                memset(&sp, 0, sizeof(sp));
                sp.idq = cd.lq_off; sp.idr = cd.lr_off; sp.id = cd.lpair;
                sp.len2 = s.qbeg; sp.len1 = (int32_t)tmp; sp.h0 = s.len * P.match;
                const uint8_t* q = query + ch.query_off;
                const uint8_t* r = ref + ch.ref_off;
                uint8_t* dq = lq + cd.lq_off; uint8_t* dr = lr + cd.lr_off;
                for (int i = 0; i < s.qbeg; ++i) dq[i] = q[s.qbeg - 1 - i];
                for (int64_t i = 0; i < tmp; ++i) dr[i] = r[tmp - 1 - i];
            }
        });
        if (n_left > 0) {
            band.assign(n_left, w);
            lap(1);
            if (int rc = bsw_extend_retry(eng, lp, lr, lq, (int64_t)n_left, w, max_try, nullptr, band.data()))
                return rc;
            add_stats();
            lap(2);
        }
        eng->pool->for_range((int64_t)cand.size(), 1024, [&](int64_t xb, int64_t xe, int) {
        for (int64_t x = xb; x < xe; ++x) {
            ChainCand& cd = cand[(size_t)x];
            const bsw_chain& ch = chains[cd.chain];
            const bsw_seed& s = seeds[ch.seed_first + cd.seed];
            bsw_alnreg& a = out[ch.seed_first + out_count[cd.chain]];
            if (s.qbeg > 0) {
                SeqPair r;
                if (cd.lpair >= 0) { r = lp[(size_t)cd.lpair]; cd.aw0 = band[(size_t)cd.lpair]; }
                else chain_empty_target(s.len * P.match, w, max_try, -1, r, cd.aw0);
                a.score = r.score;
                if (r.gscore <= 0 || r.gscore <= a.score - opt->pen_clip5) {    // local extension (:755-758)
                    a.qb = s.qbeg - r.qle; a.rb = s.rbeg - r.tle;
                    a.truesc = a.score;
                } else {                                                        // to-end extension (:759-761)
                    a.qb = 0; a.rb = s.rbeg - r.gtle;
                    a.truesc = r.gscore;
                }
            } else { a.score = a.truesc = s.len * P.match; a.qb = 0; a.rb = s.rbeg; }       // :763
        }
        });

        // ---- right flanks, h0 = the left score (:765-800): read in place when the caller's read / window buffers
        // are page-locked, else copied into page-locked staging like the left flanks -- either way the extension call
        // takes the engine's direct route (the staged route's host passes cost twice the left flanks' GPU phase)
        const bool in_place = is_pinned(query) && is_pinned(ref);
        size_t n_right = 0, rqb = 0, rrb = 0;
        for (ChainCand& cd : cand) {
            const bsw_chain& ch = chains[cd.chain];
            const bsw_seed& s = seeds[ch.seed_first + cd.seed];
            cd.rpair = -1;
            if (s.qbeg + s.len == ch.l_query) continue;
            const int64_t tlen = ch.rmax1 - ch.rmax0 - (s.rbeg + s.len - ch.rmax0);
            if (tlen <= 0) continue;
            if (tlen > 32767) { eng->err = "bsw_extend_chains: right reference flank longer than 32767"; return BSW_ERR_DOMAIN; }
            cd.rpair = (int)n_right++;
            cd.rq_off = (int64_t)rqb; cd.rr_off = (int64_t)rrb;
            rqb += (size_t)(ch.l_query - (s.qbeg + s.len)); rrb += (size_t)tlen;
        }
        prev.resize(n_right);
        if (!CB.grow(CB.rpp, CB.rpp_cap, (n_right + 16) * sizeof(SeqPair)) ||
            (!in_place && (!CB.grow(CB.rq, CB.rq_cap, rqb + 64) || !CB.grow(CB.rr, CB.rr_cap, rrb + 64)))) {
            eng->err = "bsw_extend_chains: page-locked staging allocation failed";
            return BSW_ERR_NOMEM;
        }
        SeqPair* const rp = static_cast<SeqPair*>(CB.rpp);
        uint8_t* const rq = static_cast<uint8_t*>(CB.rq); uint8_t* const rr = static_cast<uint8_t*>(CB.rr);
        eng->pool->for_range((int64_t)cand.size(), 1024, [&](int64_t xb, int64_t xe, int) {
            for (int64_t x = xb; x < xe; ++x) {
                const ChainCand& cd = cand[(size_t)x];
                if (cd.rpair < 0) continue;
                const bsw_chain& ch = chains[cd.chain];
                const bsw_seed& s = seeds[ch.seed_first + cd.seed];
                const bsw_alnreg& a = out[ch.seed_first + out_count[cd.chain]];
                const int qe = s.qbeg + s.len;
                const int64_t re = s.rbeg + s.len - ch.rmax0;
                SeqPair& sp = rp[(size_t)cd.rpair];
                memset(&sp, 0, sizeof(sp));
                sp.id = cd.rpair;
                sp.len2 = ch.l_query - qe; sp.len1 = (int32_t)(ch.rmax1 - ch.rmax0 - re); sp.h0 = a.score;
                if (in_place) { sp.idq = ch.query_off + qe; sp.idr = ch.ref_off + re; }
                else {
                    sp.idq = cd.rq_off; sp.idr = cd.rr_off;
                    memcpy(rq + cd.rq_off, query + ch.query_off + qe, (size_t)sp.len2);
                    memcpy(rr + cd.rr_off, ref + ch.ref_off + re, (size_t)sp.len1);
                }
                prev[(size_t)cd.rpair] = a.score;
            }
        });
        if (n_right > 0) {
            band.assign(n_right, w);
            lap(3);
            if (int rc = bsw_extend_retry(eng, rp, in_place ? ref : rr, in_place ? query : rq, (int64_t)n_right, w, max_try,
                                          prev.data(), band.data()))
                return rc;
            add_stats();
            lap(4);
        }
        eng->pool->for_range((int64_t)cand.size(), 1024, [&](int64_t xb, int64_t xe, int) {
        for (int64_t x = xb; x < xe; ++x) {
            ChainCand& cd = cand[(size_t)x];
            const bsw_chain& ch = chains[cd.chain];
            const bsw_seed* S = seeds + ch.seed_first;
            const bsw_seed& s = S[cd.seed];
            bsw_alnreg& a = out[ch.seed_first + out_count[cd.chain]];
            if (s.qbeg + s.len != ch.l_query) {
                const int qe = s.qbeg + s.len, sc0 = a.score;
                const int64_t re = s.rbeg + s.len - ch.rmax0;
                SeqPair r;
                if (cd.rpair >= 0) { r = rp[(size_t)cd.rpair]; cd.aw1 = band[(size_t)cd.rpair]; }
                else chain_empty_target(sc0, w, max_try, sc0, r, cd.aw1);
                a.score = r.score;
                if (r.gscore <= 0 || r.gscore <= a.score - opt->pen_clip3) {    // :802-805
                    a.qe = qe + r.qle; a.re = ch.rmax0 + re + r.tle;
                    a.truesc += a.score - sc0;
                } else {                                                        // :806-808
                    a.qe = ch.l_query; a.re = ch.rmax0 + re + r.gtle;
                    a.truesc += r.gscore - sc0;
                }
            } else { a.qe = ch.l_query; a.re = s.rbeg + s.len; }                // :810
            a.seedcov = 0;                                                      // :812-818
            for (int i = 0; i < ch.n_seeds; ++i) {
                const bsw_seed& t = S[i];
                if (t.qbeg >= a.qb && t.qbeg + t.len <= a.qe && t.rbeg >= a.rb && t.rbeg + t.len <= a.re) a.seedcov += t.len;
            }
            a.w = cd.aw0 > cd.aw1 ? cd.aw0 : cd.aw1;
            a.seedlen0 = s.len;
            ++out_count[cd.chain];
            --run[(size_t)cd.chain].k;
        }
        });
    }
    lap(5);
    if (g_timeline)
        fprintf(stderr, "bsw_extend_chains: %d rounds; ms: pick %.2f, left build %.2f, left GPU %.2f, right build + decisions %.2f, "
                        "right GPU %.2f, finish %.2f\n", rounds, tl[0], tl[1], tl[2], tl[3], tl[4], tl[5]);
    eng->stats = total;
    return BSW_OK;
}

static void bsw_chain_lanes_release(bsw_engine* eng)
{
    delete static_cast<ChainLanes*>(eng->clanes);
    eng->clanes = nullptr;
}
