// bsw_global.inl -- host side of bsw_global (banded global alignment + CIGAR);
// included by bsw_engine.cu inside extern "C".  Chunks of alignments go to the engine's first device:
// gather the two byte strings of every alignment, one thread per alignment computes score and
// operation list, a second kernel packs the lists, and they land in the caller's buffer one after the
// other (cigar_off[]).
// Two forms.  run_global2 is the default: the chunk planner of bsw_global_plan.h on the thread pool (sizes,
// descriptors + gather, parallel radix sort into work order, one launch per shared-memory class), the kernel of
// bsw_global2.cuh, class launches side by side on the device's DP streams.  run_global1 is the first form
// (bsw_global.cuh: 32-bit rows as a circular buffer or in HBM, byte directions); it takes what the second kernel's
// domain excludes -- scores beyond a signed byte, rows beyond shared memory -- and BSW_GLOBAL_KERNEL=1 selects it
// for A/B runs.
namespace {

// One chunk in flight: grow-only device buffers with page-locked twins for everything that crosses PCIe.  Two of
// them alternate (owned by the engine), so that the host prepares chunk k + 1 -- gather, work order -- while the
// device runs chunk k and chunk k - 1 drains.
struct GlobalSlot {
    Buf<uint8_t> q, r, z;
    Buf<GlobalDesc> desc;
    Buf<int2> eh;
    Buf<uint32_t> cig, packed;
    Buf<int32_t> score, ncig;
    Buf<long long> off;
    cudaStream_t st{};
    cudaEvent_t e0{}, e1{};
    int64_t first = 0, m = 0;            // alignments [first, first + m) of the call
    long long cap_words = 0;             // upper bound of the chunk's packed operations
    long long cells = 0;
};

constexpr size_t G2_SMEM_MAX = 200 * 1024;       // dynamic shared memory of a block of the second kernel

struct GlobalBufs {
    GlobalSlot slot[2];
    g2::ChunkPlan plan[2];
    std::vector<GlobalDesc> hd;
    std::vector<uint64_t> key, tmp;
};

// stable LSD radix sort on bits 20 .. 51 of the keys (two 16-bit digits: target length, band): the chunk's work order
void radix_sort_work(std::vector<uint64_t>& a, std::vector<uint64_t>& tmp)
{
    tmp.resize(a.size());
    std::vector<uint32_t> cnt(65536);
    for (int pass = 0; pass < 2; ++pass) {
        const int sh = 20 + 16 * pass;
        std::fill(cnt.begin(), cnt.end(), 0u);
        for (uint64_t v : a) ++cnt[(v >> sh) & 0xffff];
        uint32_t run = 0;
        for (uint32_t& c : cnt) { const uint32_t t = c; c = run; run += t; }
        for (uint64_t v : a) tmp[cnt[(v >> sh) & 0xffff]++] = v;
        a.swap(tmp);
    }
}

} // namespace

static void bsw_global_release(bsw_engine* eng)
{
    if (!eng->gbufs) return;
    GlobalBufs* B = static_cast<GlobalBufs*>(eng->gbufs);
    if (!eng->devs.empty()) cudaSetDevice(eng->devs[0].dev);
    for (GlobalSlot& S : B->slot) {
        release(S.q); release(S.r); release(S.z); release(S.desc); release(S.eh); release(S.cig);
        release(S.packed); release(S.score); release(S.ncig); release(S.off);
        if (S.st) cudaStreamDestroy(S.st);
        if (S.e0) cudaEventDestroy(S.e0);
        if (S.e1) cudaEventDestroy(S.e1);
    }
    delete B;
    eng->gbufs = nullptr;
}

// first form: arguments validated by bsw_global, statistics zeroed, n > 0
static int run_global1(bsw_engine* eng, const SeqPair* pairs, const uint8_t* seq_ref, const uint8_t* seq_qer, int64_t n,
                       const int32_t* w, int32_t* score, int32_t* n_cigar, uint32_t* cigar, int64_t cigar_cap,
                       int64_t* cigar_off)
{
    bsw_stats& S = eng->stats;
    const double t_begin = now_ms();
    DevCtx& c = eng->devs[0];
    CUDA_TRY(cudaSetDevice(c.dev));
    if (!eng->gbufs) eng->gbufs = new GlobalBufs();
    GlobalBufs& B = *static_cast<GlobalBufs*>(eng->gbufs);
    for (GlobalSlot& G : B.slot)
        if (!G.st) {
            CUDA_TRY(cudaStreamCreateWithFlags(&G.st, cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreate(&G.e0));
            CUDA_TRY(cudaEventCreate(&G.e1));
        }
    GlobalParams GP{eng->p.o_del, eng->p.e_del, eng->p.o_ins, eng->p.e_ins, eng->p.match, -eng->p.mismatch, eng->p.ambig};
    const long long Z_CAP = 3ll << 29, C_CAP = 1ll << 27, EH_CAP = 1ll << 26;      // bytes / words / cells per chunk
    const int64_t M_CAP = 65536;                                                    // alignments per chunk
    std::vector<GlobalDesc>& hd = B.hd;

    // host part of a chunk: descriptors, gather into page-locked staging, work order; then everything the device
    // does with it, enqueued on the slot's stream (nothing here waits for the device)
    auto launch = [&](GlobalSlot& G, int64_t first) -> int {
        const double t_host0 = now_ms();
        hd.clear();
        long long zb = 0, cw = 0, qb = 0, rb = 0;
        int qmax = 0, wmax = 0;
        int64_t m = 0;
        while (first + m < n && m < M_CAP) {
            const SeqPair& sp = pairs[first + m];
            const int wv = w[first + m];
            const long long n_col = sp.len2 < 2 * wv + 1 ? sp.len2 : 2 * wv + 1;
            const long long zi = ((n_col + 7) & ~7ll) * sp.len1, ci = (long long)sp.len1 + sp.len2;   // row pitch of bsw_global_kernel
            const int qm = std::max(qmax, sp.len2);
            if (m > 0 && (zb + zi > Z_CAP || cw + ci > C_CAP || (long long)(qm + 1) * (m + 1) > EH_CAP)) break;
            GlobalDesc d;
            d.qoff = (uint32_t)qb; d.roff = (uint32_t)rb; d.qlen = sp.len2; d.tlen = sp.len1; d.w = wv; d.idx = (int32_t)m;
            d.zoff = zb; d.coff = cw;
            hd.push_back(d);
            zb += zi; cw += ci; qb += sp.len2; rb += sp.len1; qmax = qm; wmax = std::max(wmax, wv);
            ++m;
        }
        G.first = first; G.m = m; G.cap_words = cw;
        if (int rc = ensure(eng, G.q, (size_t)qb + 16, true)) return rc;      // pinned staging twins: the copies run at link speed
        if (int rc = ensure(eng, G.r, (size_t)rb + 16, true)) return rc;
        if (int rc = ensure(eng, G.desc, (size_t)m, true)) return rc;
        uint8_t* const hq = G.q.h; uint8_t* const hr = G.r.h;
        std::vector<long long> cells_part((size_t)eng->pool->size(), 0);
        eng->pool->for_range(m, 1024, [&](int64_t b, int64_t e, int tid) {
            long long cells = 0;
            for (int64_t k = b; k < e; ++k) {
                const SeqPair& sp = pairs[first + k];
                const GlobalDesc& d = hd[(size_t)k];
                memcpy(hq + d.qoff, seq_qer + sp.idq, (size_t)sp.len2);
                memcpy(hr + d.roff, seq_ref + sp.idr, (size_t)sp.len1);
                // DP cells inside the band: sum over rows i < tlen of min(i + w + 1, qlen) - max(i - w, 0), in closed
                // form (|tlen - qlen| <= w keeps every row's window non-empty)
                const long long T = d.tlen, Q = d.qlen, W = d.w;
                const long long a = std::min(T, std::max(0ll, Q - W - 1));           // rows whose window ends before qlen
                const long long kb = std::max(0ll, T - W - 1);                        // rows whose window starts after 0
                cells += a * (a - 1) / 2 + a * (W + 1) + (T - a) * Q - kb * (kb + 1) / 2;
            }
            cells_part[(size_t)tid] += cells;
        });
        G.cells = 0;
        for (long long v : cells_part) G.cells += v;
        // threads run in order of decreasing band width, then target length: the column loop's trip count is what
        // the lanes of a warp share row by row, and the longest alignments start first (idx keeps the input position)
        B.key.resize((size_t)m);
        for (int64_t k = 0; k < m; ++k) {
            const GlobalDesc& d = hd[(size_t)k];
            const uint64_t band = (uint64_t)(d.qlen < 2 * d.w + 1 ? d.qlen : 2 * d.w + 1);
            B.key[(size_t)k] = (~((band << 16 | (uint64_t)d.tlen) << 20) & 0xfffffffffff00000ull) | (uint64_t)k;
        }
        radix_sort_work(B.key, B.tmp);
        for (int64_t k = 0; k < m; ++k) G.desc.h[k] = hd[(size_t)(B.key[(size_t)k] & 0xfffff)];
        const int threads = (int)m, stride = ((threads + 31) / 32) * 32;
        if (int rc = ensure(eng, G.z, (size_t)zb + 16)) return rc;
        if (int rc = ensure(eng, G.cig, (size_t)cw + 16)) return rc;
        if (int rc = ensure(eng, G.packed, (size_t)cw + 16, true)) return rc;
        const int W = std::max(2 * wmax + 2, 16);                    // live columns of a row (bsw_global.cuh; >= 16: its 8-column blocks wrap once)
        const int qstride = 4 * (((qmax + 3) / 4) | 1);               // query bytes per thread in shared memory
        const size_t smem = (size_t)W * GLOBAL_BLOCK * sizeof(int2) + (size_t)qstride * GLOBAL_BLOCK;
        const bool use_smem = smem <= 200 * 1024;
        if (!use_smem) if (int rc = ensure(eng, G.eh, (size_t)(qmax + 1) * (size_t)stride)) return rc;
        if (int rc = ensure(eng, G.score, (size_t)m, true)) return rc;
        if (int rc = ensure(eng, G.ncig, (size_t)m, true)) return rc;
        if (int rc = ensure(eng, G.off, (size_t)m + 1, true)) return rc;
        cudaStream_t st = G.st;
        CUDA_TRY(cudaMemcpyAsync(G.desc.d, G.desc.h, sizeof(GlobalDesc) * (size_t)m, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(G.q.d, hq, (size_t)qb, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(G.r.d, hr, (size_t)rb, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaEventRecord(G.e0, st));
        const int gblocks = (threads + GLOBAL_BLOCK - 1) / GLOBAL_BLOCK;
        if (use_smem) {
            if (!eng->global_attr_set) {
                CUDA_TRY(cudaFuncSetAttribute(bsw_global_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                eng->global_attr_set = true;
            }
            bsw_global_kernel<true><<<gblocks, GLOBAL_BLOCK, smem, st>>>(G.desc.d, threads, G.q.d, G.r.d, nullptr, 0, W, qstride, G.z.d,
                                                                         G.cig.d, G.score.d, G.ncig.d, GP);
        } else {
            bsw_global_kernel<false><<<gblocks, GLOBAL_BLOCK, 0, st>>>(G.desc.d, threads, G.q.d, G.r.d, G.eh.d, stride, 0, 0, G.z.d,
                                                                       G.cig.d, G.score.d, G.ncig.d, GP);
        }
        CUDA_TRY(cudaEventRecord(G.e1, st));
        // offsets of the packed operation lists (device scan), compaction, and the way out -- the packed lists are
        // copied as far as the chunk's upper bound allows without knowing their total: first the scalars, and after
        // the host has seen the total, exactly that many operations
        bsw_cigar_offsets<<<1, 1024, 0, st>>>(G.ncig.d, threads, G.off.d);
        bsw_cigar_compact<<<(threads + 127) / 128, 128, 0, st>>>(G.desc.d, threads, G.cig.d, G.ncig.d, G.off.d, G.packed.d);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(G.score.h, G.score.d, sizeof(int32_t) * (size_t)m, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(G.ncig.h, G.ncig.d, sizeof(int32_t) * (size_t)m, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(G.off.h, G.off.d, sizeof(long long) * ((size_t)m + 1), cudaMemcpyDeviceToHost, st));
        S.kernel_launches += 3;
        S.h2d_bytes += (int64_t)(sizeof(GlobalDesc) * (size_t)m + (size_t)qb + (size_t)rb);
        S.ms_pack += now_ms() - t_host0;                 // host: descriptors, gather, work order, enqueue
        return BSW_OK;
    };
    long long out_pos = 0;
    // results of a chunk into the caller's arrays (behind those of the earlier chunks)
    auto finish = [&](GlobalSlot& G) -> int {
        const double t_w0 = now_ms();
        CUDA_TRY(cudaStreamSynchronize(G.st));
        const double t_w1 = now_ms();
        S.ms_d2h += t_w1 - t_w0;                         // host waiting for the device
        const int64_t m = G.m, first = G.first;
        const long long run = G.off.h[m];
        if (out_pos + run > cigar_cap) {
            eng->err = "bsw_global: cigar buffer too small (len1 + len2 entries per alignment always suffice)";
            return BSW_ERR_PARAM;
        }
        if (run > 0) CUDA_TRY(cudaMemcpyAsync(G.packed.h, G.packed.d, sizeof(uint32_t) * (size_t)run, cudaMemcpyDeviceToHost, G.st));
        memcpy(score + first, G.score.h, sizeof(int32_t) * (size_t)m);
        memcpy(n_cigar + first, G.ncig.h, sizeof(int32_t) * (size_t)m);
        for (int64_t k = 0; k < m; ++k) cigar_off[first + k + 1] = out_pos + G.off.h[k + 1];
        for (int64_t k = 0; k < m; ++k) S.cells_nominal += (int64_t)pairs[first + k].len1 * pairs[first + k].len2;
        CUDA_TRY(cudaStreamSynchronize(G.st));
        if (run > 0) memcpy(cigar + out_pos, G.packed.h, sizeof(uint32_t) * (size_t)run);
        float ms = 0;
        if (cudaEventElapsedTime(&ms, G.e0, G.e1) == cudaSuccess) S.ms_kernel += (double)ms;
        S.cells_effective += G.cells;
        S.d2h_bytes += (int64_t)(16 * m + 4 * run);
        out_pos += run;
        S.ms_scatter += now_ms() - t_w1;                 // host: results into the caller's arrays (incl. the wait for the packed lists)
        return BSW_OK;
    };
    int64_t next = 0;
    int cur = 0;
    bool pending[2] = {false, false};
    while (next < n) {
        GlobalSlot& G = B.slot[cur];
        if (pending[cur]) { if (int rc = finish(G)) { cudaDeviceSynchronize(); return rc; } pending[cur] = false; }
        if (int rc = launch(G, next)) { cudaDeviceSynchronize(); return rc; }
        pending[cur] = true;
        next += G.m;
        cur ^= 1;
    }
    // drain in submission order: the slot that was launched first is the one `cur` points at now
    for (int k = 0; k < 2; ++k) {
        GlobalSlot& G = B.slot[cur];
        if (pending[cur]) { if (int rc = finish(G)) { cudaDeviceSynchronize(); return rc; } pending[cur] = false; }
        cur ^= 1;
    }
    S.ms_total = now_ms() - t_begin;
    return BSW_OK;
}

// second form (bsw_global2.cuh + bsw_global_plan.h): arguments validated by bsw_global, statistics zeroed, n > 0,
// every alignment inside the second kernel's domain with the slot width r16 says
static int run_global2(bsw_engine* eng, const SeqPair* pairs, const uint8_t* seq_ref, const uint8_t* seq_qer, int64_t n,
                       const int32_t* w, int32_t* score, int32_t* n_cigar, uint32_t* cigar, int64_t cigar_cap,
                       int64_t* cigar_off, bool r16)
{
    bsw_stats& S = eng->stats;
    const double t_begin = now_ms();
    DevCtx& c = eng->devs[0];
    CUDA_TRY(cudaSetDevice(c.dev));
    if (!eng->gbufs) eng->gbufs = new GlobalBufs();
    GlobalBufs& B = *static_cast<GlobalBufs*>(eng->gbufs);
    for (GlobalSlot& G : B.slot)
        if (!G.st) {
            CUDA_TRY(cudaStreamCreateWithFlags(&G.st, cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreate(&G.e0));
            CUDA_TRY(cudaEventCreate(&G.e1));
        }
    g2::Params GP{};
    GP.o_del = eng->p.o_del; GP.e_del = eng->p.e_del; GP.o_ins = eng->p.o_ins; GP.e_ins = eng->p.e_ins;
    g2::fill_table(GP, eng->p.match, -eng->p.mismatch, eng->p.ambig);
    if (!eng->global2_attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(g2::bsw_global2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G2_SMEM_MAX));
        CUDA_TRY(cudaFuncSetAttribute(g2::bsw_global2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G2_SMEM_MAX));
        // resident blocks are bounded by shared memory alone (47 registers, 20 - 40 KB a block): all of it to shared memory
        CUDA_TRY(cudaFuncSetAttribute(g2::bsw_global2_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CUDA_TRY(cudaFuncSetAttribute(g2::bsw_global2_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        eng->global2_attr_set = true;
    }
    // alignments per chunk: enough of them that the device's threads stay occupied while the longest alignments of the chunk
    // finish (they start first), few enough that the host plans chunk k + 1 while chunk k runs; BSW_GLOBAL_CHUNK overrides
    g2::Caps caps;
    caps.m = 1 << 17;
    if (const char* ch = getenv("BSW_GLOBAL_CHUNK")) caps.m = std::min<int64_t>(1 << 18, std::max<int64_t>(1, atoll(ch)));
    auto par = [&](int64_t cnt, int64_t grain, auto&& fn) { eng->pool->for_range(cnt, grain, fn); };
    const int slices = eng->pool->size();

    // host part of a chunk: plan (sizes, descriptors, gather into page-locked staging, work order, launches); then
    // everything the device does with it, enqueued on the slot's stream (nothing here waits for the device)
    auto launch = [&](GlobalSlot& G, g2::ChunkPlan& pl, int64_t first) -> int {
        const double t_host0 = now_ms();
        g2::plan_sizes(pairs, w, first, n, caps, par, pl);
        const int64_t m = pl.m;
        G.first = first; G.m = m; G.cap_words = pl.cig_words; G.cells = pl.cells;
        if (int rc = ensure(eng, G.q, (size_t)pl.q_bytes + 16, true)) return rc;      // pinned staging twins: the copies run at link speed
        if (int rc = ensure(eng, G.r, (size_t)pl.r_bytes + 16, true)) return rc;
        if (int rc = ensure(eng, G.desc, (size_t)m, true)) return rc;
        if (int rc = ensure(eng, G.z, (size_t)pl.z_bytes + 16)) return rc;
        if (int rc = ensure(eng, G.cig, (size_t)pl.cig_words + 16)) return rc;
        if (int rc = ensure(eng, G.packed, (size_t)pl.cig_words + 16, true)) return rc;
        if (int rc = ensure(eng, G.score, (size_t)m, true)) return rc;
        if (int rc = ensure(eng, G.ncig, (size_t)m, true)) return rc;
        if (int rc = ensure(eng, G.off, (size_t)m + 1, true)) return rc;
        g2::plan_fill(pairs, w, seq_ref, seq_qer, slices, par, pl, G.desc.h, G.q.h, G.r.h);
        cudaStream_t st = G.st;
        const int threads = (int)m;
        CUDA_TRY(cudaMemcpyAsync(G.desc.d, G.desc.h, sizeof(GlobalDesc) * (size_t)m, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(G.q.d, G.q.h, (size_t)pl.q_bytes + 8, cudaMemcpyHostToDevice, st));     // (+ 8: the last query's padding)
        CUDA_TRY(cudaMemcpyAsync(G.r.d, G.r.h, (size_t)pl.r_bytes + 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaEventRecord(G.e0, st));
        // one launch per class, largest rows first; several classes run side by side on the device's DP streams
        const int nl = (int)pl.launches.size();
        const int used = nl > 1 && !getenv("BSW_GLOBAL_SERIAL") ? std::min(nl, NSTREAMS) : 0;      // (BSW_GLOBAL_SERIAL: A/B, one stream)
        for (int k = 0; k < used; ++k) CUDA_TRY(cudaStreamWaitEvent(c.cs[k], G.e0, 0));
        for (int k = 0; k < nl; ++k) {
            const g2::Launch& L = pl.launches[(size_t)k];
            cudaStream_t ks = used ? c.cs[k % NSTREAMS] : st;
            const int gblocks = (L.count + g2::BLOCK - 1) / g2::BLOCK;
            const size_t smem = g2::smem_bytes(r16, L.slots, L.qwords);
            if (r16)
                g2::bsw_global2_kernel<true><<<gblocks, g2::BLOCK, smem, ks>>>(G.desc.d + L.first, L.count, G.q.d, G.r.d, L.slots, G.z.d,
                                                                               G.cig.d, G.score.d, G.ncig.d, GP);
            else
                g2::bsw_global2_kernel<false><<<gblocks, g2::BLOCK, smem, ks>>>(G.desc.d + L.first, L.count, G.q.d, G.r.d, L.slots, G.z.d,
                                                                                G.cig.d, G.score.d, G.ncig.d, GP);
        }
        CUDA_TRY(cudaGetLastError());
        for (int k = 0; k < used; ++k) {
            CUDA_TRY(cudaEventRecord(c.ev_join[k], c.cs[k]));
            CUDA_TRY(cudaStreamWaitEvent(st, c.ev_join[k], 0));
        }
        CUDA_TRY(cudaEventRecord(G.e1, st));
        bsw_cigar_offsets<<<1, 1024, 0, st>>>(G.ncig.d, threads, G.off.d);
        bsw_cigar_compact<<<(threads + 127) / 128, 128, 0, st>>>(G.desc.d, threads, G.cig.d, G.ncig.d, G.off.d, G.packed.d);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(G.score.h, G.score.d, sizeof(int32_t) * (size_t)m, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(G.ncig.h, G.ncig.d, sizeof(int32_t) * (size_t)m, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(G.off.h, G.off.d, sizeof(long long) * ((size_t)m + 1), cudaMemcpyDeviceToHost, st));
        S.kernel_launches += nl + 2;
        S.h2d_bytes += (int64_t)(sizeof(GlobalDesc) * (size_t)m + (size_t)pl.q_bytes + (size_t)pl.r_bytes);
        S.cells_nominal += pl.cells_nominal;
        S.ms_pack += now_ms() - t_host0;                 // host: plan, gather, work order, enqueue
        return BSW_OK;
    };
    long long out_pos = 0;
    // results of a chunk into the caller's arrays (behind those of the earlier chunks), the copies on the pool
    auto finish = [&](GlobalSlot& G) -> int {
        const double t_w0 = now_ms();
        CUDA_TRY(cudaStreamSynchronize(G.st));
        const double t_w1 = now_ms();
        S.ms_d2h += t_w1 - t_w0;                         // host waiting for the device
        const int64_t m = G.m, first = G.first;
        const long long run = G.off.h[m];
        if (out_pos + run > cigar_cap) {
            eng->err = "bsw_global: cigar buffer too small (len1 + len2 entries per alignment always suffice)";
            return BSW_ERR_PARAM;
        }
        if (run > 0) CUDA_TRY(cudaMemcpyAsync(G.packed.h, G.packed.d, sizeof(uint32_t) * (size_t)run, cudaMemcpyDeviceToHost, G.st));
        const long long base = out_pos;
        eng->pool->for_range(m, 16384, [&](int64_t b, int64_t e, int) {
            memcpy(score + first + b, G.score.h + b, sizeof(int32_t) * (size_t)(e - b));
            memcpy(n_cigar + first + b, G.ncig.h + b, sizeof(int32_t) * (size_t)(e - b));
            for (int64_t k = b; k < e; ++k) cigar_off[first + k + 1] = base + G.off.h[k + 1];
        });
        CUDA_TRY(cudaStreamSynchronize(G.st));
        if (run > 0)
            eng->pool->for_range(run, 1 << 18, [&](int64_t b, int64_t e, int) {
                memcpy(cigar + base + b, G.packed.h + b, sizeof(uint32_t) * (size_t)(e - b));
            });
        float ms = 0;
        if (cudaEventElapsedTime(&ms, G.e0, G.e1) == cudaSuccess) S.ms_kernel += (double)ms;
        S.cells_effective += G.cells;
        S.d2h_bytes += (int64_t)(16 * m + 4 * run);
        out_pos += run;
        S.ms_scatter += now_ms() - t_w1;                 // host: results into the caller's arrays (incl. the wait for the packed lists)
        return BSW_OK;
    };
    int64_t next = 0;
    int cur = 0;
    bool pending[2] = {false, false};
    while (next < n) {
        GlobalSlot& G = B.slot[cur];
        if (pending[cur]) { if (int rc = finish(G)) { cudaDeviceSynchronize(); return rc; } pending[cur] = false; }
        if (int rc = launch(G, B.plan[cur], next)) { cudaDeviceSynchronize(); return rc; }
        pending[cur] = true;
        next += G.m;
        cur ^= 1;
    }
    // drain in submission order: the slot that was launched first is the one `cur` points at now
    for (int k = 0; k < 2; ++k) {
        GlobalSlot& G = B.slot[cur];
        if (pending[cur]) { if (int rc = finish(G)) { cudaDeviceSynchronize(); return rc; } pending[cur] = false; }
        cur ^= 1;
    }
    S.ms_total = now_ms() - t_begin;
    return BSW_OK;
}

int bsw_global(bsw_engine* eng, const SeqPair* pairs, const uint8_t* seq_ref, const uint8_t* seq_qer, int64_t n,
               const int32_t* w, int32_t* score, int32_t* n_cigar, uint32_t* cigar, int64_t cigar_cap,
               int64_t* cigar_off)
{
    if (!eng) return BSW_ERR_PARAM;
    eng->err.clear();
    if (n < 0 || (n > 0 && (!pairs || !seq_ref || !seq_qer || !w || !score || !n_cigar || !cigar || !cigar_off)) || cigar_cap < 0) {
        eng->err = "bsw_global: bad arguments";
        return BSW_ERR_PARAM;
    }
    bsw_stats& S = eng->stats;
    memset(&S, 0, sizeof(S));
    S.pairs = n;
    if (cigar_off) cigar_off[0] = 0;
    if (n == 0) return BSW_OK;
    // one pass: the domain of the entry point, and which kernel takes the call -- the second wants every score in a
    // signed byte and every alignment's rows + query in shared memory; its 16-bit slots want every value in 16 bits
    g2::Params GP{};
    GP.o_del = eng->p.o_del; GP.e_del = eng->p.e_del; GP.o_ins = eng->p.o_ins; GP.e_ins = eng->p.e_ins;
    const int match = eng->p.match, mm = -eng->p.mismatch, ambig = eng->p.ambig;
    std::atomic<int> bad{0}, wide{0}, big16{0}, big32{0};
    eng->pool->for_range(n, 16384, [&](int64_t b, int64_t e, int) {
        int bad_l = 0, wide_l = 0, big16_l = 0, big32_l = 0;
        for (int64_t i = b; i < e; ++i) {
            const SeqPair& sp = pairs[i];
            const long long dl = (long long)sp.len1 - (long long)sp.len2;
            if (sp.len1 < 1 || sp.len1 > 32767 || sp.len2 < 1 || sp.len2 > 32767 || sp.idr < 0 || sp.idq < 0 || w[i] < 0 ||
                w[i] > 32767 || (dl < 0 ? -dl : dl) > w[i]) { bad_l = 1; continue; }
            const int wv = g2::eff_w(sp.len2, sp.len1, w[i]);
            const int sc = g2::slots_class(g2::row_slots(sp.len2, wv)), qc = g2::qwords_class(g2::query_words(sp.len2));
            if (!g2::rows16_ok(GP, match, mm, ambig, sp.len2, sp.len1, wv)) wide_l = 1;
            if (g2::smem_bytes(true, sc, qc) > G2_SMEM_MAX) big16_l = 1;
            if (g2::smem_bytes(false, sc, qc) > G2_SMEM_MAX) big32_l = 1;
        }
        if (bad_l) bad.store(1, std::memory_order_relaxed);
        if (wide_l) wide.store(1, std::memory_order_relaxed);
        if (big16_l) big16.store(1, std::memory_order_relaxed);
        if (big32_l) big32.store(1, std::memory_order_relaxed);
    });
    if (bad.load()) {
        eng->err = "bsw_global: need 1 <= len1, len2 <= 32767, offsets >= 0 and |len1 - len2| <= w <= 32767 "
                   "(outside the band the reference's backtrack leaves its matrix, ksw.c:593-595)";
        return BSW_ERR_DOMAIN;
    }
    // one slot width per call: 16 bits when every value fits them and every alignment fits shared memory at that
    // width, else 64-bit slots if every alignment fits at that width, else the first kernel.
    // BSW_GLOBAL_KERNEL (A/B runs): 1 = the first kernel, 2w = the second with 64-bit slots
    const char* pick = getenv("BSW_GLOBAL_KERNEL");
    const bool want1 = pick && pick[0] == '1', want_wide = pick && pick[0] == '2' && pick[1] == 'w';
    const bool can16 = !wide.load() && !big16.load(), can32 = !big32.load();
    const bool r16 = can16 && !want_wide;
    const bool second = g2::scores_ok(match, mm, ambig) && !want1 && (r16 || can32);
    if (second) return run_global2(eng, pairs, seq_ref, seq_qer, n, w, score, n_cigar, cigar, cigar_cap, cigar_off, r16);
    return run_global1(eng, pairs, seq_ref, seq_qer, n, w, score, n_cigar, cigar, cigar_cap, cigar_off);
}
