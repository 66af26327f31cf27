// bsw_global.inl -- host side of bsw_global (banded global alignment + CIGAR);
// included by bsw_engine.cu inside extern "C".  Chunks of alignments go to the engine's first device:
// gather the two byte strings of every alignment, one thread per alignment computes score and
// operation list (bsw_global.cuh), a second kernel packs the lists, and they land in the caller's
// buffer one after the other (cigar_off[]).
namespace {

struct GlobalBufs {                      // grow-only device buffers of bsw_global, owned by the engine
    Buf<uint8_t> q, r, z;
    Buf<GlobalDesc> desc;
    Buf<int2> eh;
    Buf<uint32_t> cig, packed;
    Buf<int32_t> score, ncig;
    Buf<long long> off;
};

} // namespace

static void bsw_global_release(bsw_engine* eng)
{
    if (!eng->gbufs) return;
    GlobalBufs* B = static_cast<GlobalBufs*>(eng->gbufs);
    if (!eng->devs.empty()) cudaSetDevice(eng->devs[0].dev);
    release(B->q); release(B->r); release(B->z); release(B->desc); release(B->eh); release(B->cig);
    release(B->packed); release(B->score); release(B->ncig); release(B->off);
    delete B;
    eng->gbufs = nullptr;
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
    const double t_begin = now_ms();
    for (int64_t i = 0; i < n; ++i) {
        const SeqPair& sp = pairs[i];
        const long long dl = (long long)sp.len1 - (long long)sp.len2;
        if (sp.len1 < 1 || sp.len1 > 32767 || sp.len2 < 1 || sp.len2 > 32767 || sp.idr < 0 || sp.idq < 0 || w[i] < 0 ||
            w[i] > 32767 || (dl < 0 ? -dl : dl) > w[i]) {
            eng->err = "bsw_global: need 1 <= len1, len2 <= 32767, offsets >= 0 and |len1 - len2| <= w <= 32767 "
                       "(outside the band the reference's backtrack leaves its matrix, ksw.c:593-595)";
            return BSW_ERR_DOMAIN;
        }
    }
    DevCtx& c = eng->devs[0];
    CUDA_TRY(cudaSetDevice(c.dev));
    if (!eng->gbufs) eng->gbufs = new GlobalBufs();
    GlobalBufs& B = *static_cast<GlobalBufs*>(eng->gbufs);
    cudaStream_t st = c.cs[0];
    GlobalParams GP{eng->p.o_del, eng->p.e_del, eng->p.o_ins, eng->p.e_ins, eng->p.match, -eng->p.mismatch, eng->p.ambig};
    std::vector<GlobalDesc> hd;
    std::vector<uint64_t> order;
    std::vector<long long> hoff;
    std::vector<int32_t> hn;
    const long long Z_CAP = 3ll << 30, C_CAP = 1ll << 28, EH_CAP = 1ll << 27;      // bytes / words / cells per chunk
    int64_t done = 0;
    long long out_pos = 0;
    while (done < n) {
        // chunk: as many alignments as fit the direction-matrix, operation-list and row-scratch budgets
        hd.clear();
        long long zb = 0, cw = 0, qb = 0, rb = 0;
        int qmax = 0, wmax = 0;
        int64_t m = 0;
        while (done + m < n && m < 131072) {
            const SeqPair& sp = pairs[done + m];
            const int wv = w[done + m];
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
        if (int rc = ensure(eng, B.q, (size_t)qb + 16, true)) return rc;      // pinned staging twins: the H2D copies run at link speed
        if (int rc = ensure(eng, B.r, (size_t)rb + 16, true)) return rc;
        if (int rc = ensure(eng, B.desc, (size_t)m, true)) return rc;
        uint8_t* const hq = B.q.h; uint8_t* const hr = B.r.h;
        std::vector<long long> cells_part((size_t)eng->pool->size(), 0);
        eng->pool->for_range(m, 1024, [&](int64_t b, int64_t e, int tid) {
            long long cells = 0;
            for (int64_t k = b; k < e; ++k) {
                const SeqPair& sp = pairs[done + k];
                const GlobalDesc& d = hd[(size_t)k];
                memcpy(hq + d.qoff, seq_qer + sp.idq, (size_t)sp.len2);
                memcpy(hr + d.roff, seq_ref + sp.idr, (size_t)sp.len1);
                for (int i = 0; i < d.tlen; ++i) {                   // DP cells inside the band
                    const int beg = i > d.w ? i - d.w : 0, end = i + d.w + 1 < d.qlen ? i + d.w + 1 : d.qlen;
                    if (end > beg) cells += end - beg;
                }
            }
            cells_part[(size_t)tid] += cells;
        });
        for (long long v : cells_part) S.cells_effective += v;
        // threads run in order of decreasing target length: the lanes of a warp then finish together and
        // the longest alignments start first (idx keeps the input position for the outputs)
        // (band width first: the column loop's trip count is what the lanes of a warp share row by row)
        order.resize((size_t)m);
        for (int64_t k = 0; k < m; ++k) {
            const GlobalDesc& d = hd[(size_t)k];
            const uint64_t band = (uint64_t)(d.qlen < 2 * d.w + 1 ? d.qlen : 2 * d.w + 1);
            order[(size_t)k] = ~((band << 16 | (uint64_t)d.tlen) << 20) & ~0xfffffull | (uint64_t)k;   // descending work, then input order
        }
        std::sort(order.begin(), order.end());
        for (int64_t k = 0; k < m; ++k) B.desc.h[k] = hd[(size_t)(order[(size_t)k] & 0xfffff)];
        const int threads = (int)m, stride = ((threads + 31) / 32) * 32;
        if (int rc = ensure(eng, B.z, (size_t)zb + 16)) return rc;
        if (int rc = ensure(eng, B.cig, (size_t)cw + 16)) return rc;
        const int W = std::max(2 * wmax + 2, 16);                    // live columns of a row (bsw_global.cuh; >= 16: its 8-column blocks wrap once)
        const int qstride = 4 * (((qmax + 3) / 4) | 1);               // query bytes per thread in shared memory
        const size_t smem = (size_t)W * GLOBAL_BLOCK * sizeof(int2) + (size_t)qstride * GLOBAL_BLOCK;
        const bool use_smem = smem <= 200 * 1024;
        if (!use_smem) if (int rc = ensure(eng, B.eh, (size_t)(qmax + 1) * (size_t)stride)) return rc;
        if (int rc = ensure(eng, B.score, (size_t)m)) return rc;
        if (int rc = ensure(eng, B.ncig, (size_t)m)) return rc;
        if (int rc = ensure(eng, B.off, (size_t)m + 1)) return rc;
        CUDA_TRY(cudaMemcpyAsync(B.desc.d, B.desc.h, sizeof(GlobalDesc) * (size_t)m, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(B.q.d, hq, (size_t)qb, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(B.r.d, hr, (size_t)rb, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaEventRecord(c.ev_t0, st));
        const int gblocks = (threads + GLOBAL_BLOCK - 1) / GLOBAL_BLOCK;
        if (use_smem) {
            if (!eng->global_attr_set) {
                CUDA_TRY(cudaFuncSetAttribute(bsw_global_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                eng->global_attr_set = true;
            }
            bsw_global_kernel<true><<<gblocks, GLOBAL_BLOCK, smem, st>>>(B.desc.d, threads, B.q.d, B.r.d, nullptr, 0, W, qstride, B.z.d,
                                                                         B.cig.d, B.score.d, B.ncig.d, GP);
        } else {
            bsw_global_kernel<false><<<gblocks, GLOBAL_BLOCK, 0, st>>>(B.desc.d, threads, B.q.d, B.r.d, B.eh.d, stride, 0, 0, B.z.d,
                                                                       B.cig.d, B.score.d, B.ncig.d, GP);
        }
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(c.ev_t1, st));
        CUDA_TRY(cudaMemcpyAsync(score + done, B.score.d, sizeof(int32_t) * (size_t)m, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(n_cigar + done, B.ncig.d, sizeof(int32_t) * (size_t)m, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c.ev_t0, c.ev_t1) == cudaSuccess) S.ms_kernel += (double)ms;
        // pack the operation lists behind those of the earlier chunks
        hoff.resize((size_t)m + 1);
        long long run = 0;
        for (int64_t k = 0; k < m; ++k) { hoff[(size_t)k] = run; run += n_cigar[done + k]; cigar_off[done + k + 1] = out_pos + run; }
        hoff[(size_t)m] = run;
        if (out_pos + run > cigar_cap) {
            eng->err = "bsw_global: cigar buffer too small (len1 + len2 entries per alignment always suffice)";
            return BSW_ERR_PARAM;
        }
        if (run > 0) {
            if (int rc = ensure(eng, B.packed, (size_t)run)) return rc;
            CUDA_TRY(cudaMemcpyAsync(B.off.d, hoff.data(), sizeof(long long) * ((size_t)m + 1), cudaMemcpyHostToDevice, st));
            bsw_cigar_compact<<<(threads + 127) / 128, 128, 0, st>>>(B.desc.d, threads, B.cig.d, B.ncig.d, B.off.d, B.packed.d);
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaMemcpyAsync(cigar + out_pos, B.packed.d, sizeof(uint32_t) * (size_t)run, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
        }
        S.kernel_launches += 2;
        S.h2d_bytes += (int64_t)(sizeof(GlobalDesc) * (size_t)m + (size_t)qb + (size_t)rb);
        S.d2h_bytes += (int64_t)(8 * m + 4 * run);

        for (int64_t k = 0; k < m; ++k) S.cells_nominal += (int64_t)pairs[done + k].len1 * pairs[done + k].len2;
        out_pos += run;
        done += m;
    }
    S.ms_total = now_ms() - t_begin;
    return BSW_OK;
}
