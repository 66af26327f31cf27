// bsw_engine.cu -- engine object, host pipeline and kernel launches behind the C ABI (include/bsw.h).
//
// Stands in for BandedPairWiseSW's batch wrapper smithWatermanBatchWrapper16
// (benchmarks/bsw/bandedSWA.cpp:1150-1431): where the reference pads to the SIMD width, sorts by
// len1, transposes AoS->SoA per 16 pairs and calls the AVX kernel, this engine
//   1. buckets the batch by (len2, h0, len1)                       (bsw_host.cpp)
//   2. cuts the processing order into chunks and, per chunk, packs sequences to 2 bits/base
//      straight into pinned staging, copies them to HBM asynchronously, launches the sm_100a
//      kernels per shared-memory class and copies the packed results back -- chunk k's GPU work
//      overlaps the packing of chunk k+1 and the scatter of chunk k-1
//   3. scatters the six result fields into the caller's SeqPair[] in input order.
// Pairs that contain N (code 4) and pairs whose query exceeds the short kernel's shared-memory
// limit are staged as bytes and run by the byte variant of the short kernel / the warp-per-pair
// long kernel after the last chunk.
#include "bsw_common.h"
#include "bsw_kernels.cuh"
#include <cstdio>
#include <cstring>
#include <string>
#include <memory>
#include <deque>

namespace bsw {
void partition_blocks(const SortedBatch& sb, int32_t w, int32_t n_shards, std::vector<std::vector<int64_t>>& blocks_of);
}
using namespace bsw;

namespace {

constexpr int SHORT_BLOCK = 64;           // threads (= pairs) per block of the short kernel
constexpr int SHORT_MAX_QLEN = 832;       // eh words + query byte plane of SHORT_BLOCK threads must fit 227 KB
constexpr int NSTREAMS = 4;               // compute streams per device
constexpr int MAX_CHUNKS = 16;

std::string g_create_error;
std::mutex g_err_mutex;

struct Launch {
    int first, count;     // range in the shard's processing order
    int qstride;          // shared-memory rows per thread (>= qmax + 1)
};

template <class T>
struct Buf {              // grow-only device + pinned-host buffer pair
    T* d = nullptr; T* h = nullptr; size_t cap = 0;
};

struct Chunk {
    int64_t s0 = 0, s1 = 0;               // range in the shard's processing order
    size_t q0 = 0, q1 = 0, t0 = 0, t1 = 0; // word ranges in the packed sequence buffers
    std::vector<Launch> plan;
    cudaEvent_t ev_h2d{}, ev_done{};
    bool scattered = false;
};

struct DevCtx {
    int dev = 0;
    int sms = 148;
    cudaStream_t st_copy{}, st_d2h{}, st[NSTREAMS] = {};
    cudaEvent_t ev_a{}, ev_b{}, ev_k0{}, ev_k1{}, ev_join[NSTREAMS] = {};
    Chunk chunks[MAX_CHUNKS];
    int nchunks = 0;
    Buf<int4> meta, res, meta_n, res_n;
    Buf<uint32_t> q, t;
    Buf<uint8_t> qb, tb;
    std::vector<uint32_t> pos_n;          // byte-staged pair k -> position in the shard order
    std::vector<uint8_t> hasn;            // per position: pair contains an N
    std::vector<uint32_t> qoff, toff;     // word offsets per position (+1)
    unsigned long long* d_cells = nullptr;
    unsigned long long* h_cells = nullptr;
    unsigned int* d_queue = nullptr;      // work queue of the long-pair kernel
    Buf<uint32_t> scratch;                // eh[] rows of the long-pair kernel (device only)
    // the shard of the sorted batch this device owns (copied views, processing order)
    RawBuf<uint32_t> idx;
    RawBuf<uint16_t> len2, len1, h0;
    int64_t n = 0;                        // pairs in the shard
    int64_t n_short = 0;                  // positions [0, n_short) go to the short kernel
    int64_t n_bytes_short = 0;            // byte list: [0, n_bytes_short) short pairs with N,
    int64_t n_long = 0;                   //            [n_bytes_short, +n_long) long pairs
    int qmax_bytes_short = 0;
    int long_stride = 0, long_blocks = 0;
    size_t qb_bytes = 0, tb_bytes = 0;
    bool attr_set = false;
};

} // namespace

struct bsw_engine {
    bsw_params p;
    KParams kp;
    std::deque<DevCtx> devs;
    std::unique_ptr<ThreadPool> pool;
    std::string err;
    bsw_stats stats;
    int short_max = SHORT_MAX_QLEN;       // longest query the short kernel takes
    // staged batch
    bool staged = false, ran = false;
    int64_t n = 0;
    int32_t w = 0;
    SortedBatch sb;
};

namespace {

#define CUDA_TRY(call)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            eng->err = std::string(#call) + ": " + cudaGetErrorString(e_);                      \
            return BSW_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

template <class T>
int ensure(bsw_engine* eng, Buf<T>& b, size_t need, bool host = true)
{
    if (need <= b.cap) return BSW_OK;
    size_t cap = std::max(need + need / 4, (size_t)1024);
    if (b.d) cudaFree(b.d);
    if (b.h) cudaFreeHost(b.h);
    b.d = nullptr; b.h = nullptr; b.cap = 0;
    CUDA_TRY(cudaMalloc((void**)&b.d, cap * sizeof(T)));
    if (host) CUDA_TRY(cudaHostAlloc((void**)&b.h, cap * sizeof(T), cudaHostAllocDefault));
    b.cap = cap;
    return BSW_OK;
}

template <class T>
void release(Buf<T>& b)
{
    if (b.d) cudaFree(b.d);
    if (b.h) cudaFreeHost(b.h);
    b.d = nullptr; b.h = nullptr; b.cap = 0;
}

// 2-bit packing of n base codes (one per byte) into 16-bases-per-word little-endian words.
// Returns true if a code > 3 (N) was seen; such bases are packed as (code & 3).
inline bool pack2(const uint8_t* src, int n, uint32_t* dst)
{
    uint64_t bad = 0;
    int i = 0, wi = 0;
    for (; i + 16 <= n; i += 16, ++wi) {
        uint64_t a, b;
        memcpy(&a, src + i, 8); memcpy(&b, src + i + 8, 8);
        bad |= (a | b) & 0xFCFCFCFCFCFCFCFCull;
        a &= 0x0303030303030303ull; b &= 0x0303030303030303ull;
        a = (a | (a >> 6)) & 0x000F000F000F000Full;  a = (a | (a >> 12)) & 0x000000FF000000FFull;
        a = (a | (a >> 24)) & 0xFFFFull;
        b = (b | (b >> 6)) & 0x000F000F000F000Full;  b = (b | (b >> 12)) & 0x000000FF000000FFull;
        b = (b | (b >> 24)) & 0xFFFFull;
        dst[wi] = (uint32_t)(a | (b << 16));
    }
    if (i < n) {
        uint32_t wv = 0;
        for (int k = 0; i < n; ++i, ++k) {
            const uint8_t c = src[i];
            bad |= c & 0xFC;
            wv |= (uint32_t)(c & 3) << (2 * k);
        }
        dst[wi] = wv;
    }
    return bad != 0;
}

// Shared-memory rows per thread come in steps (one launch per step present in a chunk): fine
// steps where occupancy is most sensitive to them, coarser ones for long queries.
inline int stride_for(int qmax)
{
    const int need = qmax + 8;                    // + one prefetched group (bsw_kernels.cuh)
    if (need > SHORT_MAX_QLEN + 8) return -1;
    int s;
    if (need <= 136) s = (need + 7) & ~7;
    else if (need <= 520) s = (need + 15) & ~15;
    else s = (need + 31) & ~31;
    return std::min(s, SHORT_MAX_QLEN + 8);
}

// dynamic shared memory of one short-kernel block: eh words + the 2-bit query byte plane
inline size_t short_smem_bytes(int qstride)
{
    return (size_t)qstride * SHORT_BLOCK * sizeof(uint32_t) + (size_t)((qstride + 3) / 4) * SHORT_BLOCK;
}

int set_kernel_attrs(bsw_engine* eng, DevCtx& c)
{
    if (c.attr_set) return BSW_OK;
    const int maxsm = 227 * 1024;
    CUDA_TRY(cudaFuncSetAttribute(bsw_short_kernel<SHORT_BLOCK, false>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm));
    CUDA_TRY(cudaFuncSetAttribute(bsw_short_kernel<SHORT_BLOCK, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm));
    c.attr_set = true;
    return BSW_OK;
}

int validate_params(const bsw_params* p, std::string& why)
{
    auto bad = [&](const char* m) { why = m; return BSW_ERR_PARAM; };
    if (p->e_del < 1 || p->e_ins < 1) return bad("e_del / e_ins must be >= 1");
    if (p->o_del < 0 || p->o_ins < 0) return bad("o_del / o_ins must be >= 0");
    if (p->o_del + p->e_del > 16000 || p->o_ins + p->e_ins > 16000) return bad("gap penalties too large");
    if (p->match < 1 || p->match > 127) return bad("match must be in 1..127");
    if (p->mismatch < 0 || p->mismatch > 127) return bad("mismatch penalty must be in 0..127");
    if (p->ambig > p->match || p->ambig < -127) return bad("ambig must be in -127..match");
    if (p->zdrop_mode == BSW_ZDROP_VECTOR && (p->zdrop < 1 || p->zdrop > 32767))
        return bad("zdrop must be in 1..32767 (32767 = off); the reference's vector z-drop is not 'off' for <= 0");
    if (p->zdrop_mode != BSW_ZDROP_VECTOR && p->zdrop_mode != BSW_ZDROP_SCALAR) return bad("zdrop_mode");
    if (p->zdrop > 32767) return bad("zdrop must be <= 32767");
    if (p->end_bonus < 0 || p->end_bonus > 16000) return bad("end_bonus out of range");
    if (p->n_devices < 0 || p->n_devices > 16) return bad("n_devices must be in 0..16");
    if (p->long_min_qlen < 0) return bad("long_min_qlen must be >= 0");
    return BSW_OK;
}

// ------------------------------------------------------------------------------------------
// per-shard preparation: offsets, chunking, launch plans, buffer sizing
// ------------------------------------------------------------------------------------------
int prepare_shard(bsw_engine* eng, DevCtx& c, bool pipelined)
{
    CUDA_TRY(cudaSetDevice(c.dev));
    if (int rc = set_kernel_attrs(eng, c)) return rc;
    c.nchunks = 0;
    c.n_short = c.n_bytes_short = c.n_long = 0; c.qb_bytes = c.tb_bytes = 0; c.pos_n.clear();
    if (c.n == 0) return BSW_OK;
    // queries longer than the short kernel's limit sit at the end of the (len2-ascending) order
    c.n_short = std::upper_bound(c.len2.data(), c.len2.data() + c.n, (uint16_t)eng->short_max) - c.len2.data();
    c.n_long = c.n - c.n_short;
    c.qoff.resize((size_t)c.n_short + 1); c.toff.resize((size_t)c.n_short + 1);
    uint64_t qo = 0, to = 0;
    for (int64_t s = 0; s < c.n_short; ++s) {
        c.qoff[s] = (uint32_t)qo; c.toff[s] = (uint32_t)to;
        qo += (uint32_t)(c.len2[s] + 15) >> 4; to += (uint32_t)(c.len1[s] + 15) >> 4;
    }
    if (qo > 0xffffffffull || to > 0xffffffffull) { eng->err = "batch too large for 32-bit word offsets"; return BSW_ERR_PARAM; }
    c.qoff[c.n_short] = (uint32_t)qo; c.toff[c.n_short] = (uint32_t)to;
    if (int rc = ensure(eng, c.meta, (size_t)c.n_short + 1)) return rc;
    if (int rc = ensure(eng, c.res, (size_t)c.n_short + 1)) return rc;
    if (int rc = ensure(eng, c.q, qo + 4)) return rc;
    if (int rc = ensure(eng, c.t, to + 4)) return rc;
    c.hasn.assign((size_t)c.n_short, 0);

    // chunks: equal shares of the estimated work, cut at block boundaries
    int want = 1;
    if (pipelined && c.n_short >= 4 * 16384) want = (int)std::min<int64_t>(8, c.n_short / 65536 + 1);
    std::vector<int64_t> cuts{0};
    if (want > 1) {
        const int64_t nblk = (c.n_short + SHORT_BLOCK - 1) / SHORT_BLOCK;
        std::vector<double> pre((size_t)nblk + 1, 0.0);
        const double band = 2.0 * eng->w + 1;
        for (int64_t b = 0; b < nblk; ++b) {
            const int64_t s = std::min(c.n_short - 1, b * SHORT_BLOCK + SHORT_BLOCK - 1);
            const int64_t cnt = std::min<int64_t>(SHORT_BLOCK, c.n_short - b * SHORT_BLOCK);
            pre[b + 1] = pre[b] + (double)cnt * (64.0 + c.len1[s] * std::min<double>(c.len2[s], band) +
                                                 40.0 * (c.len1[s] + c.len2[s]));
        }
        for (int k = 1; k < want; ++k) {
            const double target = pre[nblk] * k / want;
            const int64_t b = std::lower_bound(pre.begin(), pre.end(), target) - pre.begin();
            const int64_t cut = std::min(c.n_short, b * SHORT_BLOCK);
            if (cut > cuts.back()) cuts.push_back(cut);
        }
    }
    if (c.n_short > cuts.back()) cuts.push_back(c.n_short);
    c.nchunks = (int)cuts.size() - 1;
    for (int k = 0; k < c.nchunks; ++k) {
        Chunk& ch = c.chunks[k];
        ch.s0 = cuts[k]; ch.s1 = cuts[k + 1];
        ch.q0 = c.qoff[ch.s0]; ch.q1 = c.qoff[ch.s1]; ch.t0 = c.toff[ch.s0]; ch.t1 = c.toff[ch.s1];
        ch.scattered = false;
        ch.plan.clear();
        for (int64_t s = ch.s0; s < ch.s1;) {
            const int64_t e = std::min(ch.s1, s + SHORT_BLOCK);
            const int qs = stride_for(c.len2[e - 1]);                 // ascending in len2
            if (!ch.plan.empty() && ch.plan.back().qstride == qs) ch.plan.back().count += (int)(e - s);
            else ch.plan.push_back(Launch{(int)s, (int)(e - s), qs});
            s = e;
        }
    }
    return BSW_OK;
}

// pack one chunk of a shard into pinned staging (host threads), flag N-containing pairs
void pack_chunk(bsw_engine* eng, DevCtx& c, const Chunk& ch, const SeqPair* pairs, const uint8_t* seq_ref,
                const uint8_t* seq_qer)
{
    (void)pairs;
    const uint32_t* idx = c.idx.data();
    const uint64_t* offr = eng->sb.offr.data();
    const uint64_t* offq = eng->sb.offq.data();
    eng->pool->for_range(ch.s1 - ch.s0, 1024, [&](int64_t b, int64_t e, int) {
        b += ch.s0; e += ch.s0;
        for (int64_t s = b; s < e; ++s) {
            if (s + 16 < e) { __builtin_prefetch(&offq[idx[s + 16]], 0, 0); __builtin_prefetch(&offr[idx[s + 16]], 0, 0); }
            if (s + 4 < e) {
                const uint32_t nx = idx[s + 4];
                __builtin_prefetch(seq_qer + offq[nx], 0, 0);
                __builtin_prefetch(seq_ref + offr[nx], 0, 0);
                __builtin_prefetch(seq_ref + offr[nx] + 64, 0, 0);
            }
            const uint32_t i = idx[s];
            const int l2 = c.len2[s], l1 = c.len1[s];
            const bool nq = pack2(seq_qer + offq[i], l2, c.q.h + c.qoff[s]);
            const bool nr = pack2(seq_ref + offr[i], l1, c.t.h + c.toff[s]);
            const int flag = (nq | nr) ? BSW_META_NFLAG : 0;
            c.hasn[s] = (uint8_t)(flag != 0);
            c.meta.h[s] = make_int4((int)c.qoff[s], (int)c.toff[s], l2 | (l1 << 16), c.h0[s] | flag);
        }
    });
}

int h2d_chunk(bsw_engine* eng, DevCtx& c, Chunk& ch)
{
    cudaStream_t st = c.st_copy;
    CUDA_TRY(cudaMemcpyAsync(c.meta.d + ch.s0, c.meta.h + ch.s0, sizeof(int4) * (size_t)(ch.s1 - ch.s0), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(c.q.d + ch.q0, c.q.h + ch.q0, sizeof(uint32_t) * (ch.q1 - ch.q0), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(c.t.d + ch.t0, c.t.h + ch.t0, sizeof(uint32_t) * (ch.t1 - ch.t0), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaEventRecord(ch.ev_h2d, st));
    eng->stats.h2d_bytes += (int64_t)(sizeof(int4) * (size_t)(ch.s1 - ch.s0) + 4 * ((ch.q1 - ch.q0) + (ch.t1 - ch.t0)));
    return BSW_OK;
}

// launches of one chunk, spread over the compute streams; longest class first
int launch_chunk(bsw_engine* eng, DevCtx& c, Chunk& ch, bool wait_h2d)
{
    int li = 0;
    bool used[NSTREAMS] = {};
    for (int k = (int)ch.plan.size() - 1; k >= 0; --k, ++li) {
        const Launch& L = ch.plan[k];
        const int si = li % NSTREAMS;
        cudaStream_t st = c.st[si];
        if (wait_h2d && !used[si]) CUDA_TRY(cudaStreamWaitEvent(st, ch.ev_h2d, 0));
        used[si] = true;
        const int grid = (L.count + SHORT_BLOCK - 1) / SHORT_BLOCK;
        const size_t smem = short_smem_bytes(L.qstride);
        bsw_short_kernel<SHORT_BLOCK, false><<<grid, SHORT_BLOCK, smem, st>>>(
            c.meta.d, c.q.d, c.t.d, c.res.d, L.first, L.count, L.qstride, eng->kp, c.d_cells);
        eng->stats.kernel_launches++;
    }
    CUDA_TRY(cudaGetLastError());
    return BSW_OK;
}

int d2h_chunk(bsw_engine* eng, DevCtx& c, Chunk& ch)
{
    for (int s = 0; s < NSTREAMS; ++s) {
        CUDA_TRY(cudaEventRecord(c.ev_join[s], c.st[s]));
        CUDA_TRY(cudaStreamWaitEvent(c.st_d2h, c.ev_join[s], 0));
    }
    CUDA_TRY(cudaMemcpyAsync(c.res.h + ch.s0, c.res.d + ch.s0, sizeof(int4) * (size_t)(ch.s1 - ch.s0), cudaMemcpyDeviceToHost, c.st_d2h));
    CUDA_TRY(cudaEventRecord(ch.ev_done, c.st_d2h));
    eng->stats.d2h_bytes += (int64_t)(sizeof(int4) * (size_t)(ch.s1 - ch.s0));
    return BSW_OK;
}

inline void write_result(SeqPair& sp, const int4 v)
{
    sp.score = (int16_t)(v.x & 0xffff);  sp.qle = (int16_t)(v.x >> 16);
    sp.tle = (int16_t)(v.y & 0xffff);    sp.gtle = (int16_t)(v.y >> 16);
    sp.gscore = (int16_t)(v.z & 0xffff); sp.max_off = (int16_t)(v.z >> 16);
}

void scatter_chunk(bsw_engine* eng, DevCtx& c, Chunk& ch, SeqPair* pairs)
{
    const uint32_t* idx = c.idx.data();
    const int4* r = c.res.h;
    eng->pool->for_range(ch.s1 - ch.s0, 4096, [&](int64_t b, int64_t e, int) {
        b += ch.s0; e += ch.s0;
        for (int64_t s = b; s < e; ++s) {
            if (s + 8 < e) __builtin_prefetch(&pairs[idx[s + 8]].score, 1, 0);
            if (!c.hasn[s]) write_result(pairs[idx[s]], r[s]);
        }
    });
    ch.scattered = true;
}

// byte-staged pairs: short pairs with N + all long pairs.  Staging (host) and H2D.
int stage_bytes(bsw_engine* eng, DevCtx& c, const SeqPair* pairs, const uint8_t* seq_ref, const uint8_t* seq_qer)
{
    (void)pairs;
    c.pos_n.clear();
    c.qmax_bytes_short = 0;
    for (int64_t s = 0; s < c.n_short; ++s)
        if (c.hasn[s]) { c.pos_n.push_back((uint32_t)s); c.qmax_bytes_short = std::max<int>(c.qmax_bytes_short, c.len2[s]); }
    c.n_bytes_short = (int64_t)c.pos_n.size();
    for (int64_t s = c.n_short; s < c.n; ++s) c.pos_n.push_back((uint32_t)s);
    const size_t nb = c.pos_n.size();
    if (nb == 0) return BSW_OK;
    if (int rc = ensure(eng, c.meta_n, nb)) return rc;
    if (int rc = ensure(eng, c.res_n, nb)) return rc;
    uint64_t qo = 0, to = 0;
    for (size_t k = 0; k < nb; ++k) {
        const uint32_t s = c.pos_n[k];
        c.meta_n.h[k] = make_int4((int)qo, (int)to, c.len2[s] | (c.len1[s] << 16), c.h0[s]);
        qo += ((uint64_t)c.len2[s] + 7) & ~3ull; to += ((uint64_t)c.len1[s] + 7) & ~3ull;   // 4-byte aligned, >= 4 B slack
    }
    if (qo > 0x7fffffffull || to > 0x7fffffffull) { eng->err = "too many byte-staged bases in one batch"; return BSW_ERR_PARAM; }
    c.qb_bytes = qo; c.tb_bytes = to;
    if (int rc = ensure(eng, c.qb, c.qb_bytes + 16)) return rc;
    if (int rc = ensure(eng, c.tb, c.tb_bytes + 16)) return rc;
    eng->pool->for_range((int64_t)nb, 256, [&](int64_t b, int64_t e, int) {
        for (int64_t k = b; k < e; ++k) {
            const uint32_t s = c.pos_n[k], i = c.idx[s];
            memcpy(c.qb.h + c.meta_n.h[k].x, seq_qer + eng->sb.offq[i], (size_t)c.len2[s]);
            memcpy(c.tb.h + c.meta_n.h[k].y, seq_ref + eng->sb.offr[i], (size_t)c.len1[s]);
        }
    });
    if (c.n_long > 0) {
        const int qmax = c.len2[c.n - 1];
        c.long_stride = (qmax + 12) & ~3;
        int64_t blocks = std::min<int64_t>((c.n_long + LONG_WARPS - 1) / LONG_WARPS, (int64_t)c.sms * 8);
        const int64_t cap_words = (int64_t)(256ll << 20) / 4;       // <= 256 MB of eh rows
        blocks = std::max<int64_t>(1, std::min(blocks, cap_words / ((int64_t)c.long_stride * LONG_WARPS)));
        c.long_blocks = (int)blocks;
        if (int rc = ensure(eng, c.scratch, (size_t)blocks * LONG_WARPS * c.long_stride, false)) return rc;
    }
    cudaStream_t st = c.st_copy;
    CUDA_TRY(cudaMemcpyAsync(c.meta_n.d, c.meta_n.h, sizeof(int4) * nb, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(c.qb.d, c.qb.h, c.qb_bytes, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(c.tb.d, c.tb.h, c.tb_bytes, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaEventRecord(c.ev_a, st));
    eng->stats.h2d_bytes += (int64_t)(sizeof(int4) * nb + c.qb_bytes + c.tb_bytes);
    return BSW_OK;
}

int launch_bytes(bsw_engine* eng, DevCtx& c, bool wait_h2d)
{
    if (c.pos_n.empty()) return BSW_OK;
    cudaStream_t st = c.st[0];
    if (wait_h2d) CUDA_TRY(cudaStreamWaitEvent(st, c.ev_a, 0));
    if (c.n_bytes_short > 0) {
        const int grid = (int)((c.n_bytes_short + SHORT_BLOCK - 1) / SHORT_BLOCK);
        const int qstride = stride_for(c.qmax_bytes_short);
        bsw_short_kernel<SHORT_BLOCK, true><<<grid, SHORT_BLOCK, short_smem_bytes(qstride), st>>>(
            c.meta_n.d, reinterpret_cast<const uint32_t*>(c.qb.d), reinterpret_cast<const uint32_t*>(c.tb.d),
            c.res_n.d, 0, (int)c.n_bytes_short, qstride, eng->kp, c.d_cells);
        eng->stats.kernel_launches++;
    }
    if (c.n_long > 0) {
        CUDA_TRY(cudaMemsetAsync(c.d_queue, 0, sizeof(unsigned int), st));
        bsw_long_kernel<<<c.long_blocks, LONG_WARPS * 32, 0, st>>>(
            c.meta_n.d + c.n_bytes_short, c.qb.d, c.tb.d, c.res_n.d + c.n_bytes_short, (int)c.n_long, eng->kp,
            c.scratch.d, c.long_stride, c.d_queue, c.d_cells);
        eng->stats.kernel_launches++;
    }
    CUDA_TRY(cudaGetLastError());
    return BSW_OK;
}

int d2h_bytes(bsw_engine* eng, DevCtx& c)
{
    if (c.pos_n.empty()) return BSW_OK;
    CUDA_TRY(cudaEventRecord(c.ev_join[0], c.st[0]));
    CUDA_TRY(cudaStreamWaitEvent(c.st_d2h, c.ev_join[0], 0));
    CUDA_TRY(cudaMemcpyAsync(c.res_n.h, c.res_n.d, sizeof(int4) * c.pos_n.size(), cudaMemcpyDeviceToHost, c.st_d2h));
    eng->stats.d2h_bytes += (int64_t)(sizeof(int4) * c.pos_n.size());
    return BSW_OK;
}

void scatter_bytes(bsw_engine* eng, DevCtx& c, SeqPair* pairs)
{
    eng->pool->for_range((int64_t)c.pos_n.size(), 4096, [&](int64_t b, int64_t e, int) {
        for (int64_t k = b; k < e; ++k) write_result(pairs[c.idx[c.pos_n[k]]], c.res_n.h[k]);
    });
}

// Splits the sorted batch across the engine's devices (views copied per shard).
int build_shards(bsw_engine* eng)
{
    SortedBatch& sb = eng->sb;
    const int ndev = (int)eng->devs.size();
    if (ndev == 1) {
        DevCtx& c = eng->devs[0];
        c.n = sb.n;
        c.idx.swap(sb.idx); c.len2.swap(sb.len2); c.len1.swap(sb.len1); c.h0.swap(sb.h0);
        return BSW_OK;
    }
    std::vector<std::vector<int64_t>> blocks_of;
    partition_blocks(sb, eng->w, ndev, blocks_of);
    for (int g = 0; g < ndev; ++g) {
        DevCtx& c = eng->devs[g];
        size_t cnt = 0;
        for (int64_t b : blocks_of[g]) cnt += (size_t)(std::min(sb.n, b * 1024 + 1024) - b * 1024);
        c.idx.reserve(cnt); c.len2.reserve(cnt); c.len1.reserve(cnt); c.h0.reserve(cnt);
        size_t pos = 0;
        for (int64_t b : blocks_of[g]) {
            const int64_t lo = b * 1024, hi = std::min(sb.n, lo + 1024);
            const size_t m = (size_t)(hi - lo);
            memcpy(c.idx.data() + pos, sb.idx.data() + lo, m * sizeof(uint32_t));
            memcpy(c.len2.data() + pos, sb.len2.data() + lo, m * sizeof(uint16_t));
            memcpy(c.len1.data() + pos, sb.len1.data() + lo, m * sizeof(uint16_t));
            memcpy(c.h0.data() + pos, sb.h0.data() + lo, m * sizeof(uint16_t));
            pos += m;
        }
        c.n = (int64_t)cnt;
    }
    return BSW_OK;
}

int begin_batch(bsw_engine* eng, const SeqPair* pairs, const uint8_t* seq_ref, const uint8_t* seq_qer,
                int64_t n, int32_t w)
{
    eng->err.clear();
    eng->staged = false; eng->ran = false;
    if (n < 0 || w < 0 || (n > 0 && (!pairs || !seq_ref || !seq_qer))) { eng->err = "bad arguments"; return BSW_ERR_PARAM; }
    if (n > 0x7fffffff) { eng->err = "more than 2^31-1 pairs per call"; return BSW_ERR_PARAM; }
    bsw_stats& S = eng->stats;
    memset(&S, 0, sizeof(S));
    S.pairs = n;
    eng->n = n; eng->w = w; eng->kp.w = w;
    const double t0 = now_ms();
    build_sorted_batch(*eng->pool, pairs, n, eng->p.match, eng->sb);
    if (!eng->sb.domain_ok) {
        eng->err = "pair outside the domain: need 1<=len1,len2<=32767, h0>=1, h0+len2*match<=32767, offsets>=0 (bandedSWA.h:84, SURVEY 8b)";
        return BSW_ERR_DOMAIN;
    }
    S.cells_nominal = eng->sb.cells_nominal;
    if (int rc = build_shards(eng)) return rc;
    S.ms_sort = now_ms() - t0;
    return BSW_OK;
}

} // namespace

extern "C" {

const char* bsw_version(void) { return "bsw_b200 0.2 sm_100a"; }

void bsw_default_params(bsw_params* p)
{
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->o_del = p->o_ins = 6; p->e_del = p->e_ins = 1;   // main_banded.cpp:51-52
    p->zdrop = 100; p->end_bonus = 5;                   // main_banded.cpp:250
    p->match = 1; p->mismatch = 4; p->ambig = -1;       // main_banded.cpp:49-50,53
    p->zdrop_mode = BSW_ZDROP_VECTOR;
}

const char* bsw_last_error(const bsw_engine* eng)
{
    if (eng) return eng->err.c_str();
    std::lock_guard<std::mutex> g(g_err_mutex);
    static thread_local std::string copy;
    copy = g_create_error;
    return copy.c_str();
}

bsw_engine* bsw_create(const bsw_params* params, int* err)
{
    auto fail = [&](int code, const std::string& msg) -> bsw_engine* {
        { std::lock_guard<std::mutex> g(g_err_mutex); g_create_error = msg; }
        if (err) *err = code;
        return nullptr;
    };
    if (!params) return fail(BSW_ERR_PARAM, "params == NULL");
    std::string why;
    if (validate_params(params, why) != BSW_OK) return fail(BSW_ERR_PARAM, why);

    int ndev_avail = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev_avail);
    if (ce != cudaSuccess || ndev_avail < 1)
        return fail(BSW_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(ce) +
                                  " (this engine has no CPU fallback)");
    bsw_engine* eng = new bsw_engine();
    eng->p = *params;
    KParams& k = eng->kp;
    k.match = params->match; k.mismatch_neg = -params->mismatch; k.ambig = params->ambig;
    k.o_del = params->o_del; k.e_del = params->e_del; k.o_ins = params->o_ins; k.e_ins = params->e_ins;
    k.oe_del = k.o_del + k.e_del; k.oe_ins = k.o_ins + k.e_ins;
    k.zdrop = params->zdrop; k.end_bonus = params->end_bonus; k.zmode = params->zdrop_mode;
    k.mx = std::max(std::max(k.match, k.mismatch_neg), k.ambig);
    k.w = 0; k.kone = 1;
    eng->pool.reset(new ThreadPool(auto_threads(params->host_threads)));
    eng->short_max = params->long_min_qlen > 0 ? std::min(params->long_min_qlen - 1, SHORT_MAX_QLEN) : SHORT_MAX_QLEN;
    memset(&eng->stats, 0, sizeof(eng->stats));

    std::vector<int> ids;
    if (params->n_devices == 0) { int cur = 0; cudaGetDevice(&cur); ids.push_back(cur); }
    else for (int i = 0; i < params->n_devices; ++i) ids.push_back(params->devices[i]);
    for (int id : ids)
        if (id < 0 || id >= ndev_avail) { delete eng; return fail(BSW_ERR_PARAM, "device ordinal out of range"); }
    eng->devs.resize(ids.size());
    for (size_t i = 0; i < ids.size(); ++i) {
        DevCtx& c = eng->devs[i];
        c.dev = ids[i];
        bool ok = cudaSetDevice(c.dev) == cudaSuccess;
        cudaDeviceProp prop{};
        ok = ok && cudaGetDeviceProperties(&prop, c.dev) == cudaSuccess;
        if (ok && prop.major != 10) {
            delete eng;
            return fail(BSW_ERR_CUDA, std::string("device ") + prop.name +
                                      " is not sm_100: this library carries sm_100a code only");
        }
        c.sms = prop.multiProcessorCount;
        ok = ok && cudaStreamCreateWithFlags(&c.st_copy, cudaStreamNonBlocking) == cudaSuccess;
        ok = ok && cudaStreamCreateWithFlags(&c.st_d2h, cudaStreamNonBlocking) == cudaSuccess;
        for (int s = 0; ok && s < NSTREAMS; ++s) {
            ok = cudaStreamCreateWithFlags(&c.st[s], cudaStreamNonBlocking) == cudaSuccess;
            ok = ok && cudaEventCreateWithFlags(&c.ev_join[s], cudaEventDisableTiming) == cudaSuccess;
        }
        for (int k2 = 0; ok && k2 < MAX_CHUNKS; ++k2) {
            ok = cudaEventCreateWithFlags(&c.chunks[k2].ev_h2d, cudaEventDisableTiming) == cudaSuccess;
            ok = ok && cudaEventCreateWithFlags(&c.chunks[k2].ev_done, cudaEventDisableTiming) == cudaSuccess;
        }
        ok = ok && cudaEventCreate(&c.ev_a) == cudaSuccess && cudaEventCreate(&c.ev_b) == cudaSuccess;
        ok = ok && cudaEventCreate(&c.ev_k0) == cudaSuccess && cudaEventCreate(&c.ev_k1) == cudaSuccess;
        ok = ok && cudaMalloc((void**)&c.d_cells, sizeof(unsigned long long)) == cudaSuccess;
        ok = ok && cudaMalloc((void**)&c.d_queue, sizeof(unsigned int)) == cudaSuccess;
        ok = ok && cudaHostAlloc((void**)&c.h_cells, sizeof(unsigned long long), cudaHostAllocDefault) == cudaSuccess;
        if (!ok) {
            std::string m = std::string("device setup failed: ") + cudaGetErrorString(cudaGetLastError());
            bsw_destroy(eng);
            return fail(BSW_ERR_CUDA, m);
        }
    }
    if (err) *err = BSW_OK;
    return eng;
}

void bsw_destroy(bsw_engine* eng)
{
    if (!eng) return;
    for (DevCtx& c : eng->devs) {
        cudaSetDevice(c.dev);
        cudaDeviceSynchronize();
        for (cudaStream_t s : {c.st_copy, c.st_d2h}) if (s) cudaStreamDestroy(s);
        for (int s = 0; s < NSTREAMS; ++s) {
            if (c.st[s]) cudaStreamDestroy(c.st[s]);
            if (c.ev_join[s]) cudaEventDestroy(c.ev_join[s]);
        }
        for (Chunk& ch : c.chunks) {
            if (ch.ev_h2d) cudaEventDestroy(ch.ev_h2d);
            if (ch.ev_done) cudaEventDestroy(ch.ev_done);
        }
        for (cudaEvent_t e : {c.ev_a, c.ev_b, c.ev_k0, c.ev_k1})
            if (e) cudaEventDestroy(e);
        release(c.meta); release(c.res); release(c.meta_n); release(c.res_n); release(c.q); release(c.t);
        release(c.qb); release(c.tb); release(c.scratch);
        if (c.d_cells) cudaFree(c.d_cells);
        if (c.d_queue) cudaFree(c.d_queue);
        if (c.h_cells) cudaFreeHost(c.h_cells);
    }
    delete eng;
}

int bsw_get_stats(const bsw_engine* eng, bsw_stats* out)
{
    if (!eng || !out) return BSW_ERR_PARAM;
    *out = eng->stats;
    return BSW_OK;
}

// ------------------------------------------------------------------------------------------
// resident form: stage (host -> HBM), run (kernels only, repeatable), fetch (HBM -> SeqPair[])
// ------------------------------------------------------------------------------------------
int bsw_stage(bsw_engine* eng, const SeqPair* pairs, const uint8_t* seq_ref, const uint8_t* seq_qer,
              int64_t n, int32_t w)
{
    if (!eng) return BSW_ERR_PARAM;
    const double t_begin = now_ms();
    if (int rc = begin_batch(eng, pairs, seq_ref, seq_qer, n, w)) return rc;
    bsw_stats& S = eng->stats;
    for (DevCtx& c : eng->devs) {
        if (int rc = prepare_shard(eng, c, false)) return rc;
        if (c.n == 0) continue;
        const double tp = now_ms();
        CUDA_TRY(cudaEventRecord(c.ev_a, c.st_copy));
        for (int k = 0; k < c.nchunks; ++k) {
            pack_chunk(eng, c, c.chunks[k], pairs, seq_ref, seq_qer);
            if (int rc = h2d_chunk(eng, c, c.chunks[k])) return rc;
        }
        CUDA_TRY(cudaEventRecord(c.ev_b, c.st_copy));
        S.ms_pack += now_ms() - tp;
    }
    for (DevCtx& c : eng->devs) {
        if (c.n == 0) continue;
        CUDA_TRY(cudaSetDevice(c.dev));
        CUDA_TRY(cudaStreamSynchronize(c.st_copy));
        float ms = 0; CUDA_TRY(cudaEventElapsedTime(&ms, c.ev_a, c.ev_b));
        S.ms_h2d = std::max(S.ms_h2d, (double)ms);
        const double tp = now_ms();
        if (int rc = stage_bytes(eng, c, pairs, seq_ref, seq_qer)) return rc;
        CUDA_TRY(cudaStreamSynchronize(c.st_copy));
        S.ms_pack += now_ms() - tp;
        S.n_short += (int32_t)c.n_short;
        S.n_long += (int32_t)c.n_long;
    }
    eng->staged = true;
    S.ms_total = now_ms() - t_begin;
    return BSW_OK;
}

int bsw_run_staged(bsw_engine* eng)
{
    if (!eng) return BSW_ERR_PARAM;
    if (!eng->staged) { eng->err = "bsw_run_staged before bsw_stage"; return BSW_ERR_STATE; }
    bsw_stats& S = eng->stats;
    S.kernel_launches = 0; S.ms_kernel = 0; S.cells_effective = 0;
    for (DevCtx& c : eng->devs) {
        if (c.n == 0) continue;
        CUDA_TRY(cudaSetDevice(c.dev));
        CUDA_TRY(cudaMemsetAsync(c.d_cells, 0, sizeof(unsigned long long), c.st[0]));
        CUDA_TRY(cudaEventRecord(c.ev_k0, c.st[0]));
        for (int s = 1; s < NSTREAMS; ++s) CUDA_TRY(cudaStreamWaitEvent(c.st[s], c.ev_k0, 0));
        for (int k = c.nchunks - 1; k >= 0; --k)
            if (int rc = launch_chunk(eng, c, c.chunks[k], false)) return rc;
        for (int s = 1; s < NSTREAMS; ++s) {
            CUDA_TRY(cudaEventRecord(c.ev_join[s], c.st[s]));
            CUDA_TRY(cudaStreamWaitEvent(c.st[0], c.ev_join[s], 0));
        }
        if (int rc = launch_bytes(eng, c, false)) return rc;
        CUDA_TRY(cudaEventRecord(c.ev_k1, c.st[0]));
        CUDA_TRY(cudaMemcpyAsync(c.h_cells, c.d_cells, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c.st[0]));
    }
    for (DevCtx& c : eng->devs) {
        if (c.n == 0) continue;
        CUDA_TRY(cudaSetDevice(c.dev));
        CUDA_TRY(cudaStreamSynchronize(c.st[0]));
        float ms = 0; CUDA_TRY(cudaEventElapsedTime(&ms, c.ev_k0, c.ev_k1));
        S.ms_kernel = std::max(S.ms_kernel, (double)ms);
        S.cells_effective += (int64_t)*c.h_cells;
    }
    eng->ran = true;
    return BSW_OK;
}

int bsw_fetch(bsw_engine* eng, SeqPair* pairs, int64_t n)
{
    if (!eng) return BSW_ERR_PARAM;
    if (!eng->ran) { eng->err = "bsw_fetch before bsw_run_staged"; return BSW_ERR_STATE; }
    if (n != eng->n || (n > 0 && !pairs)) { eng->err = "bsw_fetch: pair count differs from the staged batch"; return BSW_ERR_PARAM; }
    bsw_stats& S = eng->stats;
    S.d2h_bytes = 0; S.ms_d2h = 0; S.ms_scatter = 0;
    for (DevCtx& c : eng->devs) {
        if (c.n == 0) continue;
        CUDA_TRY(cudaSetDevice(c.dev));
        CUDA_TRY(cudaEventRecord(c.ev_a, c.st_d2h));
        for (int k = 0; k < c.nchunks; ++k) if (int rc = d2h_chunk(eng, c, c.chunks[k])) return rc;
        if (int rc = d2h_bytes(eng, c)) return rc;
        CUDA_TRY(cudaEventRecord(c.ev_b, c.st_d2h));
    }
    for (DevCtx& c : eng->devs) {
        if (c.n == 0) continue;
        CUDA_TRY(cudaSetDevice(c.dev));
        CUDA_TRY(cudaStreamSynchronize(c.st_d2h));
        float ms = 0; CUDA_TRY(cudaEventElapsedTime(&ms, c.ev_a, c.ev_b));
        S.ms_d2h = std::max(S.ms_d2h, (double)ms);
        const double t0 = now_ms();
        for (int k = 0; k < c.nchunks; ++k) scatter_chunk(eng, c, c.chunks[k], pairs);
        scatter_bytes(eng, c, pairs);
        S.ms_scatter += now_ms() - t0;
    }
    return BSW_OK;
}

// ------------------------------------------------------------------------------------------
// the hot path: host buffers in, results in place.  Chunk pipeline (see the file header).
// ------------------------------------------------------------------------------------------
int bsw_extend(bsw_engine* eng, SeqPair* pairs, const uint8_t* seq_ref, const uint8_t* seq_qer,
               int64_t n, int32_t w)
{
    if (!eng) return BSW_ERR_PARAM;
    const double t_begin = now_ms();
    if (int rc = begin_batch(eng, pairs, seq_ref, seq_qer, n, w)) return rc;
    bsw_stats& S = eng->stats;
    for (DevCtx& c : eng->devs) {
        if (int rc = prepare_shard(eng, c, true)) return rc;
        if (c.n == 0) continue;
        CUDA_TRY(cudaMemsetAsync(c.d_cells, 0, sizeof(unsigned long long), c.st_copy));
        CUDA_TRY(cudaEventRecord(c.ev_k0, c.st_copy));
        for (int s = 0; s < NSTREAMS; ++s) CUDA_TRY(cudaStreamWaitEvent(c.st[s], c.ev_k0, 0));
    }
    // longest chunk first on every device; the devices' pipelines are interleaved chunk by chunk
    int max_chunks = 0;
    for (DevCtx& c : eng->devs) max_chunks = std::max(max_chunks, c.nchunks);
    for (int step = 0; step < max_chunks; ++step) {
        for (DevCtx& c : eng->devs) {
            const int k = c.nchunks - 1 - step;
            if (k < 0) continue;
            CUDA_TRY(cudaSetDevice(c.dev));
            double t0 = now_ms();
            pack_chunk(eng, c, c.chunks[k], pairs, seq_ref, seq_qer);
            S.ms_pack += now_ms() - t0;
            if (int rc = h2d_chunk(eng, c, c.chunks[k])) return rc;
            if (int rc = launch_chunk(eng, c, c.chunks[k], true)) return rc;
            if (int rc = d2h_chunk(eng, c, c.chunks[k])) return rc;
            // scatter whatever already came back while the GPU works on this chunk
            t0 = now_ms();
            for (int j = c.nchunks - 1; j > k; --j) {
                Chunk& done = c.chunks[j];
                if (!done.scattered && cudaEventQuery(done.ev_done) == cudaSuccess) scatter_chunk(eng, c, done, pairs);
            }
            S.ms_scatter += now_ms() - t0;
        }
    }
    for (DevCtx& c : eng->devs) {
        if (c.n == 0) continue;
        CUDA_TRY(cudaSetDevice(c.dev));
        const double t0 = now_ms();
        if (int rc = stage_bytes(eng, c, pairs, seq_ref, seq_qer)) return rc;
        S.ms_pack += now_ms() - t0;
        if (int rc = launch_bytes(eng, c, true)) return rc;
        if (int rc = d2h_bytes(eng, c)) return rc;
        for (int s = 0; s < NSTREAMS; ++s) {
            CUDA_TRY(cudaEventRecord(c.ev_join[s], c.st[s]));
            CUDA_TRY(cudaStreamWaitEvent(c.st_d2h, c.ev_join[s], 0));
        }
        CUDA_TRY(cudaEventRecord(c.ev_k1, c.st_d2h));
        CUDA_TRY(cudaMemcpyAsync(c.h_cells, c.d_cells, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c.st_d2h));
        S.n_short += (int32_t)c.n_short;
        S.n_long += (int32_t)c.n_long;
    }
    for (DevCtx& c : eng->devs) {
        if (c.n == 0) continue;
        CUDA_TRY(cudaSetDevice(c.dev));
        for (int k = c.nchunks - 1; k >= 0; --k) {
            Chunk& ch = c.chunks[k];
            if (ch.scattered) continue;
            CUDA_TRY(cudaEventSynchronize(ch.ev_done));
            const double t0 = now_ms();
            scatter_chunk(eng, c, ch, pairs);
            S.ms_scatter += now_ms() - t0;
        }
        CUDA_TRY(cudaStreamSynchronize(c.st_d2h));
        const double t0 = now_ms();
        scatter_bytes(eng, c, pairs);
        S.ms_scatter += now_ms() - t0;
        float ms = 0; CUDA_TRY(cudaEventElapsedTime(&ms, c.ev_k0, c.ev_k1));
        S.ms_kernel = std::max(S.ms_kernel, (double)ms);      // first H2D .. last kernel, copies overlapped
        S.cells_effective += (int64_t)*c.h_cells;
    }
    S.ms_total = now_ms() - t_begin;
    return BSW_OK;
}

double bsw_measure_int_peak(bsw_engine* eng)
{
    if (!eng || eng->devs.empty()) return 0.0;
    DevCtx& c = eng->devs[0];
    if (cudaSetDevice(c.dev) != cudaSuccess) return 0.0;
    const int threads = 256, blocks = c.sms * 8, iters = 4096;
    int* d_out = nullptr;
    if (cudaMalloc((void**)&d_out, sizeof(int) * (size_t)threads * blocks) != cudaSuccess) return 0.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, c.st[0]);
        bsw_int_peak_kernel<<<blocks, threads, 0, c.st[0]>>>(d_out, iters, 12345 + rep);
        cudaEventRecord(e1, c.st[0]);
        if (cudaStreamSynchronize(c.st[0]) != cudaSuccess) { best = 0.0; break; }
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        const double ops = (double)threads * blocks * (double)iters * 64.0;
        if (rep > 0 && ms > 0) best = std::max(best, ops / (ms * 1e-3));
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_out);
    return best;
}

} // extern "C"
