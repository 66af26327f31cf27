// bsw_engine.cu -- engine object, host pipeline and kernel launches behind the C ABI (include/bsw.h).
//
// Stands in for BandedPairWiseSW's batch wrapper smithWatermanBatchWrapper16
// (benchmarks/bsw/bandedSWA.cpp:1150-1431): where the reference pads to the SIMD width, sorts by
// len1, transposes AoS->SoA per 16 pairs and calls the AVX kernel, this engine buckets by
// (len2, len1, h0), packs sequences to 2 bits/base straight into pinned staging, copies
// asynchronously to HBM, launches the sm_100a kernels per length class and scatters the six
// result fields back into the caller's SeqPair[] in input order.
#include "bsw_common.h"
#include "bsw_kernels.cuh"
#include <cstdio>
#include <cstring>
#include <string>
#include <mutex>

namespace bsw {
void bucket_order(const SeqPair* pairs, int64_t n, int64_t* order, int nthreads);
}
using namespace bsw;

namespace {

constexpr int SHORT_BLOCK = 64;           // threads (= pairs) per block of the short kernel
constexpr int SHORT_MAX_QLEN = 880;       // (qlen+1) * SHORT_BLOCK * 4 B must fit 227 KB
constexpr int NSTREAMS = 4;

std::string g_create_error;
std::mutex g_err_mutex;

struct Launch {
    int first, count;     // range in the shard's processing order
    int qstride;          // shared-memory rows per thread (>= qmax + 1)
    bool bytes;           // N-containing pairs: byte sequences
};

template <class T>
struct Buf {              // grow-only device + pinned-host buffer pair
    T* d = nullptr; T* h = nullptr; size_t cap = 0;
};

struct DevCtx {
    int dev = 0;
    cudaStream_t st[NSTREAMS] = {};
    cudaEvent_t ev_h2d0{}, ev_h2d1{}, ev_k0{}, ev_k1{}, ev_d2h0{}, ev_d2h1{}, ev_join[NSTREAMS] = {};
    Buf<int4> meta, res, meta_n;
    Buf<uint32_t> q, t;
    Buf<uint8_t> qb, tb;
    Buf<int> pos_n;
    unsigned long long* d_cells = nullptr;
    unsigned long long* h_cells = nullptr;
    unsigned int* d_queue = nullptr;      // work queue of the long-pair kernel
    Buf<uint32_t> scratch;                // eh[] rows of the long-pair kernel (device only)
    int64_t n_short = 0;                  // sorted positions [0, n_short) go to the short kernel
    int64_t n_bytes_short = 0;            // byte-staged list: [0, n_bytes_short) short pairs with N,
    int64_t n_long = 0;                   //                   [n_bytes_short, +n_long) long pairs
    int long_stride = 0, long_blocks = 0;
    // staged shard
    int64_t first = 0, n = 0, n_bytes_pairs = 0;
    size_t q_words = 0, t_words = 0, qb_bytes = 0, tb_bytes = 0;
    std::vector<Launch> plan;
    bool attr_set = false;
};

} // namespace

struct bsw_engine {
    bsw_params p;
    KParams kp;
    std::vector<DevCtx> devs;
    std::string err;
    bsw_stats stats;
    int nthreads = 1;
    int short_max = 0;                   // longest query the short kernel takes
    // staged batch
    bool staged = false, ran = false;
    int64_t n = 0;
    int32_t w = 0;
    std::vector<int64_t> order;          // processing order -> caller index
    std::vector<int64_t> shard_begin;
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
// Returns true if a code > 3 (N) was seen; such bases are packed as 0.
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

const int kStrideSteps[] = {9, 17, 25, 33, 41, 49, 57, 65, 73, 81, 89, 97, 105, 113, 121, 129, 145, 153, 161,
                            177, 193, 209, 225, 241, 257, 273, 289, 305, 321, 353, 385, 417, 449, 513,
                            577, 641, 705, 769, 833, SHORT_MAX_QLEN + 1};

inline int stride_for(int qmax)
{
    for (int s : kStrideSteps) if (s >= qmax + 1) return s;
    return -1;
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

} // namespace

extern "C" {

const char* bsw_version(void) { return "bsw_b200 0.1 sm_100a"; }

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
    k.w = 0;
    eng->nthreads = auto_threads(params->host_threads);
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
        for (int s = 0; ok && s < NSTREAMS; ++s) {
            ok = cudaStreamCreateWithFlags(&c.st[s], cudaStreamNonBlocking) == cudaSuccess;
            ok = ok && cudaEventCreateWithFlags(&c.ev_join[s], cudaEventDisableTiming) == cudaSuccess;
        }
        ok = ok && cudaEventCreate(&c.ev_h2d0) == cudaSuccess && cudaEventCreate(&c.ev_h2d1) == cudaSuccess;
        ok = ok && cudaEventCreate(&c.ev_k0) == cudaSuccess && cudaEventCreate(&c.ev_k1) == cudaSuccess;
        ok = ok && cudaEventCreate(&c.ev_d2h0) == cudaSuccess && cudaEventCreate(&c.ev_d2h1) == cudaSuccess;
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
        for (int s = 0; s < NSTREAMS; ++s) {
            if (c.st[s]) { cudaStreamSynchronize(c.st[s]); cudaStreamDestroy(c.st[s]); }
            if (c.ev_join[s]) cudaEventDestroy(c.ev_join[s]);
        }
        for (cudaEvent_t e : {c.ev_h2d0, c.ev_h2d1, c.ev_k0, c.ev_k1, c.ev_d2h0, c.ev_d2h1})
            if (e) cudaEventDestroy(e);
        release(c.meta); release(c.res); release(c.meta_n); release(c.q); release(c.t);
        release(c.qb); release(c.tb); release(c.pos_n);
        if (c.d_cells) cudaFree(c.d_cells);
        if (c.d_queue) cudaFree(c.d_queue);
        release(c.scratch);
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
// stage: validate, bucket, partition across devices, pack into pinned staging, async H2D
// ------------------------------------------------------------------------------------------
int bsw_stage(bsw_engine* eng, const SeqPair* pairs, const uint8_t* seq_ref, const uint8_t* seq_qer,
              int64_t n, int32_t w)
{
    if (!eng) return BSW_ERR_PARAM;
    eng->err.clear();
    eng->staged = false; eng->ran = false;
    if (n < 0 || w < 0 || (n > 0 && (!pairs || !seq_ref || !seq_qer))) { eng->err = "bad arguments"; return BSW_ERR_PARAM; }
    if (n > 0x7fffffff) { eng->err = "more than 2^31-1 pairs per call"; return BSW_ERR_PARAM; }
    const double t_begin = now_ms();
    bsw_stats& S = eng->stats;
    memset(&S, 0, sizeof(S));
    S.pairs = n;
    eng->n = n; eng->w = w;
    eng->kp.w = w;
    const int nt = eng->nthreads;
    const int ndev = (int)eng->devs.size();

    // ---- domain check + nominal cells
    {
        std::vector<int64_t> nominal((size_t)nt + 1, 0);
        std::vector<int> bad((size_t)nt + 1, 0);
        const int match = eng->p.match;
        parallel_chunks(n, 1 << 14, nt, [&](int64_t b, int64_t e, int t) {
            int64_t acc = 0; int bd = 0;
            for (int64_t k = b; k < e; ++k) {
                const SeqPair& sp = pairs[k];
                if (sp.len1 < 1 || sp.len1 > 32767 || sp.len2 < 1 || sp.len2 > 32767 || sp.h0 < 1 ||
                    (int64_t)sp.h0 + (int64_t)sp.len2 * match > 32767 || sp.idr < 0 || sp.idq < 0) bd = 1;
                acc += (int64_t)sp.len1 * sp.len2;
            }
            nominal[t] += acc; bad[t] |= bd;
        });
        for (int t = 0; t <= nt; ++t) {
            S.cells_nominal += nominal[t];
            if (bad[t]) {
                eng->err = "pair outside the domain: need 1<=len1,len2<=32767, h0>=1, h0+len2*match<=32767 (bandedSWA.h:84, SURVEY 8b)";
                return BSW_ERR_DOMAIN;
            }
        }
    }

    // ---- bucket + partition
    double t0 = now_ms();
    eng->order.resize((size_t)n);
    eng->shard_begin.assign((size_t)ndev + 1, 0);
    if (ndev == 1) {
        bucket_order(pairs, n, eng->order.data(), nt);
        eng->shard_begin[1] = n;
    } else {
        int rc = bsw_partition(pairs, n, w, ndev, eng->order.data(), eng->shard_begin.data());
        if (rc != BSW_OK) { eng->err = "partition failed"; return rc; }
    }
    S.ms_sort = now_ms() - t0;

    // ---- per device: offsets, pack, H2D
    for (int d = 0; d < ndev; ++d) {
        DevCtx& c = eng->devs[d];
        CUDA_TRY(cudaSetDevice(c.dev));
        if (int rc = set_kernel_attrs(eng, c)) return rc;
        c.first = eng->shard_begin[d];
        c.n = eng->shard_begin[d + 1] - c.first;
        c.plan.clear();
        c.n_bytes_pairs = 0; c.q_words = c.t_words = c.qb_bytes = c.tb_bytes = 0;
        if (c.n == 0) continue;
        const int64_t* ord = eng->order.data() + c.first;
        double tp0 = now_ms();
        if (int rc = ensure(eng, c.meta, (size_t)c.n)) return rc;
        if (int rc = ensure(eng, c.res, (size_t)c.n)) return rc;
        // word offsets (serial prefix over the sorted order); queries longer than the short
        // kernel's shared-memory limit sit at the end of the order and are staged as bytes only
        std::vector<uint32_t> qoff((size_t)c.n + 1), toff((size_t)c.n + 1);
        {
            uint64_t qo = 0, to = 0;
            c.n_short = c.n;
            for (int64_t s = 0; s < c.n; ++s) {
                const SeqPair& sp = pairs[ord[s]];
                qoff[s] = (uint32_t)qo; toff[s] = (uint32_t)to;
                if (sp.len2 > eng->short_max) { if (c.n_short == c.n) c.n_short = s; continue; }
                qo += (uint64_t)(sp.len2 + 15) >> 4; to += (uint64_t)(sp.len1 + 15) >> 4;
            }
            if (qo > 0xffffffffull || to > 0xffffffffull) { eng->err = "batch too large for 32-bit word offsets"; return BSW_ERR_PARAM; }
            qoff[c.n] = (uint32_t)qo; toff[c.n] = (uint32_t)to;
            c.q_words = qo; c.t_words = to;
        }
        if (int rc = ensure(eng, c.q, c.q_words + 4)) return rc;
        if (int rc = ensure(eng, c.t, c.t_words + 4)) return rc;
        std::vector<uint8_t> hasn((size_t)c.n, 0);
        parallel_chunks(c.n_short, 2048, nt, [&](int64_t b, int64_t e, int) {
            for (int64_t s = b; s < e; ++s) {
                const SeqPair& sp = pairs[ord[s]];
                bool nq = pack2(seq_qer + sp.idq, sp.len2, c.q.h + qoff[s]);
                bool nr = pack2(seq_ref + sp.idr, sp.len1, c.t.h + toff[s]);
                hasn[s] = (uint8_t)(nq | nr);
                c.meta.h[s] = make_int4((int)qoff[s], (int)toff[s], sp.len2 | (sp.len1 << 16), sp.h0);
            }
        });
        // byte-staged pairs: short pairs containing N (recomputed by the byte-sequence variant of
        // the short kernel) followed by all long pairs (warp-per-pair kernel, N-safe by construction)
        std::vector<int64_t> nlist;
        for (int64_t s = 0; s < c.n_short; ++s) if (hasn[s]) nlist.push_back(s);
        c.n_bytes_short = (int64_t)nlist.size();
        for (int64_t s = c.n_short; s < c.n; ++s) nlist.push_back(s);
        c.n_long = c.n - c.n_short;
        c.n_bytes_pairs = (int64_t)nlist.size();
        if (!nlist.empty()) {
            if (int rc = ensure(eng, c.meta_n, nlist.size())) return rc;
            if (int rc = ensure(eng, c.pos_n, nlist.size())) return rc;
            uint64_t qo = 0, to = 0;
            for (size_t k = 0; k < nlist.size(); ++k) {
                const SeqPair& sp = pairs[ord[nlist[k]]];
                c.meta_n.h[k] = make_int4((int)qo, (int)to, sp.len2 | (sp.len1 << 16), sp.h0);
                c.pos_n.h[k] = (int)nlist[k];
                qo += ((uint64_t)sp.len2 + 7) & ~3ull; to += ((uint64_t)sp.len1 + 7) & ~3ull;   // 4-byte aligned, >= 4 B slack
            }
            if (qo > 0x7fffffffull || to > 0x7fffffffull) { eng->err = "too many N-containing bases in one batch"; return BSW_ERR_PARAM; }
            c.qb_bytes = qo; c.tb_bytes = to;
            if (int rc = ensure(eng, c.qb, c.qb_bytes + 16)) return rc;
            if (int rc = ensure(eng, c.tb, c.tb_bytes + 16)) return rc;
            for (size_t k = 0; k < nlist.size(); ++k) {
                const SeqPair& sp = pairs[ord[nlist[k]]];
                memcpy(c.qb.h + c.meta_n.h[k].x, seq_qer + sp.idq, (size_t)sp.len2);
                memcpy(c.tb.h + c.meta_n.h[k].y, seq_ref + sp.idr, (size_t)sp.len1);
            }
        }
        S.ms_pack += now_ms() - tp0;

        // launch plan: blocks of SHORT_BLOCK consecutive pairs, merged by shared-memory class
        {
            int64_t s = 0;
            while (s < c.n_short) {
                const int64_t blk_end = std::min<int64_t>(c.n_short, s + SHORT_BLOCK);
                const int qmax = pairs[ord[blk_end - 1]].len2;       // ascending in len2
                const int qs = stride_for(qmax);
                if (!c.plan.empty() && c.plan.back().qstride == qs && !c.plan.back().bytes)
                    c.plan.back().count += (int)(blk_end - s);
                else
                    c.plan.push_back(Launch{(int)s, (int)(blk_end - s), qs, false});
                s = blk_end;
            }
            if (c.n_bytes_short > 0) {
                int qmax = 0;
                for (int64_t k = 0; k < c.n_bytes_short; ++k) qmax = std::max(qmax, pairs[ord[nlist[k]]].len2);
                c.plan.push_back(Launch{0, (int)c.n_bytes_short, stride_for(qmax), true});
            }
            if (c.n_long > 0) {
                const int qmax = pairs[ord[c.n - 1]].len2;
                c.long_stride = (qmax + 12) & ~3;
                cudaDeviceProp prop{};
                CUDA_TRY(cudaGetDeviceProperties(&prop, c.dev));
                int64_t blocks = std::min<int64_t>((c.n_long + LONG_WARPS - 1) / LONG_WARPS,
                                                   (int64_t)prop.multiProcessorCount * 8);
                const int64_t cap_words = (int64_t)(256ll << 20) / 4;       // <= 256 MB of eh rows
                blocks = std::max<int64_t>(1, std::min(blocks, cap_words / ((int64_t)c.long_stride * LONG_WARPS)));
                c.long_blocks = (int)blocks;
                if (int rc = ensure(eng, c.scratch, (size_t)blocks * LONG_WARPS * c.long_stride, false)) return rc;
            }
        }

        // async H2D on stream 0
        cudaStream_t st = c.st[0];
        CUDA_TRY(cudaEventRecord(c.ev_h2d0, st));
        CUDA_TRY(cudaMemcpyAsync(c.meta.d, c.meta.h, sizeof(int4) * (size_t)c.n, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(c.q.d, c.q.h, sizeof(uint32_t) * c.q_words, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(c.t.d, c.t.h, sizeof(uint32_t) * c.t_words, cudaMemcpyHostToDevice, st));
        S.h2d_bytes += (int64_t)(sizeof(int4) * (size_t)c.n + 4 * (c.q_words + c.t_words));
        if (!nlist.empty()) {
            CUDA_TRY(cudaMemcpyAsync(c.meta_n.d, c.meta_n.h, sizeof(int4) * nlist.size(), cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(c.pos_n.d, c.pos_n.h, sizeof(int) * nlist.size(), cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(c.qb.d, c.qb.h, c.qb_bytes, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(c.tb.d, c.tb.h, c.tb_bytes, cudaMemcpyHostToDevice, st));
            S.h2d_bytes += (int64_t)(20 * nlist.size() + c.qb_bytes + c.tb_bytes);
        }
        CUDA_TRY(cudaEventRecord(c.ev_h2d1, st));
        S.n_short += (int32_t)c.n_short;
        S.n_long += (int32_t)c.n_long;
    }
    for (DevCtx& c : eng->devs) {
        if (c.n == 0) continue;
        CUDA_TRY(cudaSetDevice(c.dev));
        CUDA_TRY(cudaStreamSynchronize(c.st[0]));
        float ms = 0; CUDA_TRY(cudaEventElapsedTime(&ms, c.ev_h2d0, c.ev_h2d1));
        S.ms_h2d = std::max(S.ms_h2d, (double)ms);
    }
    eng->staged = true;
    S.ms_total = now_ms() - t_begin;
    return BSW_OK;
}

// ------------------------------------------------------------------------------------------
// run: DP kernels only, on every device's staged shard; repeatable
// ------------------------------------------------------------------------------------------
static int launch_device(bsw_engine* eng, DevCtx& c)
{
    CUDA_TRY(cudaSetDevice(c.dev));
    CUDA_TRY(cudaMemsetAsync(c.d_cells, 0, sizeof(unsigned long long), c.st[0]));
    CUDA_TRY(cudaEventRecord(c.ev_k0, c.st[0]));
    for (int s = 1; s < NSTREAMS; ++s) CUDA_TRY(cudaStreamWaitEvent(c.st[s], c.ev_k0, 0));
    int li = 0;
    // longest class first so the tail of the grid is made of the cheapest blocks
    for (int k = (int)c.plan.size() - 1; k >= 0; --k, ++li) {
        const Launch& L = c.plan[k];
        cudaStream_t st = c.st[L.bytes ? 0 : li % NSTREAMS];
        const int grid = (L.count + SHORT_BLOCK - 1) / SHORT_BLOCK;
        const size_t smem = (size_t)L.qstride * SHORT_BLOCK * sizeof(uint32_t);
        if (L.bytes)
            continue;   // byte-variant launches go last, see below
        bsw_short_kernel<SHORT_BLOCK, false><<<grid, SHORT_BLOCK, smem, st>>>(
            c.meta.d, c.q.d, c.t.d, c.res.d, nullptr, L.first, L.count, eng->kp, c.d_cells);
        eng->stats.kernel_launches++;
    }
    for (int s = 1; s < NSTREAMS; ++s) {
        CUDA_TRY(cudaEventRecord(c.ev_join[s], c.st[s]));
        CUDA_TRY(cudaStreamWaitEvent(c.st[0], c.ev_join[s], 0));
    }
    for (const Launch& L : c.plan) {
        if (!L.bytes) continue;
        const int grid = (L.count + SHORT_BLOCK - 1) / SHORT_BLOCK;
        const size_t smem = (size_t)L.qstride * SHORT_BLOCK * sizeof(uint32_t);
        bsw_short_kernel<SHORT_BLOCK, true><<<grid, SHORT_BLOCK, smem, c.st[0]>>>(
            c.meta_n.d, reinterpret_cast<const uint32_t*>(c.qb.d), reinterpret_cast<const uint32_t*>(c.tb.d),
            c.res.d, c.pos_n.d, L.first, L.count, eng->kp, c.d_cells);
        eng->stats.kernel_launches++;
    }
    if (c.n_long > 0) {
        CUDA_TRY(cudaMemsetAsync(c.d_queue, 0, sizeof(unsigned int), c.st[0]));
        bsw_long_kernel<<<c.long_blocks, LONG_WARPS * 32, 0, c.st[0]>>>(
            c.meta_n.d + c.n_bytes_short, c.qb.d, c.tb.d, c.res.d, c.pos_n.d + c.n_bytes_short,
            (int)c.n_long, eng->kp, c.scratch.d, c.long_stride, c.d_queue, c.d_cells);
        eng->stats.kernel_launches++;
    }
    CUDA_TRY(cudaEventRecord(c.ev_k1, c.st[0]));
    CUDA_TRY(cudaMemcpyAsync(c.h_cells, c.d_cells, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c.st[0]));
    CUDA_TRY(cudaGetLastError());
    return BSW_OK;
}

int bsw_run_staged(bsw_engine* eng)
{
    if (!eng) return BSW_ERR_PARAM;
    if (!eng->staged) { eng->err = "bsw_run_staged before bsw_stage"; return BSW_ERR_STATE; }
    bsw_stats& S = eng->stats;
    S.kernel_launches = 0; S.ms_kernel = 0; S.cells_effective = 0;
    for (DevCtx& c : eng->devs)
        if (c.n) { if (int rc = launch_device(eng, c)) return rc; }
    for (DevCtx& c : eng->devs) {
        if (c.n == 0) continue;
        CUDA_TRY(cudaSetDevice(c.dev));
        CUDA_TRY(cudaStreamSynchronize(c.st[0]));
        float ms = 0; CUDA_TRY(cudaEventElapsedTime(&ms, c.ev_k0, c.ev_k1));
        S.ms_kernel = std::max(S.ms_kernel, (double)ms);
        // N-containing pairs are computed twice (2-bit pass result is overwritten); count them once
        S.cells_effective += (int64_t)*c.h_cells;
    }
    eng->ran = true;
    return BSW_OK;
}

// ------------------------------------------------------------------------------------------
// fetch: D2H of the packed results, scatter into the caller's array in input order
// ------------------------------------------------------------------------------------------
int bsw_fetch(bsw_engine* eng, SeqPair* pairs, int64_t n)
{
    if (!eng) return BSW_ERR_PARAM;
    if (!eng->ran) { eng->err = "bsw_fetch before bsw_run_staged"; return BSW_ERR_STATE; }
    if (n != eng->n || (n > 0 && !pairs)) { eng->err = "bsw_fetch: pair count differs from the staged batch"; return BSW_ERR_PARAM; }
    bsw_stats& S = eng->stats;
    S.d2h_bytes = 0; S.ms_d2h = 0;
    for (DevCtx& c : eng->devs) {
        if (c.n == 0) continue;
        CUDA_TRY(cudaSetDevice(c.dev));
        CUDA_TRY(cudaEventRecord(c.ev_d2h0, c.st[0]));
        CUDA_TRY(cudaMemcpyAsync(c.res.h, c.res.d, sizeof(int4) * (size_t)c.n, cudaMemcpyDeviceToHost, c.st[0]));
        CUDA_TRY(cudaEventRecord(c.ev_d2h1, c.st[0]));
        S.d2h_bytes += (int64_t)(sizeof(int4) * (size_t)c.n);
    }
    double ts = 0;
    for (DevCtx& c : eng->devs) {
        if (c.n == 0) continue;
        CUDA_TRY(cudaSetDevice(c.dev));
        CUDA_TRY(cudaStreamSynchronize(c.st[0]));
        float ms = 0; CUDA_TRY(cudaEventElapsedTime(&ms, c.ev_d2h0, c.ev_d2h1));
        S.ms_d2h = std::max(S.ms_d2h, (double)ms);
        const double t0 = now_ms();
        const int64_t* ord = eng->order.data() + c.first;
        const int4* r = c.res.h;
        parallel_chunks(c.n, 1 << 13, eng->nthreads, [&](int64_t b, int64_t e, int) {
            for (int64_t s = b; s < e; ++s) {
                SeqPair& sp = pairs[ord[s]];
                const int4 v = r[s];
                sp.score = (int16_t)(v.x & 0xffff);  sp.qle = (int16_t)(v.x >> 16);
                sp.tle = (int16_t)(v.y & 0xffff);    sp.gtle = (int16_t)(v.y >> 16);
                sp.gscore = (int16_t)(v.z & 0xffff); sp.max_off = (int16_t)(v.z >> 16);
            }
        });
        ts += now_ms() - t0;
    }
    S.ms_scatter = ts;
    return BSW_OK;
}

int bsw_extend(bsw_engine* eng, SeqPair* pairs, const uint8_t* seq_ref, const uint8_t* seq_qer,
               int64_t n, int32_t w)
{
    if (!eng) return BSW_ERR_PARAM;
    const double t0 = now_ms();
    if (int rc = bsw_stage(eng, pairs, seq_ref, seq_qer, n, w)) return rc;
    if (int rc = bsw_run_staged(eng)) return rc;
    if (int rc = bsw_fetch(eng, pairs, n)) return rc;
    eng->stats.ms_total = now_ms() - t0;
    return BSW_OK;
}

double bsw_measure_int_peak(bsw_engine* eng)
{
    if (!eng || eng->devs.empty()) return 0.0;
    DevCtx& c = eng->devs[0];
    if (cudaSetDevice(c.dev) != cudaSuccess) return 0.0;
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, c.dev) != cudaSuccess) return 0.0;
    const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 4096;
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
