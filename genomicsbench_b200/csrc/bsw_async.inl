// bsw_async.inl -- asynchronous submit and call coalescing (bsw_extend_async / bsw_wait); included by
// bsw_engine.cu inside extern "C".
//
// The reference driver feeds getScores16 512 pairs per call from T OpenMP threads (scripts/run-cpu.sh:30,
// main_banded.cpp:279-291): right for a CPU core, far too little for a GPU (a 512-pair call costs 0.3 ms of latency,
// 1.6 M pairs/s, against 50 M pairs/s for the same pairs in one large batch).  bsw_extend_async queues the call and
// returns a ticket; a worker thread owned by the engine takes EVERYTHING that is queued at that moment -- calls of
// different threads, different buffers -- rebases the records onto one pair array and runs them as one batch on a
// private child engine, then hands every call its own six result fields back.  bsw_wait blocks until the ticket's
// results are in the caller's records.  The C++ drop-in class routes small getScores* calls of all its instances
// through one shared engine this way (csrc/bsw_shim.cpp), so the unmodified driver's -t T -b 512 habit turns into
// batches of T x 512 pairs.
namespace {

struct AsyncReq {
    SeqPair* pairs; const uint8_t* ref; const uint8_t* qer; int64_t n; int32_t w;
    int64_t ticket; int rc = BSW_OK; bool done = false; int64_t cells = 0;
};

struct AsyncQueue {
    std::mutex m;
    std::condition_variable cv_work, cv_done;
    std::deque<AsyncReq*> pending;
    std::vector<AsyncReq*> open;          // submitted, not yet waited for
    int64_t next_ticket = 1;
    bool stop = false;
    std::thread worker;
    bsw_engine* child = nullptr;
    std::string child_err;
    int64_t batches = 0, calls = 0;       // statistics: calls coalesced into batches
};

constexpr int64_t ASYNC_MAX_PAIRS = 1 << 18;      // pairs per coalesced batch

void async_run_batch(AsyncQueue& Q, std::vector<AsyncReq*>& batch, std::vector<SeqPair>& comb)
{
    int64_t total = 0;
    const uint8_t* ref0 = batch[0]->ref; const uint8_t* qer0 = batch[0]->qer;
    for (AsyncReq* r : batch) { total += r->n; ref0 = std::min(ref0, r->ref); qer0 = std::min(qer0, r->qer); }
    comb.resize((size_t)total);
    int64_t pos = 0;
    for (AsyncReq* r : batch) {
        const int64_t dr = r->ref - ref0, dq = r->qer - qer0;
        for (int64_t k = 0; k < r->n; ++k) {
            SeqPair sp = r->pairs[k];
            sp.idr += dr; sp.idq += dq;
            comb[(size_t)(pos + k)] = sp;
        }
        pos += r->n;
    }
    int rc = bsw_extend(Q.child, comb.data(), ref0, qer0, total, batch[0]->w);
    if (rc != BSW_OK && batch.size() > 1) {
        // one call's bad input must not fail the others: run them one by one
        for (AsyncReq* r : batch) {
            r->rc = bsw_extend(Q.child, r->pairs, r->ref, r->qer, r->n, r->w);
            r->cells = Q.child->stats.cells_effective;
            if (r->rc != BSW_OK) Q.child_err = Q.child->err;
        }
        return;
    }
    if (rc != BSW_OK) Q.child_err = Q.child->err;
    const int64_t cells = Q.child->stats.cells_effective;
    pos = 0;
    for (AsyncReq* r : batch) {
        r->rc = rc;
        if (rc == BSW_OK)
            for (int64_t k = 0; k < r->n; ++k) {
                const SeqPair& s = comb[(size_t)(pos + k)];
                SeqPair& d = r->pairs[k];
                d.score = s.score; d.tle = s.tle; d.gtle = s.gtle; d.qle = s.qle; d.gscore = s.gscore; d.max_off = s.max_off;
            }
        r->cells = total > 0 ? cells * r->n / total : 0;      // the batch's effective cells, shared out by pair count
        pos += r->n;
    }
}

void async_worker(AsyncQueue* Qp)
{
    AsyncQueue& Q = *Qp;
    std::vector<AsyncReq*> batch;
    std::vector<SeqPair> comb;
    for (;;) {
        batch.clear();
        {
            std::unique_lock<std::mutex> lk(Q.m);
            Q.cv_work.wait(lk, [&] { return Q.stop || !Q.pending.empty(); });
            if (Q.pending.empty()) return;                    // stop, queue drained
            // Linger: callers that block on their previous call (the driver's threads) are all released by the same
            // notify and resubmit within microseconds of each other -- taking the first arrival alone would ping-pong
            // one call per batch for ever.  Keep collecting while calls still arrive, up to ~100 us or 16 k pairs.
            for (int spin = 0; spin < 6 && !Q.stop; ++spin) {
                int64_t have = 0;
                for (const AsyncReq* r : Q.pending) have += r->n;
                if (have >= 16384) break;
                const size_t before = Q.pending.size();
                Q.cv_work.wait_for(lk, std::chrono::microseconds(spin == 0 ? 30 : 15));
                if (Q.pending.size() == before && spin > 0) break;
            }
            const int32_t w = Q.pending.front()->w;
            int64_t total = 0;
            while (!Q.pending.empty() && Q.pending.front()->w == w && (batch.empty() || total + Q.pending.front()->n <= ASYNC_MAX_PAIRS)) {
                batch.push_back(Q.pending.front());
                total += Q.pending.front()->n;
                Q.pending.pop_front();
            }
        }
        async_run_batch(Q, batch, comb);
        {
            std::lock_guard<std::mutex> g(Q.m);
            for (AsyncReq* r : batch) r->done = true;
            Q.batches += 1; Q.calls += (int64_t)batch.size();
        }
        Q.cv_done.notify_all();
    }
}

} // namespace

static void bsw_async_release(bsw_engine* eng)
{
    AsyncQueue* Q = static_cast<AsyncQueue*>(eng->aq);
    if (!Q) return;
    {
        std::lock_guard<std::mutex> g(Q->m);
        Q->stop = true;
    }
    Q->cv_work.notify_all();
    if (Q->worker.joinable()) Q->worker.join();
    for (AsyncReq* r : Q->open) delete r;
    bsw_destroy(Q->child);
    delete Q;
    eng->aq = nullptr;
}

int bsw_extend_async(bsw_engine* eng, SeqPair* pairs, const uint8_t* seq_ref, const uint8_t* seq_qer, int64_t n, int32_t w,
                     int64_t* ticket)
{
    if (!eng || !ticket) return BSW_ERR_PARAM;
    if (n < 0 || w < 0 || n > ASYNC_MAX_PAIRS * 64 || (n > 0 && (!pairs || !seq_ref || !seq_qer))) return BSW_ERR_PARAM;
    std::call_once(eng->aq_once, [&] {
        AsyncQueue* Q = new AsyncQueue();
        bsw_params p = eng->p;                               // same scoring, same devices
        int err = 0;
        Q->child = bsw_create(&p, &err);
        if (Q->child) Q->worker = std::thread(async_worker, Q);
        eng->aq = Q;
    });
    AsyncQueue* Q = static_cast<AsyncQueue*>(eng->aq);
    if (!Q || !Q->child) return BSW_ERR_CUDA;
    AsyncReq* r = new AsyncReq{pairs, seq_ref, seq_qer, n, w, 0};
    {
        std::lock_guard<std::mutex> g(Q->m);
        r->ticket = Q->next_ticket++;
        *ticket = r->ticket;
        Q->open.push_back(r);
        if (n == 0) r->done = true;
        else Q->pending.push_back(r);
    }
    Q->cv_work.notify_one();
    return BSW_OK;
}

int bsw_wait(bsw_engine* eng, int64_t ticket, int64_t* cells_effective)
{
    if (!eng || !eng->aq) return BSW_ERR_STATE;
    AsyncQueue* Q = static_cast<AsyncQueue*>(eng->aq);
    std::unique_lock<std::mutex> lk(Q->m);
    AsyncReq* r = nullptr;
    size_t at = 0;
    for (size_t k = 0; k < Q->open.size(); ++k) if (Q->open[k]->ticket == ticket) { r = Q->open[k]; at = k; break; }
    if (!r) return BSW_ERR_STATE;                             // unknown ticket, or waited for twice
    (void)at;
    Q->cv_done.wait(lk, [&] { return r->done; });
    const int rc = r->rc;
    if (cells_effective) *cells_effective = r->cells;
    if (rc != BSW_OK) eng->err = Q->child_err;
    // (other waiters erased their entries while this one slept: look the entry up again)
    Q->open.erase(std::find(Q->open.begin(), Q->open.end(), r));
    delete r;
    return rc;
}

int bsw_async_stats(const bsw_engine* eng, int64_t* calls, int64_t* batches)
{
    if (!eng) return BSW_ERR_PARAM;
    AsyncQueue* Q = static_cast<AsyncQueue*>(eng->aq);
    if (calls) *calls = 0;
    if (batches) *batches = 0;
    if (!Q) return BSW_OK;
    std::lock_guard<std::mutex> g(Q->m);
    if (calls) *calls = Q->calls;
    if (batches) *batches = Q->batches;
    return BSW_OK;
}
