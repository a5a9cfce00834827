// queue.cu -- coalescing queue for SMALL MSMs issued concurrently by many host threads (BASELINE configs[3]: block
// validation).
//
// Where it sits in the reference: a validator checks the transactions of a block in parallel -- up to 512 tokio blocking
// threads (/root/reference/cli/src/commands/start.rs:623-640) run `check_transaction_basic` / `check_next_block`
// (/root/reference/node/bft/ledger-service/src/ledger.rs:341-347, node/consensus/src/lib.rs:499), and each of them ends
// in Varuna::verify_batch -> KZG10::batch_check, whose linear combinations of commitments are VariableBase::msm calls
// over a few tens of points [UPSTREAM algorithms/src/polycommit/kzg10/mod.rs].  One such call is ~25 kernel launches of
// latency-bound work; hundreds of them on hundreds of streams serialise on the launch path.  Here every caller SUBMITS
// its MSM and blocks on its own ticket; one dispatcher thread takes whatever is pending, packs it into ONE segmented
// launch set (msm_run_batch_device: "virtual windows", see msm.cu) on its own stream, and completes the tickets.  While a
// batch is on the GPU the next one accumulates, so the batch size adapts to the arrival rate without a timer.  The
// queue mutex is held only to append / swap the pending list -- never across a copy, a launch or a synchronise.
//
// The Rust call sites stay unchanged: with option `msm_queue_threshold` = T (B200_MSM_QUEUE_THRESHOLD), every
// b200_msm_g1_bls12_377 call of <= T points goes through submit + wait internally.
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <memory>
#include <system_error>
#include <thread>
#include <unordered_map>

#include "common.cuh"

#define QUEUE_MAX_MSMS_PER_BATCH 4096u
#define QUEUE_MAX_POINTS_PER_BATCH ((size_t)1 << 20)

struct QueueJob {
    const void* points;
    const void* scalars;
    size_t n, stride;
    std::mutex mu;
    std::condition_variable cv;
    bool done = false;
    b200_error_t err{0, "ok"};
    uint8_t out[144];
};

struct QueueState {
    std::mutex mu;
    std::condition_variable cv;                     // dispatcher sleeps here
    std::deque<std::shared_ptr<QueueJob>> pending;
    std::unordered_map<uint64_t, std::shared_ptr<QueueJob>> tickets;
    uint64_t next_ticket = 1;
    std::thread worker;
    bool running = false, stop = false;
    // dispatcher-owned staging (grown on demand, released at shutdown)
    void* h_stage = nullptr;
    size_t h_cap = 0;
    void* d_stage = nullptr;
    size_t d_cap = 0;
    cudaStream_t stream = nullptr;
};
// Never destroyed: a process that exits without b200_shutdown still has the dispatcher parked on its condition
// variable, and destroying a joinable std::thread (or a mutex a thread sleeps on) during static destruction aborts.
static QueueState& g_q = *new QueueState();

static void job_complete(QueueJob& j, b200_error_t err) {
    std::lock_guard<std::mutex> lock(j.mu);
    j.err = err;
    j.done = true;
    j.cv.notify_all();
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// one batch: pack into pinned staging, one H2D, one segmented MSM, one D2H
static b200_error_t queue_run_batch(std::vector<std::shared_ptr<QueueJob>>& batch) {
    B200_TRY(b200_require_device());
    if (!g_q.stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&g_q.stream, cudaStreamNonBlocking));
        b200_scratch_register_stream(g_q.stream);
    }
    const size_t nmsm = batch.size(), stride = batch[0]->stride;
    size_t npts = 0;
    for (auto& j : batch) npts += j->n;
    // staging layout (host and device identical): points | scalars | offsets | results
    const size_t o_pts = 0, o_sc = align_up(npts * stride, 256), o_off = o_sc + align_up(npts * 32, 256),
                 o_out = o_off + align_up((nmsm + 1) * 8, 256), total = o_out + nmsm * 144;
    if (total > g_q.h_cap) {
        if (g_q.h_stage) cudaFreeHost(g_q.h_stage);
        if (g_q.d_stage) cudaFree(g_q.d_stage);
        g_q.h_stage = g_q.d_stage = nullptr;
        g_q.h_cap = g_q.d_cap = 0;
        const size_t cap = align_up(total * 2, (size_t)1 << 20);
        CUDA_TRY(cudaHostAlloc(&g_q.h_stage, cap, cudaHostAllocDefault));
        CUDA_TRY(cudaMalloc(&g_q.d_stage, cap));
        g_q.h_cap = g_q.d_cap = cap;
    }
    uint8_t* h = reinterpret_cast<uint8_t*>(g_q.h_stage);
    uint8_t* d = reinterpret_cast<uint8_t*>(g_q.d_stage);
    uint64_t* off = reinterpret_cast<uint64_t*>(h + o_off);
    size_t at = 0;
    for (size_t m = 0; m < nmsm; m++) {
        const QueueJob& j = *batch[m];
        off[m] = at;
        if (j.n) {
            memcpy(h + o_pts + at * stride, j.points, j.n * stride);
            memcpy(h + o_sc + at * 32, j.scalars, j.n * 32);
        }
        at += j.n;
    }
    off[nmsm] = at;
    cudaStream_t s = g_q.stream;
    CUDA_TRY(cudaMemcpyAsync(d, h, o_out, cudaMemcpyHostToDevice, s));
    B200_TRY(msm_run_batch_device(d + o_out, d + o_pts, npts, d + o_sc, stride, nullptr,
                                  reinterpret_cast<const unsigned long long*>(d + o_off), (uint32_t)nmsm, s));
    CUDA_TRY(cudaMemcpyAsync(h + o_out, d + o_out, nmsm * 144, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    for (size_t m = 0; m < nmsm; m++) memcpy(batch[m]->out, h + o_out + m * 144, 144);
    return b200_ok();
}

static void queue_worker() {
    for (;;) {
        std::vector<std::shared_ptr<QueueJob>> batch;
        {
            std::unique_lock<std::mutex> lock(g_q.mu);
            g_q.cv.wait(lock, [] { return g_q.stop || !g_q.pending.empty(); });
            if (g_q.pending.empty()) {
                if (g_q.stop) return;
                continue;
            }
            // Linger: a launch set costs ~2 ms however few MSMs it carries, and the callers of a burst (the transactions
            // of one block) arrive within a fraction of that -- so wait while submissions keep coming (no arrival for
            // `linger` microseconds ends it, four times that is the cap).  A lone caller pays `linger` once.
            const int linger = b200_config().msm_queue_linger_us;
            if (linger > 0 && !g_q.stop) {
                const auto t_end = std::chrono::steady_clock::now() + std::chrono::microseconds(4 * linger);
                for (;;) {
                    const size_t seen = g_q.pending.size();
                    if (seen >= QUEUE_MAX_MSMS_PER_BATCH) break;
                    g_q.cv.wait_for(lock, std::chrono::microseconds(linger), [&] { return g_q.stop || g_q.pending.size() != seen; });
                    if (g_q.stop || g_q.pending.size() == seen || std::chrono::steady_clock::now() >= t_end) break;
                }
            }
            // everything pending with the stride of the first job, up to the batch limits
            const size_t stride = g_q.pending.front()->stride;
            size_t npts = 0;
            for (auto it = g_q.pending.begin(); it != g_q.pending.end() && batch.size() < QUEUE_MAX_MSMS_PER_BATCH;) {
                if ((*it)->stride != stride) { ++it; continue; }
                if (!batch.empty() && npts + (*it)->n > QUEUE_MAX_POINTS_PER_BATCH) break;
                npts += (*it)->n;
                batch.push_back(*it);
                it = g_q.pending.erase(it);
            }
        }
        g_counters.queue_batches.fetch_add(1, std::memory_order_relaxed);
        const b200_error_t r = queue_run_batch(batch);
        if (r.code != 0) (void)cudaGetLastError();
        for (auto& j : batch) job_complete(*j, r);
    }
}

static b200_error_t queue_start_locked() {
    if (g_q.running) return b200_ok();
    try {
        g_q.stop = false;
        g_q.worker = std::thread(queue_worker);
    } catch (const std::system_error&) {
        return b200_err(B200_ERR_NOT_INITIALIZED, "msm_submit: could not start the dispatcher thread");
    }
    g_q.running = true;
    return b200_ok();
}

extern "C" b200_error_t b200_msm_submit(const void* points, size_t npoints, const void* scalars, size_t affine_stride,
                                        uint64_t* out_ticket) {
    B200_TRY(b200_require_device());
    if (!out_ticket) return b200_err(B200_ERR_INVALID_ARG, "msm_submit: null ticket pointer");
    if (npoints && (!points || !scalars)) return b200_err(B200_ERR_INVALID_ARG, "msm_submit: null input pointer");
    if (affine_stride < 104 || (affine_stride & 7)) return b200_err(B200_ERR_INVALID_ARG, "msm_submit: affine stride must be >= 104 and 8-byte aligned");
    if (npoints > QUEUE_MAX_POINTS_PER_BATCH) return b200_err(B200_ERR_TOO_LARGE, "msm_submit: the queue is for small MSMs (<= 2^20 points); call b200_msm_g1_bls12_377");
    auto job = std::make_shared<QueueJob>();
    job->points = points;
    job->scalars = scalars;
    job->n = npoints;
    job->stride = affine_stride;
    {
        std::lock_guard<std::mutex> lock(g_q.mu);
        B200_TRY(queue_start_locked());
        const uint64_t t = g_q.next_ticket++;
        g_q.tickets.emplace(t, job);
        g_q.pending.push_back(job);
        *out_ticket = t;
    }
    g_counters.queue_submits.fetch_add(1, std::memory_order_relaxed);
    g_q.cv.notify_one();
    return b200_ok();
}

extern "C" b200_error_t b200_msm_wait(uint64_t ticket, void* out_jacobian_144B) {
    std::shared_ptr<QueueJob> job;
    {
        std::lock_guard<std::mutex> lock(g_q.mu);
        auto it = g_q.tickets.find(ticket);
        if (it == g_q.tickets.end()) return b200_err(B200_ERR_BAD_HANDLE, "msm_wait: unknown ticket");
        job = it->second;
        g_q.tickets.erase(it);                        // a ticket is waited on exactly once
    }
    std::unique_lock<std::mutex> lock(job->mu);
    job->cv.wait(lock, [&] { return job->done; });
    if (job->err.code == 0 && out_jacobian_144B) memcpy(out_jacobian_144B, job->out, 144);
    return job->err;
}

void b200_queue_shutdown() {
    std::thread worker;
    {
        std::lock_guard<std::mutex> lock(g_q.mu);
        if (!g_q.running) return;
        g_q.stop = true;                              // the dispatcher drains what is pending, then exits
        g_q.running = false;
        worker.swap(g_q.worker);
    }
    g_q.cv.notify_all();
    if (worker.joinable()) worker.join();
    if (g_q.h_stage) cudaFreeHost(g_q.h_stage);
    if (g_q.d_stage) cudaFree(g_q.d_stage);
    if (g_q.stream) {
        b200_scratch_forget_stream(g_q.stream);             // its cached scratch goes back while the stream still exists
        cudaStreamDestroy(g_q.stream);
    }
    g_q.h_stage = g_q.d_stage = nullptr;
    g_q.h_cap = g_q.d_cap = 0;
    g_q.stream = nullptr;
}
