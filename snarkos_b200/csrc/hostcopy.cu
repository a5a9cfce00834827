// hostcopy.cu -- host <-> device copies of the extern "C" boundary that stay fast for PAGEABLE caller memory.
//
// The reference's callers hand over ordinary Rust `Vec`s (`VariableBase::msm(&[G1Affine], &[BigInteger256])`,
// `fft_in_place(&mut Vec<Fr>)`; SURVEY.md 8b "Ownership": the library copies H2D / D2H internally, pinned staging or
// cudaHostRegister on demand).  cudaMemcpyAsync out of pageable memory is staged by the driver through one bounce
// buffer on the calling thread: measured 247 ms instead of 104 ms for a 2^24-point MSM from host buffers.  Pinning the
// caller's pages in place (cudaHostRegister) costs more than the copy for buffers used once.  So: a pinned or
// registered source goes straight through cudaMemcpyAsync; a pageable one is cut into 4 MiB slices that a few worker
// threads memcpy into their own pinned slots (two per worker, reused under an event) and upload from there, all
// ordered inside the caller's stream: the uploads start after everything already queued on it and the stream resumes
// when the last slice has landed.
#include <cstring>
#include <system_error>
#include <thread>

#include "common.cuh"

#define HC_SLICE ((size_t)4 << 20)
#define HC_WORKERS 8
#define HC_MIN_STAGED ((size_t)8 << 20)      // below this the driver's own path is as good

struct HcWorker {
    cudaStream_t st = nullptr;
    void* slot[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaEvent_t done = nullptr;
};
// One staging set = 8 workers x 2 pinned 4 MiB slots (64 MiB), 8 streams, 25 events.  Sets live in a small PROCESS-WIDE
// pool: a call checks one out under the mutex and returns it when its copies are queued, so short-lived caller threads
// (tokio's blocking pool creates and retires threads) neither leak pinned memory nor pay the pinning again, and
// b200_shutdown releases everything.  Concurrent callers never share a set; at most HC_MAX_SETS exist, a caller that
// finds none free takes the driver's own pageable path for that copy (slower, never wrong).
struct HcPool {
    HcWorker w[HC_WORKERS];
    cudaEvent_t start = nullptr;
    bool ready = false;
    bool busy = false;
};
#define HC_MAX_SETS 4
static std::mutex g_hc_mu;
static HcPool g_hc_sets[HC_MAX_SETS];

static void hc_destroy(HcPool& p) {
    if (p.start) cudaEventDestroy(p.start);
    for (int k = 0; k < HC_WORKERS; k++) {
        HcWorker& w = p.w[k];
        if (w.st) cudaStreamDestroy(w.st);
        if (w.done) cudaEventDestroy(w.done);
        for (int s = 0; s < 2; s++) {
            if (w.slot[s]) cudaFreeHost(w.slot[s]);
            if (w.ev[s]) cudaEventDestroy(w.ev[s]);
        }
    }
    p = HcPool();
}

static bool hc_prepare(HcPool& p) {
    if (p.ready) return true;
    bool ok = cudaEventCreateWithFlags(&p.start, cudaEventDisableTiming) == cudaSuccess;
    for (int k = 0; k < HC_WORKERS && ok; k++) {
        HcWorker& w = p.w[k];
        ok = cudaStreamCreateWithFlags(&w.st, cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming) == cudaSuccess;
        for (int s = 0; s < 2 && ok; s++)
            ok = cudaHostAlloc(&w.slot[s], HC_SLICE, cudaHostAllocDefault) == cudaSuccess &&
                 cudaEventCreateWithFlags(&w.ev[s], cudaEventDisableTiming) == cudaSuccess;
    }
    if (!ok) {
        (void)cudaGetLastError();
        hc_destroy(p);                         // fall back to the driver's pageable path, never to a CPU computation
        return false;
    }
    p.ready = true;
    return true;
}

// a free staging set, or nullptr (all in use / pinning failed)
static HcPool* hc_checkout() {
    std::lock_guard<std::mutex> lock(g_hc_mu);
    for (int i = 0; i < HC_MAX_SETS; i++) {
        HcPool& p = g_hc_sets[i];
        if (p.busy) continue;
        if (!hc_prepare(p)) return nullptr;
        p.busy = true;
        return &p;
    }
    return nullptr;
}
static void hc_return(HcPool* p) {
    std::lock_guard<std::mutex> lock(g_hc_mu);
    p->busy = false;
}
void hostcopy_release() {
    std::lock_guard<std::mutex> lock(g_hc_mu);
    for (int i = 0; i < HC_MAX_SETS; i++)
        if (!g_hc_sets[i].busy) hc_destroy(g_hc_sets[i]);
}

static bool hc_is_pageable(const void* host) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, host) != cudaSuccess) {
        (void)cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

// direction 0: host -> device, 1: device -> host
static b200_error_t hc_staged(HcPool& p, void* dst, const void* src, size_t bytes, int direction, cudaStream_t stream) {
    int device = 0;
    CUDA_TRY(cudaGetDevice(&device));
    CUDA_TRY(cudaEventRecord(p.start, stream));
    const size_t nslices = (bytes + HC_SLICE - 1) / HC_SLICE;
    const int nworkers = nslices < HC_WORKERS ? (int)nslices : HC_WORKERS;
    cudaError_t errs[HC_WORKERS];
    std::thread th[HC_WORKERS];
    int started = 0;
    bool spawn_failed = false;
    for (int k = 0; k < nworkers; k++) {
        errs[k] = cudaSuccess;
        auto body = [&, k]() {
            HcWorker& w = p.w[k];
            cudaError_t e = cudaSetDevice(device);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(w.st, p.start, 0);
            uint32_t use = 0;
            size_t pending_off[2] = {0, 0}, pending_len[2] = {0, 0};
            for (size_t j = (size_t)k; j < nslices && e == cudaSuccess; j += (size_t)nworkers, use++) {
                const int s = (int)(use & 1u);
                const size_t off = j * HC_SLICE, len = bytes - off < HC_SLICE ? bytes - off : HC_SLICE;
                e = cudaEventSynchronize(w.ev[s]);        // the slot's previous transfer (this call or an earlier one) is over
                if (e != cudaSuccess) break;
                if (direction == 1 && use >= 2) memcpy((uint8_t*)dst + pending_off[s], w.slot[s], pending_len[s]);
                if (direction == 0) {
                    memcpy(w.slot[s], (const uint8_t*)src + off, len);
                    e = cudaMemcpyAsync((uint8_t*)dst + off, w.slot[s], len, cudaMemcpyHostToDevice, w.st);
                } else {
                    e = cudaMemcpyAsync(w.slot[s], (const uint8_t*)src + off, len, cudaMemcpyDeviceToHost, w.st);
                    pending_off[s] = off;
                    pending_len[s] = len;
                }
                if (e == cudaSuccess) e = cudaEventRecord(w.ev[s], w.st);
            }
            if (direction == 1 && e == cudaSuccess) {                               // drain the (up to two) slots in flight
                const uint32_t first = use >= 2 ? use - 2 : 0;
                for (uint32_t u = first; u < use && e == cudaSuccess; u++) {
                    const int s = (int)(u & 1u);
                    e = cudaEventSynchronize(w.ev[s]);
                    if (e == cudaSuccess) memcpy((uint8_t*)dst + pending_off[s], w.slot[s], pending_len[s]);
                }
            }
            if (e == cudaSuccess) e = cudaEventRecord(w.done, w.st);
            errs[k] = e;
        };
        // nothing may throw across the extern "C" boundary: if the OS refuses a thread, this worker's share runs here
        if (!spawn_failed) {
            try {
                th[k] = std::thread(body);
                started = k + 1;
                continue;
            } catch (const std::system_error&) {
                spawn_failed = true;
            }
        }
        body();
    }
    for (int k = 0; k < started; k++) th[k].join();
    for (int k = 0; k < nworkers; k++)
        if (errs[k] != cudaSuccess) return b200_cuda_err(errs[k]);
    for (int k = 0; k < nworkers; k++) CUDA_TRY(cudaStreamWaitEvent(stream, p.w[k].done, 0));
    return b200_ok();
}

// Upload ordered in `stream`.  On return the source may be reused (pageable path) or must stay valid until the
// stream reaches the copy (pinned path, as with cudaMemcpyAsync).
b200_error_t b200_h2d(void* d_dst, const void* h_src, size_t bytes, cudaStream_t stream) {
    if (bytes == 0) return b200_ok();
    NvtxRange range("b200_h2d");
    if (bytes >= HC_MIN_STAGED && b200_config().staged_copies && hc_is_pageable(h_src)) {
        if (HcPool* p = hc_checkout()) {
            const b200_error_t r = hc_staged(*p, d_dst, h_src, bytes, 0, stream);
            hc_return(p);
            return r;
        }
    }
    CUDA_TRY(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, stream));
    return b200_ok();
}

// Download ordered after everything queued on `stream`.  Pageable destination: the data has landed on return; pinned
// destination: as with cudaMemcpyAsync (synchronise the stream before reading).
b200_error_t b200_d2h(void* h_dst, const void* d_src, size_t bytes, cudaStream_t stream) {
    if (bytes == 0) return b200_ok();
    NvtxRange range("b200_d2h");
    if (bytes >= HC_MIN_STAGED && b200_config().staged_copies && hc_is_pageable(h_dst)) {
        if (HcPool* p = hc_checkout()) {
            const b200_error_t r = hc_staged(*p, h_dst, d_src, bytes, 1, stream);
            hc_return(p);
            return r;
        }
    }
    CUDA_TRY(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, stream));
    return b200_ok();
}
