// queue_bench -- BASELINE configs[3] shape from NATIVE threads: T host threads (the validator's tokio blocking threads /
// rayon workers, /root/reference/cli/src/commands/start.rs:623-640) each verify M transactions, i.e. issue M small
// VariableBase::msm calls of P points one after another, all threads at once.  Three ways through the C ABI:
//   direct : every thread calls b200_msm_g1_bls12_377 itself (own stream, ~25 launches per call)
//   queued : b200_msm_submit + b200_msm_wait (queue.cu: one dispatcher coalesces whatever is pending into one launch set)
//   batch  : ONE b200_msm_batch_g1_bls12_377 call with all T * M MSMs (the lower bound a shim above the loop would get)
// Results of the three are compared through b200_g1_compress.  Prints one JSON line.  Python threads cannot show this
// (the GIL serialises the submits): bench.py runs this binary for its batch_verify_msm entry.
#include <cuda_runtime.h>
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../include/snarkos_b200.h"

static uint64_t splitmix(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
#define CHECK(e) do { b200_error_t _r = (e); if (_r.code != 0) { fprintf(stderr, "error %d: %s (%s)\n", _r.code, _r.msg, #e); exit(1); } } while (0)

static void on_segv(int sig) {                      // a crash must say where: frames resolve with addr2line against the .so
    void* frames[64];
    const int n = backtrace(frames, 64);
    const char msg[] = "queue_bench: fatal signal, backtrace:\n";
    (void)!write(2, msg, sizeof(msg) - 1);
    backtrace_symbols_fd(frames, n, 2);
    _exit(128 + sig);
}

int main(int argc, char** argv) {
    signal(SIGSEGV, on_segv);
    signal(SIGABRT, on_segv);
    const int T = argc > 1 ? atoi(argv[1]) : 64, M = argc > 2 ? atoi(argv[2]) : 4, P = argc > 3 ? atoi(argv[3]) : 40;
    const int reps = argc > 4 ? atoi(argv[4]) : 7;
    const size_t nmsm = (size_t)T * M, npts = nmsm * P;
    CHECK(b200_init(0));
    void* d_pts = nullptr;
    cudaMalloc(&d_pts, npts * 104);
    CHECK(b200_g1_synthetic_bases_device(d_pts, npts, 104, 424242, nullptr));
    std::vector<uint8_t> pts(npts * 104);
    cudaMemcpy(pts.data(), d_pts, pts.size(), cudaMemcpyDeviceToHost);
    cudaFree(d_pts);
    std::vector<uint64_t> sc(npts * 4);
    uint64_t seed = 7;
    for (size_t i = 0; i < npts; i++) {
        for (int k = 0; k < 4; k++) sc[4 * i + k] = splitmix(seed);
        sc[4 * i + 3] &= (1ull << 60) - 1;
    }
    std::vector<uint64_t> off(nmsm + 1);
    for (size_t m = 0; m <= nmsm; m++) off[m] = m * P;
    std::vector<uint8_t> out_direct(nmsm * 144), out_queued(nmsm * 144), out_batch(nmsm * 144);

    auto run_threads = [&](bool queued, std::vector<uint8_t>& out) {
        std::atomic<int> ready{0};
        std::atomic<bool> go{false};
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++)
            th.emplace_back([&, t] {
                ready.fetch_add(1);
                while (!go.load(std::memory_order_acquire)) {}
                for (int m = 0; m < M; m++) {
                    const size_t i = (size_t)t * M + m;
                    if (queued) {
                        uint64_t ticket = 0;
                        CHECK(b200_msm_submit(pts.data() + i * P * 104, P, sc.data() + i * P * 4, 104, &ticket));
                        CHECK(b200_msm_wait(ticket, out.data() + i * 144));
                    } else {
                        CHECK(b200_msm_g1_bls12_377(out.data() + i * 144, pts.data() + i * P * 104, P, sc.data() + i * P * 4, 104));
                    }
                }
            });
        while (ready.load() < T) {}
        const auto t0 = std::chrono::steady_clock::now();
        go.store(true, std::memory_order_release);
        for (auto& x : th) x.join();
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    };
    auto median = [](std::vector<double> v) { std::sort(v.begin(), v.end()); return v[v.size() / 2]; };

    std::vector<double> t_direct, t_queued, t_batch, t_batch2;
    uint64_t b0 = 0, b1 = 0, s0 = 0, s1 = 0;
    for (int r = 0; r < reps + 1; r++) {                         // first repetition is the warm-up
        const double a = run_threads(false, out_direct);
        if (r == 1) { CHECK(b200_get_counter("queue_batches", &b0)); CHECK(b200_get_counter("queue_submits", &s0)); }
        const double b = run_threads(true, out_queued);
        const auto t0 = std::chrono::steady_clock::now();
        CHECK(b200_msm_batch_g1_bls12_377(out_batch.data(), pts.data(), sc.data(), off.data(), nmsm, 104));
        const double c = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        const auto t1 = std::chrono::steady_clock::now();
        CHECK(b200_msm_batch_g1_bls12_377(out_batch.data(), pts.data(), sc.data(), off.data(), nmsm, 104));
        const double c2 = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
        if (r) { t_direct.push_back(a); t_queued.push_back(b); t_batch.push_back(c); t_batch2.push_back(c2); }
    }
    // cost of one stream-ordered allocation in this process state (many caller streams hold freed blocks)
    double alloc_us = 0;
    {
        cudaStream_t st;
        cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        const auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < 200; i++) {
            void* p = nullptr;
            cudaMallocAsync(&p, 1 << 20, st);
            cudaFreeAsync(p, st);
        }
        cudaStreamSynchronize(st);
        alloc_us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / 200;
        cudaStreamDestroy(st);
    }
    char prof[4096] = {0};
    b200_profile_begin();
    CHECK(b200_msm_batch_g1_bls12_377(out_batch.data(), pts.data(), sc.data(), off.data(), nmsm, 104));
    CHECK(b200_profile_end(prof, sizeof(prof)));
    fprintf(stderr, "T=%d batch call device stages: %s\n", T, prof);
    CHECK(b200_get_counter("queue_batches", &b1));
    CHECK(b200_get_counter("queue_submits", &s1));
    std::vector<uint8_t> c_direct(nmsm * 48), c_queued(nmsm * 48), c_batch(nmsm * 48);
    CHECK(b200_g1_compress(c_direct.data(), out_direct.data(), nmsm));
    CHECK(b200_g1_compress(c_queued.data(), out_queued.data(), nmsm));
    CHECK(b200_g1_compress(c_batch.data(), out_batch.data(), nmsm));
    const bool same = c_direct == c_queued && c_direct == c_batch;
    printf("{\"threads\": %d, \"msms_per_thread\": %d, \"points\": %d, \"direct_ms\": %.3f, \"queued_ms\": %.3f, \"one_batch_call_ms\": %.3f, \"one_batch_call_again_ms\": %.3f, "
           "\"queue_batches_per_round\": %.2f, \"submits_per_round\": %.1f, \"alloc_free_pair_us\": %.1f, \"results_identical\": %s}\n",
           T, M, P, median(t_direct), median(t_queued), median(t_batch), median(t_batch2), (double)(b1 - b0) / reps, (double)(s1 - s0) / reps, alloc_us,
           same ? "true" : "false");
    b200_shutdown();
    return same ? 0 : 2;
}
