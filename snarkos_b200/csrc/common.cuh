// common.cuh -- error plumbing, per-process context and launch accounting shared by the .cu files.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <map>
#include <mutex>
#include <vector>

#include "../../include/snarkos_b200.h"

extern std::atomic<uint64_t> g_kernel_launches;
#define B200_LAUNCH_COUNT() (g_kernel_launches.fetch_add(1, std::memory_order_relaxed))

static inline b200_error_t b200_ok() { return b200_error_t{0, "ok"}; }
static inline b200_error_t b200_err(int32_t code, const char* msg) { return b200_error_t{code, msg}; }
static inline b200_error_t b200_cuda_err(cudaError_t e) { return b200_error_t{(int32_t)e, cudaGetErrorString(e)}; }

#define CUDA_TRY(expr)                                    \
    do {                                                  \
        cudaError_t _e = (expr);                          \
        if (_e != cudaSuccess) return b200_cuda_err(_e);  \
    } while (0)
#define B200_TRY(expr)                     \
    do {                                   \
        b200_error_t _r = (expr);          \
        if (_r.code != 0) return _r;       \
    } while (0)
#define KERNEL_CHECK()                                    \
    do {                                                  \
        B200_LAUNCH_COUNT();                              \
        cudaError_t _e = cudaGetLastError();              \
        if (_e != cudaSuccess) return b200_cuda_err(_e);  \
    } while (0)

// Tuning knobs.  The B200_* environment variables are read ONCE, when the library is first used (b200_init or the first
// compute call); afterwards b200_set_option changes them at run time (sweep tools, tests).  Nothing on the call path calls
// getenv.
struct B200Config {
    int msm_c = 0;                   // msm_window_bits          B200_MSM_C             0 = automatic
    bool msm_glv = true;             // msm_glv                  B200_MSM_NO_GLV
    int msm_affine_rounds = -1;      // msm_affine_rounds        B200_MSM_AFFINE_ROUNDS -1 = automatic
    int msm_slices = 1;              // msm_slices               B200_MSM_SLICES
    int msm_chunk = 0;               // msm_chunk                B200_MSM_CHUNK         0 = automatic
    int msm_seg_len = 0;             // msm_seg_len              B200_MSM_SEG_LEN       buckets per running-sum segment of the reduction, 0 = automatic
    int msm_reduce_quad_max = 8192;  // msm_reduce_quad_max      B200_MSM_REDUCE_QUAD_MAX  up to this many segments: four lanes per segment (ec_coop.cuh)
    bool msm_host_pipeline = true;   // msm_host_pipeline        B200_MSM_NO_HOST_PIPELINE
    int msm_host_first_log = 20;     // msm_host_first_log       B200_MSM_HOST_FIRST_LOG
    int msm_host_chunk_log = 23;     // msm_host_chunk_log       B200_MSM_HOST_CHUNK_LOG  host calls of >= 2^this points are streamed
                                     //                          in ranges of at most 2^this points (tests lower it)
    bool msm_stream_two = true;      // msm_stream_two           B200_MSM_NO_STREAM_TWO    streamed ranges alternate between two compute streams
    bool msm_auto_table = true;      // msm_auto_table           B200_MSM_NO_AUTO_TABLE
    unsigned long long msm_list_budget = 0;   // msm_list_budget_bytes: cap on the pair-round scratch of ONE call
                                     //                          (0 = whatever cudaMallocAsync grants); above it the call
                                     //                          runs the XYZZ-only path, exactly as on an allocation failure
    int msm_queue_threshold = 0;     // msm_queue_threshold      B200_MSM_QUEUE_THRESHOLD  host MSMs of <= this many points go
                                     //                          through the coalescing queue (queue.cu); 0 = off
    int msm_queue_linger_us = 100;   // msm_queue_linger_us      B200_MSM_QUEUE_LINGER_US  the dispatcher waits this long for further submissions
                                     //                          before it launches a batch (at most 4 x; 0 = launch at once)
    char ntt_plan[32] = {0};         // ntt_plan "a,b,c"         B200_NTT_PLAN
    int ntt_tile_log = 0;            // ntt_tile_log             B200_NTT_TILE_LOG
    bool ntt_radix4 = true;          // ntt_radix4               B200_NTT_RADIX2
    bool ntt_boundary_tables = true; // ntt_boundary_tables      B200_NTT_NO_BOUNDARY_TABLES
    bool ntt_host_pipeline = true;   // ntt_host_pipeline        B200_NTT_NO_HOST_PIPELINE
    int ntt_variant = 0;             // ntt_variant              B200_NTT_VARIANT       0 = default (128-thread CTAs on 1024-element tiles),
                                     //                          1 = bulk-copy TMA, 3 = 256-thread CTAs only, 4 = warp-column kernel for 2^8 passes
    bool staged_copies = true;       // staged_copies            B200_NO_STAGED_COPIES
    int l2_fetch_granularity = 0;    // (init only)              B200_L2_FETCH_GRANULARITY
};
B200Config& b200_config();           // loaded from the environment on first use

// fallbacks and queue activity the caller can observe (b200_get_counter)
struct B200Counters {
    std::atomic<uint64_t> msm_xyzz_fallbacks{0};     // pair rounds skipped: lists did not fit (allocation failure / budget)
    std::atomic<uint64_t> queue_submits{0}, queue_batches{0};
};
extern B200Counters g_counters;

// Stream-ordered scratch (api.cu): blocks are recycled through a per-thread, per-stream cache -- size classes are powers
// of two up to 64 MiB and eighths of an octave above, one budget (half the device memory) for all caches together, everything handed back
// to the driver when an allocation fails -- so a call in steady state makes no allocator call at all.  Reuse on the
// SAME stream needs no synchronisation, and -- unlike the driver's pool -- a block freed by one caller thread is never
// handed to another stream: with hundreds of caller threads the driver pool's cross-stream reuse put milliseconds of
// waits into every small call (tools/queue_bench.cpp: a 256-MSM batch call took 3.4 ms after 16 caller threads had run
// and 12.8 ms after 256; stage events showed the time inside the allocations), and even the GB-sized lists of a 2^24
// MSM cost ~2 ms of stream time per call to obtain from it.
cudaError_t b200_scratch_alloc(void** p, size_t bytes, cudaStream_t stream);
void b200_scratch_free(void* p, size_t bytes, cudaStream_t stream);
void b200_scratch_release_all();     // b200_shutdown
void b200_scratch_register_stream(cudaStream_t s);   // a stream the library created (alive until it says otherwise)
void b200_scratch_forget_stream(cudaStream_t s);     // ... is about to be destroyed

// stream-ordered scratch buffer (freed on the same stream when it goes out of scope)
struct DevBuf {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    size_t bytes = 0;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    cudaError_t alloc(size_t nbytes, cudaStream_t stream) {
        s = stream;
        bytes = nbytes ? nbytes : 16;
        return b200_scratch_alloc(&p, bytes, stream);
    }
    void release() { if (p) b200_scratch_free(p, bytes, s); p = nullptr; }
    ~DevBuf() { release(); }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Optional per-stage timing (b200_profile_begin / b200_profile_end): CUDA events recorded around each stage on
// the launching stream, by the calling thread.  Off by default; costs nothing when off.
struct StageTimer {
    static bool& enabled();
    static void mark(const char* name, cudaStream_t stream);     // start of stage `name` (ends the previous one)
    static void finish(cudaStream_t stream);                      // end of the last stage of a call
};
// NVTX ranges (header-only nvtx3: a no-op costing one pointer check unless a profiler is attached) around every stage
// of a call and around the host <-> device copies, so nsys / ncu timelines read H2D | stage | stage | ... | D2H.
void b200_nvtx_stage(const char* name);                          // closes the calling thread's open stage range, opens `name`
void b200_nvtx_stage_end();
struct NvtxRange {
    explicit NvtxRange(const char* name);
    ~NvtxRange();
};
#define STAGE(name, stream) do { b200_nvtx_stage(name); if (StageTimer::enabled()) StageTimer::mark(name, stream); } while (0)
#define STAGE_END(stream) do { b200_nvtx_stage_end(); if (StageTimer::enabled()) StageTimer::finish(stream); } while (0)

// internal entry points (device pointers, caller-provided stream)
b200_error_t ntt_run_device(void* d_inout, uint32_t log_n, size_t batch, size_t batch_stride, int direction,
                            int coset, cudaStream_t stream);
b200_error_t msm_run_device(void* d_out, const void* d_points, size_t n, const void* d_scalars, size_t stride,
                            const void* d_packed /* non-null: pre-packed bases */, cudaStream_t stream,
                            bool packed_glv = false /* d_packed holds (P_i, phi(P_i)) pairs: msm_pack_bases_glv_device */);
b200_error_t msm_run_batch_device(void* d_out, const void* d_points, size_t n, const void* d_scalars, size_t stride,
                                  const void* d_packed, const unsigned long long* d_seg_off, uint32_t nmsm,
                                  cudaStream_t stream, size_t n_reg = 0, uint32_t c_tab = 0, bool packed_glv = false,
                                  bool shared_bases = false /* every MSM indexes d_packed from 0: k commits, one SRS */);
// resident base sets (api.cu)
struct RegisteredBases {
    void* d_packed;      // n packed points; with a window table: nwin * n (window 0 first)
    size_t n;
    uint32_t c_tab;      // 0: plain; else the window width of the table 2^(c*w) * P_i
    bool glv;            // plain sets: stored as 2n records (P_i, phi(P_i)) for the GLV split (msm_glv.cuh)
};
b200_error_t b200_lookup_bases(uint64_t handle, RegisteredBases* out);
b200_error_t msm_pack_bases_glv_device(void* d_packed_2n, const void* d_points, size_t n, size_t stride, cudaStream_t stream);
bool msm_glv_enabled();
b200_error_t msm_run_tabulated_device(void* d_out, size_t n, const void* d_scalars, const void* d_table, size_t n_reg,
                                      uint32_t c, cudaStream_t stream);
b200_error_t msm_stream_begin(void** session, size_t n_total, cudaStream_t stream);
b200_error_t msm_stream_add(void* session, const void* d_points, size_t n, const void* d_scalars, size_t stride,
                            cudaStream_t stream);
b200_error_t msm_stream_finish(void* session, void* d_out, cudaStream_t stream);
void msm_stream_abort(void* session);
b200_error_t msm_build_window_table_device(void* d_table, size_t n, uint32_t c, cudaStream_t stream);
b200_error_t msm_pack_bases_device(void* d_packed, const void* d_points, size_t n, size_t stride,
                                   cudaStream_t stream);
// host <-> device copies that stay fast for pageable caller memory (hostcopy.cu)
b200_error_t b200_h2d(void* d_dst, const void* h_src, size_t bytes, cudaStream_t stream);
b200_error_t b200_d2h(void* h_dst, const void* d_src, size_t bytes, cudaStream_t stream);
void ntt_release_tables();
void hostcopy_release();             // pinned staging slots (hostcopy.cu)
void b200_queue_shutdown();          // submit / wait dispatcher (queue.cu)
cudaStream_t b200_thread_copy_stream();
cudaStream_t b200_thread_copy_stream2();
b200_error_t b200_require_device();
cudaStream_t b200_thread_stream();
cudaStream_t b200_thread_aux_stream();
