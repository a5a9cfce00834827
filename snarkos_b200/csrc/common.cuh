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

// stream-ordered scratch buffer (freed on the same stream when it goes out of scope)
struct DevBuf {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    cudaError_t alloc(size_t bytes, cudaStream_t stream) {
        s = stream;
        return cudaMallocAsync(&p, bytes ? bytes : 16, stream);
    }
    void release() { if (p) cudaFreeAsync(p, s); p = nullptr; }
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
#define STAGE(name, stream) do { if (StageTimer::enabled()) StageTimer::mark(name, stream); } while (0)
#define STAGE_END(stream) do { if (StageTimer::enabled()) StageTimer::finish(stream); } while (0)

// internal entry points (device pointers, caller-provided stream)
b200_error_t ntt_run_device(void* d_inout, uint32_t log_n, size_t batch, size_t batch_stride, int direction,
                            int coset, cudaStream_t stream);
b200_error_t msm_run_device(void* d_out, const void* d_points, size_t n, const void* d_scalars, size_t stride,
                            const void* d_packed /* non-null: pre-packed bases */, cudaStream_t stream,
                            bool packed_glv = false /* d_packed holds (P_i, phi(P_i)) pairs: msm_pack_bases_glv_device */);
b200_error_t msm_run_batch_device(void* d_out, const void* d_points, size_t n, const void* d_scalars, size_t stride,
                                  const void* d_packed, const unsigned long long* d_seg_off, uint32_t nmsm,
                                  cudaStream_t stream, size_t n_reg = 0, uint32_t c_tab = 0, bool packed_glv = false);
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
b200_error_t b200_require_device();
cudaStream_t b200_thread_stream();
cudaStream_t b200_thread_aux_stream();
