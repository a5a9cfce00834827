// ntt.cu -- kernels + launcher for EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place on B200.
// Algorithm and index arithmetic live in ntt_core.cuh (shared with the host test shim).
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "ntt_core.cuh"
#include "ntt_plan.h"

#define NTT_MAX_LOG_N 28

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
#ifndef NTT_MIN_CTAS
#define NTT_MIN_CTAS 3
#endif
__device__ __forceinline__ void ntt_pass_body(const NttPassParams& p);
__global__ void __launch_bounds__(256, NTT_MIN_CTAS) ntt_pass_kernel(const NttPassParams p) { ntt_pass_body(p); }
// the same pass on 128-thread CTAs (option ntt_variant = 2): with 32 KB tiles six CTAs share an SM instead of three, so
// that more CTAs are in their multiplier-bound stage phases while others load or store
__global__ void __launch_bounds__(128, 6) ntt_pass_kernel_128(const NttPassParams p) { ntt_pass_body(p); }
// Shape-specialised pass (ntt_core.cuh NttShape): 2^L x 2^CW = 1024-element tiles on 128 threads with every tile shift
// and mask known at compile time, ONE code instance of the stage-pair body for all stage pairs (rolled loop: unrolling it
// per stage pair makes the kernel slower -- 3.78 against 3.25 ms at 2^24 -- the hot loop no longer fits the
// instruction cache), and a store phase that contains only what this pass needs (STORE 1: strided pass with its
// boundary table; 2: last pass of a plain forward transform; 0: everything, chosen at run time).
// 1024-element tiles run on 128 threads (six CTAs per SM), 2048-element tiles -- the plans that save a pass with 9-bit
// passes: 2^17, 2^18, 2^25 .. 2^27 -- on 256 threads (three per SM): eight elements per thread either way.
template <int L, int CW, int STORE>
__global__ void __launch_bounds__((1 << (L + CW)) / 8, (L + CW == 10 ? 6 : 3)) ntt_pass_shaped_kernel(const NttPassParams p) {
    constexpr int NT = (1 << (L + CW)) / 8;
    typedef NttShape<L, CW, NT, 1, STORE> SH;
    static_assert(L + CW == 10 || L + CW == 11, "1024- or 2048-element tiles");
    extern __shared__ uint4 sm[];
    const uint32_t tile = blockIdx.x, batch = blockIdx.y, tid = threadIdx.x;
    uint4* sm_tw = sm + 2 * (1 << (L + CW));
    ntt_phase_stage_twiddles<SH>(p, sm_tw, tid, NT);
    const NttTwiddles twd = ntt_shared_twiddles(sm_tw, L);
    ntt_phase_load<SH>(p, sm, tile, batch, tid, NT);
    __syncthreads();
    if (STORE == 0 && p.coset_pre) {                          // the launcher sends forward coset passes to STORE 0
        ntt_phase_coset_pre(p, sm, tile, tid, NT);
        __syncthreads();
    }
#pragma unroll 1
    for (int s = 0; s + 1 < L; s += 2) {
        ntt_phase_stage2<SH>(p, sm, twd, s, tid, NT);
        __syncthreads();
    }
    if (L & 1) {
        ntt_phase_stage<SH>(p, sm, twd, L - 1, tid, NT);
        __syncthreads();
    }
    ntt_phase_store<SH>(p, sm, tile, batch, tid, NT);
}
template <int L, int CW, int STORE> static cudaError_t ntt_shaped_attr() {
    cudaError_t e = cudaFuncSetAttribute(ntt_pass_shaped_kernel<L, CW, STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << (L + CW)) * 32 + (1 << L) * 16);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ntt_pass_shaped_kernel<L, CW, STORE>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    return e;
}
template <int L, int CW> static cudaError_t ntt_shaped_attrs() {
    cudaError_t e = ntt_shaped_attr<L, CW, 0>();
    if (e == cudaSuccess) e = ntt_shaped_attr<L, CW, 1>();
    if (e == cudaSuccess) e = ntt_shaped_attr<L, CW, 2>();
    return e;
}
template <int L, int CW> static void ntt_launch_shaped(int store, dim3 grid, size_t smem, cudaStream_t stream, const NttPassParams& p) {
    constexpr int NT = (1 << (L + CW)) / 8;
    if (store == 1) ntt_pass_shaped_kernel<L, CW, 1><<<grid, NT, smem, stream>>>(p);
    else if (store == 2) ntt_pass_shaped_kernel<L, CW, 2><<<grid, NT, smem, stream>>>(p);
    else ntt_pass_shaped_kernel<L, CW, 0><<<grid, NT, smem, stream>>>(p);
}
__device__ __forceinline__ void ntt_pass_body(const NttPassParams& p) {
    extern __shared__ uint4 sm[];
    const uint32_t tile = blockIdx.x, batch = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;
    const uint32_t L = p.log_len[p.pass];
    uint4* sm_tw = sm + 2 * ((size_t)1 << (L + p.log_cw));         // after the two tile planes: the pass's L/2 twiddles
    ntt_phase_stage_twiddles(p, sm_tw, tid, nt);
    const NttTwiddles twd = ntt_shared_twiddles(sm_tw, L);
    ntt_phase_load(p, sm, tile, batch, tid, nt);
    __syncthreads();
    if (p.coset_pre) {
        ntt_phase_coset_pre(p, sm, tile, tid, nt);
        __syncthreads();
    }
    uint32_t s = 0;
    for (; p.radix4 && s + 1 < L; s += 2) {           // stage pairs on register-resident quads
        ntt_phase_stage2(p, sm, twd, s, tid, nt);
        __syncthreads();
    }
    for (; s < L; s++) {                              // odd length (or radix-2 only): single stages
        ntt_phase_stage(p, sm, twd, s, tid, nt);
        __syncthreads();
    }
    ntt_phase_store(p, sm, tile, batch, tid, nt);
}

// Warp-column pass (ntt_core.cuh): length-2^8 passes over 4 adjacent columns, one warp per column, 8 elements per lane,
// three in-register rounds with warp-level synchronisation only; two block barriers per tile.
__global__ void __launch_bounds__(128, 4) ntt_pass_wc_kernel(const NttPassParams p) {
    extern __shared__ uint4 sm[];
    const uint32_t tile = blockIdx.x, batch = blockIdx.y, tid = threadIdx.x;
    uint4* sm_tw = sm + WC_TILE_U4;
    ntt_phase_stage_twiddles(p, sm_tw, tid, 128);
    const NttTwiddles twd = ntt_shared_twiddles(sm_tw, WC_LOG_LEN);
    wc_phase_load(p, sm, tile, batch, tid, 128);
    __syncthreads();
    const uint32_t cw = tid >> 5, lane = tid & 31u;
    wc_round1(p, sm, twd, tile, cw, lane);
    __syncwarp();
    wc_round2(sm, twd, cw, lane);
    __syncwarp();
    wc_round3(sm, twd, cw, lane);
    __syncthreads();
    wc_phase_store(p, sm, tile, batch, tid, 128);
}
#define NTT_WC_SMEM (WC_TILE_U4 * 16 + (1u << WC_LOG_LEN) * 16)

// ---------------------------------------------------------------------------------------------
// Bulk-copy (TMA) variant: persistent CTAs, two tile buffers, cp.async.bulk (UBLKCP) loads of tile i + 1 completing on
// an mbarrier while tile i is being transformed (option ntt_variant = 1; phases in ntt_core.cuh).
// ---------------------------------------------------------------------------------------------
namespace tma {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// generic-proxy accesses to a buffer ordered before the async-proxy (bulk copy) writes that reuse it
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
}   // namespace tma

#define NTT_BULK_HEADER_U4 8u          // 128 bytes: the two mbarriers
#ifndef NTT_BULK_MIN_CTAS
#define NTT_BULK_MIN_CTAS 3
#endif
__global__ void __launch_bounds__(256, NTT_BULK_MIN_CTAS) ntt_pass_bulk_kernel(const NttPassParams p, uint32_t tiles_per_poly,
                                                                               uint32_t total_tiles) {
    extern __shared__ __align__(128) uint4 sm[];
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    const NttGeom g = ntt_geom(p);
    const uint32_t L = g.log_len;
    const uint32_t buf_elems = ntt_bulk_tile_elems(L, g.log_cw, g.last);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm);
    uint4* sm_tw = sm + NTT_BULK_HEADER_U4;
    uint4* bufs = sm_tw + ((size_t)1 << L);
    if (tid == 0) {
        tma::mbar_init(&bars[0], 1);
        tma::mbar_init(&bars[1], 1);
        tma::fence_barrier_init();
    }
    ntt_phase_stage_twiddles(p, sm_tw, tid, nt);
    const NttTwiddles twd = ntt_shared_twiddles(sm_tw, L);
    __syncthreads();
    const uint32_t nruns = ntt_bulk_runs(p);
    const uint32_t tile_bytes = g.tile_elems * 32u;
    auto issue = [&](uint32_t gtile, uint32_t b) {
        const uint32_t tile = gtile % tiles_per_poly, batch = gtile / tiles_per_poly;
        const uint4* src = p.src + 2ull * batch * p.batch_stride;
        uint4* dstb = bufs + (size_t)b * 2 * buf_elems;
        if (tid == 0) tma::mbar_expect_tx(&bars[b], tile_bytes);
        for (uint32_t r = tid; r < nruns; r += nt) {
            unsigned long long se;
            uint32_t de, cnt;
            ntt_bulk_run(p, tile, r, se, de, cnt);
            tma::bulk_g2s(dstb + 2 * (size_t)de, src + 2 * se, cnt * 32u, &bars[b]);
        }
    };
    uint32_t it = 0;
    if (blockIdx.x < total_tiles) issue(blockIdx.x, 0);
    for (uint32_t gt = blockIdx.x; gt < total_tiles; gt += gridDim.x, it++) {
        const uint32_t b = it & 1u;
        if (gt + gridDim.x < total_tiles) issue(gt + gridDim.x, b ^ 1u);       // next tile lands while this one is transformed
        tma::mbar_wait(&bars[b], (it >> 1) & 1u);
        const uint32_t tile = gt % tiles_per_poly, batch = gt / tiles_per_poly;
        NttBulkTile T;
        T.sm = bufs + (size_t)b * 2 * buf_elems;
        T.last = g.last;
        T.log_len = L;
        T.log_cw = g.log_cw;
        T.h = (tid >> 2) & 1u;
        if (p.coset_pre) {
            ntt_bulk_coset_pre(p, T, tile, tid, nt);
            __syncthreads();
        }
        uint32_t s = 0;
        for (; p.radix4 && s + 1 < L; s += 2) {
            ntt_bulk_stage2(p, T, twd, s, tid, nt);
            __syncthreads();
        }
        for (; s < L; s++) {
            ntt_bulk_stage(p, T, twd, s, tid, nt);
            __syncthreads();
        }
        ntt_bulk_store_out(p, T, tile, batch, tid, nt);
        tma::fence_proxy_async();
        __syncthreads();                         // buffer b is free for the bulk copies of iteration it + 1
    }
}

// out[i] = (base^(2^nsq))^(i << shift)
__global__ void ntt_pow_table_kernel(uint4* out, fr_t base, uint32_t nsq, uint32_t count, uint32_t shift) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    for (uint32_t k = 0; k < nsq; k++) base = fp_sqr(base);
    fr_t r = fp_pow_u64(base, (uint64_t)i << shift);
    uint4 lo, hi;
    fr_to_u4(r, lo, hi);
    out[2 * i] = lo;
    out[2 * i + 1] = hi;
}

// Element (r, c) of a row-major [rows x cols] matrix times w_N^((row_base + r) * (col_base + c)) (kind 0: the
// four-step twiddle between the two transform axes) or element i of a vector times g^(base + i) / g^-(base + i)
// (kind 1: the coset shift of a block of a distributed polynomial).  Powers come from the two-level tables.
__global__ void __launch_bounds__(256) ntt_mul_powers_kernel(uint4* __restrict__ data, const uint4* __restrict__ lo,
                                                             const uint4* __restrict__ hi, unsigned long long rows,
                                                             unsigned long long cols, unsigned long long row_base,
                                                             unsigned long long col_base, uint32_t log_n, int kind) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    unsigned long long e;
    if (kind == 0) {
        const unsigned long long r = i / cols, c = i - r * cols;
        e = ((row_base + r) * (col_base + c)) & ((1ull << log_n) - 1);
    } else {
        e = row_base + i;
    }
    fr_t x = fr_load(data, i);
    if (e) x = fp_mul(x, pow2level(lo, hi, e));
    uint4 a, b;
    fr_to_u4(x, a, b);
    data[2 * i] = a;
    data[2 * i + 1] = b;
}

// boundary table of one pass: out[k * S + m] = w_M^(m * k), M = 2^log_sub, S = 2^log_stride (M entries)
__global__ void __launch_bounds__(256) ntt_boundary_table_kernel(uint4* __restrict__ out, const uint4* __restrict__ lo,
                                                                 const uint4* __restrict__ hi, uint32_t log_n,
                                                                 uint32_t log_sub, uint32_t log_stride) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >> log_sub) return;
    const unsigned long long k = i >> log_stride, m = i & ((1ull << log_stride) - 1);
    const unsigned long long ex = ((m * k) & ((1ull << log_sub) - 1)) << (log_n - log_sub);
    fr_t w = pow2level(lo, hi, ex);
    uint4 a, b;
    fr_to_u4(w, a, b);
    out[2 * i] = a;
    out[2 * i + 1] = b;
}

// out = (2^log_n)^-1 in Montgomery form
__global__ void ntt_size_inv_kernel(fr_t* out, uint32_t log_n) {
    fr_t n = fp_zero<FrP>();
    n.v[log_n >> 5] = 1u << (log_n & 31);
    *out = fp_inv(fp_to_mont(n));
}

// ---------------------------------------------------------------------------------------------
// cached per-(log_n, direction) tables
// ---------------------------------------------------------------------------------------------
#define NTT_BOUNDARY_TABLE_MAX_LOG 27          // up to 2^27 entries = 4 GiB per pass boundary (180 GB of HBM); larger -> two-level powers
struct NttDomainTables {
    // per-pass inter-pass twiddle tables, valid for the plan recorded next to them
    uint4* boundary[NTT_MAX_PASSES] = {nullptr, nullptr, nullptr, nullptr};
    uint32_t boundary_plan[NTT_MAX_PASSES] = {0, 0, 0, 0};
    uint32_t boundary_npasses = 0;
    uint4* pow_lo = nullptr;
    uint4* pow_hi = nullptr;
    uint4* coset_lo = nullptr;
    uint4* coset_hi = nullptr;
    fr_t size_inv;
};
struct NttState {
    std::mutex mu;
    uint4* tile_tw[2] = {nullptr, nullptr};
    std::map<uint32_t, NttDomainTables> domains;     // key = log_n * 2 + direction
    bool smem_attr_set = false;
    int sm_count = 148;
};
static NttState g_ntt;

static fr_t fr_const(const uint32_t* limbs) {
    fr_t r;
    memcpy(r.v, limbs, sizeof(r.v));
    return r;
}

static b200_error_t build_pow_table(uint4** out, const uint32_t* base_limbs, uint32_t nsq, uint32_t count,
                                    uint32_t shift, cudaStream_t stream) {
    CUDA_TRY(cudaMalloc(out, (size_t)count * 32));
    ntt_pow_table_kernel<<<(count + 127) / 128, 128, 0, stream>>>(*out, fr_const(base_limbs), nsq, count, shift);
    KERNEL_CHECK();
    return b200_ok();
}

struct NttExchangeParams;
__global__ void ntt_exchange_transpose_kernel(NttExchangeParams p);

static b200_error_t get_tables(uint32_t log_n, int direction, cudaStream_t stream, NttDomainTables* out,
                               const uint4** tile_tw, const NttPlan* plan = nullptr) {
    std::lock_guard<std::mutex> lock(g_ntt.mu);
    if (!g_ntt.smem_attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(ntt_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (1 << NTT_MAX_TILE_LOG) * 32 + (1 << (NTT_MAX_TILE_LOG - 1)) * 32));
        // same shared-memory carve-out for the pass kernel and the exchange kernel: CTAs of kernels that ask for
        // different L1 / shared splits do not share an SM, which serialises the two streams of the overlapped
        // multi-GPU schedule (measured: NTT 4.01 ms + exchange 0.38 ms = 4.39 ms "concurrently")
        CUDA_TRY(cudaFuncSetAttribute(ntt_pass_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CUDA_TRY(cudaFuncSetAttribute(ntt_pass_kernel_128, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (1 << NTT_MAX_TILE_LOG) * 32 + (1 << (NTT_MAX_TILE_LOG - 1)) * 32));
        CUDA_TRY(cudaFuncSetAttribute(ntt_pass_kernel_128, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CUDA_TRY((ntt_shaped_attrs<8, 2>()));
        CUDA_TRY((ntt_shaped_attrs<7, 3>()));
        CUDA_TRY((ntt_shaped_attrs<6, 4>()));
        CUDA_TRY((ntt_shaped_attrs<9, 2>()));
        CUDA_TRY((ntt_shaped_attrs<8, 3>()));
        CUDA_TRY(cudaFuncSetAttribute(ntt_pass_wc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NTT_WC_SMEM));
        CUDA_TRY(cudaFuncSetAttribute(ntt_pass_wc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CUDA_TRY(cudaFuncSetAttribute(ntt_pass_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        CUDA_TRY(cudaFuncSetAttribute(ntt_pass_bulk_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        {
            int dev = 0;
            CUDA_TRY(cudaGetDevice(&dev));
            CUDA_TRY(cudaDeviceGetAttribute(&g_ntt.sm_count, cudaDevAttrMultiProcessorCount, dev));
        }
        CUDA_TRY(cudaFuncSetAttribute(ntt_exchange_transpose_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        g_ntt.smem_attr_set = true;
    }
    const uint32_t* root = direction ? FR_TWO_ADIC_ROOT_INV : FR_TWO_ADIC_ROOT;
    if (!g_ntt.tile_tw[direction]) {
        B200_TRY(build_pow_table(&g_ntt.tile_tw[direction], root, FR_TWO_ADICITY - NTT_TILE_TW_LOG,
                                 1u << (NTT_TILE_TW_LOG - 1), 0, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
    }
    *tile_tw = g_ntt.tile_tw[direction];
    uint32_t key = log_n * 2 + (uint32_t)direction;
    auto it = g_ntt.domains.find(key);
    if (it == g_ntt.domains.end()) {
        NttDomainTables t;
        uint32_t lo_count = 1u << NTT_POW_LO_LOG;
        uint32_t hi_count = log_n > NTT_POW_LO_LOG ? (1u << (log_n - NTT_POW_LO_LOG)) : 1u;
        B200_TRY(build_pow_table(&t.pow_lo, root, FR_TWO_ADICITY - log_n, lo_count, 0, stream));
        B200_TRY(build_pow_table(&t.pow_hi, root, FR_TWO_ADICITY - log_n, hi_count, NTT_POW_LO_LOG, stream));
        const uint32_t* g = direction ? FR_GENERATOR_INV : FR_GENERATOR;
        B200_TRY(build_pow_table(&t.coset_lo, g, 0, lo_count, 0, stream));
        B200_TRY(build_pow_table(&t.coset_hi, g, 0, hi_count, NTT_POW_LO_LOG, stream));
        fr_t* d_inv = nullptr;
        CUDA_TRY(cudaMalloc(&d_inv, sizeof(fr_t)));
        ntt_size_inv_kernel<<<1, 1, 0, stream>>>(d_inv, log_n);
        KERNEL_CHECK();
        CUDA_TRY(cudaMemcpyAsync(&t.size_inv, d_inv, sizeof(fr_t), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        CUDA_TRY(cudaFree(d_inv));
        it = g_ntt.domains.emplace(key, t).first;
    }
    if (plan && plan->npasses > 1 && b200_config().ntt_boundary_tables) {
        NttDomainTables& t = it->second;
        bool same = t.boundary_npasses == plan->npasses;
        for (uint32_t i = 0; same && i < plan->npasses; i++) same = (t.boundary_plan[i] == plan->log_len[i]);
        if (!same) {                                  // (re)build for this plan; happens once per (size, direction, plan)
            CUDA_TRY(cudaStreamSynchronize(stream));
            for (uint32_t i = 0; i < NTT_MAX_PASSES; i++) {
                if (t.boundary[i]) CUDA_TRY(cudaFree(t.boundary[i]));
                t.boundary[i] = nullptr;
                t.boundary_plan[i] = i < plan->npasses ? plan->log_len[i] : 0;
            }
            t.boundary_npasses = plan->npasses;
            uint32_t before = 0;
            for (uint32_t i = 0; i + 1 < plan->npasses; i++) {
                const uint32_t log_sub = log_n - before, log_stride = log_sub - plan->log_len[i];
                before += plan->log_len[i];
                if (log_sub > NTT_BOUNDARY_TABLE_MAX_LOG) continue;
                const size_t count = (size_t)1 << log_sub;
                CUDA_TRY(cudaMalloc(&t.boundary[i], count * 32));
                ntt_boundary_table_kernel<<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(t.boundary[i], t.pow_lo, t.pow_hi,
                                                                                             log_n, log_sub, log_stride);
                KERNEL_CHECK();
            }
            CUDA_TRY(cudaStreamSynchronize(stream));
        }
    }
    *out = it->second;
    return b200_ok();
}

// Multi-GPU four-step helpers (see snarkos_b200/dist.py): twiddle / coset scaling of a block of a distributed
// polynomial of total size 2^log_n.
extern "C" b200_error_t b200_fr_mul_powers_device(void* d_data, uint32_t log_n, int direction, int kind,
                                                  unsigned long long rows, unsigned long long cols,
                                                  unsigned long long row_base, unsigned long long col_base, void* stream) {
    B200_TRY(b200_require_device());
    if (log_n > NTT_MAX_LOG_N || (direction != 0 && direction != 1) || (kind != 0 && kind != 1))
        return b200_err(B200_ERR_INVALID_ARG, "fr_mul_powers: bad argument");
    if (rows == 0 || cols == 0) return b200_ok();
    if (!d_data) return b200_err(B200_ERR_INVALID_ARG, "fr_mul_powers: null pointer");
    if (kind == 1 && row_base + rows * cols > (1ull << log_n)) return b200_err(B200_ERR_INVALID_ARG, "fr_mul_powers: coset range exceeds the domain");
    NttDomainTables tabs;
    const uint4* tile_tw = nullptr;
    cudaStream_t s = (cudaStream_t)stream;
    B200_TRY(get_tables(log_n, direction, s, &tabs, &tile_tw));
    const unsigned long long total = rows * cols;
    ntt_mul_powers_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
        reinterpret_cast<uint4*>(d_data), kind == 0 ? tabs.pow_lo : tabs.coset_lo, kind == 0 ? tabs.pow_hi : tabs.coset_hi, rows,
        cols, row_base, col_base, log_n, kind);
    KERNEL_CHECK();
    return b200_ok();
}

// ---------------------------------------------------------------------------------------------
// Fused exchange of the multi-GPU four-step NTT: transpose + all-to-all (+ the twiddle between the two transform
// axes) in ONE pass.  The local slab [r_local x c] of a row-distributed matrix is read in 32 x 32 tiles (coalesced
// rows), optionally multiplied by w_N^(row * col), transposed through shared memory and every column of the tile is
// written as a run of 32 x 32 B straight to its place in the [c / world x r_local * world] slab of the rank that owns
// that column -- peer memory over NVLink when that rank is another GPU (dst[d] = that rank's buffer, opened through
// CUDA IPC).  Replaces: local re-tiling pass, NCCL all-to-all, second re-tiling pass and the separate twiddle pass.
// ---------------------------------------------------------------------------------------------
#define NTT_XCHG_TILE 32
struct NttExchangeParams {
    const uint4* src;               // first row of the part of the slab this launch handles
    uint4* dst[NTT_XCHG_MAX_WORLD];
    const uint4* pow_lo;
    const uint4* pow_hi;
    unsigned long long r_total, row_off, r_count;   // slab rows per rank, first row / number of rows of this launch
    unsigned long long c, c_local, col_lo, col_cnt; // columns [col_lo, col_lo + col_cnt) of EVERY destination's range
    unsigned long long row_base;                    // global index of slab row 0 (twiddle exponent)
    uint32_t rank, world, log_n, twiddle;
};

// A launch may cover a row range and, for every destination rank, a sub-range of its columns, so that a slab can be
// exchanged in chunks while the transforms of the chunks that have arrived are already running.
// Grid-stride over the tiles: with a bounded grid (cta_limit) the kernel keeps only a few CTAs per SM, so that the
// CTAs of a concurrently running transform still fit -- an NVLink-bound exchange needs bytes in flight, not SM slots
// (an unbounded high-priority grid fills every SM with CTAs that wait on the link and starves the transform).
__global__ void __launch_bounds__(256) ntt_exchange_transpose_kernel(NttExchangeParams p) {
    __shared__ uint4 sm_lo[NTT_XCHG_TILE][NTT_XCHG_TILE + 1];
    __shared__ uint4 sm_hi[NTT_XCHG_TILE][NTT_XCHG_TILE + 1];
    const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const unsigned long long vtotal = p.col_cnt * p.world;
    const unsigned long long tiles_x = (vtotal + NTT_XCHG_TILE - 1) / NTT_XCHG_TILE;
    const unsigned long long tiles_y = (p.r_count + NTT_XCHG_TILE - 1) / NTT_XCHG_TILE;
    const unsigned long long R = p.r_total * p.world;
    for (unsigned long long tile = blockIdx.x; tile < tiles_x * tiles_y; tile += gridDim.x) {
        // virtual column v = d * col_cnt + j  <->  column d * c_local + col_lo + j
        const unsigned long long v0 = (tile % tiles_x) * NTT_XCHG_TILE, r0 = (tile / tiles_x) * NTT_XCHG_TILE;
        for (uint32_t rr = ty; rr < NTT_XCHG_TILE; rr += 8) {
            const unsigned long long row = r0 + rr, v = v0 + tx;
            if (row < p.r_count && v < vtotal) {
                const unsigned long long d = v / p.col_cnt, col = d * p.c_local + p.col_lo + (v - d * p.col_cnt);
                fr_t x;
                ptx::ld_global_256(p.src + 2 * (row * p.c + col), x.v);
                if (p.twiddle) {
                    const unsigned long long e = ((p.row_base + p.row_off + row) * col) & ((1ull << p.log_n) - 1);
                    if (e) x = fp_mul(x, pow2level(p.pow_lo, p.pow_hi, e));
                }
                fr_to_u4(x, sm_lo[rr][tx], sm_hi[rr][tx]);
            }
        }
        __syncthreads();
        for (uint32_t cc = ty; cc < NTT_XCHG_TILE; cc += 8) {
            const unsigned long long v = v0 + cc, row = r0 + tx;
            if (row < p.r_count && v < vtotal) {
                const unsigned long long d = v / p.col_cnt, cl = p.col_lo + (v - d * p.col_cnt);
                uint4* out = p.dst[d] + 2 * (cl * R + (unsigned long long)p.rank * p.r_total + p.row_off + row);
                // one 256-bit store per element: a warp writes 1 KB of whole sectors per instruction (16-byte halves
                // with 16-byte gaps travel over NVLink as many small packets)
                const uint4 lo = sm_lo[tx][cc], hi = sm_hi[tx][cc];
                const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
                ptx::st_global_256(out, w);
            }
        }
        __syncthreads();
    }
}

extern "C" b200_error_t b200_fr_exchange_transpose_part_device(const void* d_src_slab, void* const* dst_ptrs, uint32_t world,
                                                               uint32_t rank, unsigned long long r_local,
                                                               unsigned long long row_off, unsigned long long r_count,
                                                               unsigned long long c, unsigned long long col_lo,
                                                               unsigned long long col_cnt, uint32_t log_n, int direction,
                                                               int twiddle, unsigned long long row_base, uint32_t cta_limit,
                                                               void* stream) {
    B200_TRY(b200_require_device());
    if (world == 0 || world > NTT_XCHG_MAX_WORLD || rank >= world || (direction != 0 && direction != 1) || log_n > NTT_MAX_LOG_N)
        return b200_err(B200_ERR_INVALID_ARG, "fr_exchange_transpose: bad argument");
    if (r_count == 0 || col_cnt == 0) return b200_ok();
    if (c % world) return b200_err(B200_ERR_INVALID_ARG, "fr_exchange_transpose: columns must divide by the world size");
    if (row_off + r_count > r_local || col_lo + col_cnt > c / world)
        return b200_err(B200_ERR_INVALID_ARG, "fr_exchange_transpose: part outside the slab");
    if (!d_src_slab || !dst_ptrs) return b200_err(B200_ERR_INVALID_ARG, "fr_exchange_transpose: null pointer");
    NttExchangeParams p;
    memset(&p, 0, sizeof(p));
    cudaStream_t s = (cudaStream_t)stream;
    if (twiddle) {
        NttDomainTables tabs;
        const uint4* tile_tw = nullptr;
        B200_TRY(get_tables(log_n, direction, s, &tabs, &tile_tw));
        p.pow_lo = tabs.pow_lo;
        p.pow_hi = tabs.pow_hi;
    }
    p.src = reinterpret_cast<const uint4*>(d_src_slab) + 2 * row_off * c;
    for (uint32_t d = 0; d < world; d++) {
        if (!dst_ptrs[d]) return b200_err(B200_ERR_INVALID_ARG, "fr_exchange_transpose: null destination");
        p.dst[d] = reinterpret_cast<uint4*>(dst_ptrs[d]);
    }
    p.r_total = r_local;
    p.row_off = row_off;
    p.r_count = r_count;
    p.c = c;
    p.c_local = c / world;
    p.col_lo = col_lo;
    p.col_cnt = col_cnt;
    p.row_base = row_base;
    p.rank = rank;
    p.world = world;
    p.log_n = log_n;
    p.twiddle = twiddle ? 1u : 0u;
    const unsigned long long vtotal = col_cnt * world;
    unsigned long long ntiles = ((vtotal + NTT_XCHG_TILE - 1) / NTT_XCHG_TILE) * ((r_count + NTT_XCHG_TILE - 1) / NTT_XCHG_TILE);
    if (ntiles >= (1ull << 31)) return b200_err(B200_ERR_TOO_LARGE, "fr_exchange_transpose: too many tiles per launch");
    if (cta_limit && ntiles > cta_limit) ntiles = cta_limit;
    ntt_exchange_transpose_kernel<<<(unsigned)ntiles, 256, 0, s>>>(p);
    KERNEL_CHECK();
    return b200_ok();
}

extern "C" b200_error_t b200_fr_exchange_transpose_device(const void* d_src, void* const* dst_ptrs, uint32_t world,
                                                          uint32_t rank, unsigned long long r_local, unsigned long long c,
                                                          uint32_t log_n, int direction, int twiddle,
                                                          unsigned long long row_base, void* stream) {
    if (world == 0 || c % world) return b200_err(B200_ERR_INVALID_ARG, "fr_exchange_transpose: columns must divide by the world size");
    return b200_fr_exchange_transpose_part_device(d_src, dst_ptrs, world, rank, r_local, 0, r_local, c, 0, c / world, log_n,
                                                  direction, twiddle, row_base, 0, stream);
}

// Exchange buffers other ranks' GPUs write into: plain cudaMalloc memory exported / opened through CUDA IPC
extern "C" b200_error_t b200_peer_buffer_alloc(size_t bytes, void** d_ptr, void* handle64) {
    B200_TRY(b200_require_device());
    if (!d_ptr || !handle64) return b200_err(B200_ERR_INVALID_ARG, "peer_buffer_alloc: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
    CUDA_TRY(cudaMalloc(d_ptr, bytes ? bytes : 16));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, *d_ptr);
    if (e != cudaSuccess) { cudaFree(*d_ptr); *d_ptr = nullptr; return b200_cuda_err(e); }
    memcpy(handle64, &h, 64);
    return b200_ok();
}
extern "C" b200_error_t b200_peer_buffer_open(const void* handle64, void** d_ptr) {
    B200_TRY(b200_require_device());
    if (!d_ptr || !handle64) return b200_err(B200_ERR_INVALID_ARG, "peer_buffer_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CUDA_TRY(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return b200_ok();
}
extern "C" b200_error_t b200_peer_buffer_close(void* d_ptr) {
    if (d_ptr) CUDA_TRY(cudaIpcCloseMemHandle(d_ptr));
    return b200_ok();
}
extern "C" b200_error_t b200_peer_buffer_free(void* d_ptr) {
    if (d_ptr) CUDA_TRY(cudaFree(d_ptr));
    return b200_ok();
}

void ntt_release_tables() {
    std::lock_guard<std::mutex> lock(g_ntt.mu);
    for (int d = 0; d < 2; d++) {
        if (g_ntt.tile_tw[d]) cudaFree(g_ntt.tile_tw[d]);
        g_ntt.tile_tw[d] = nullptr;
    }
    for (auto& kv : g_ntt.domains) {
        for (int i = 0; i < NTT_MAX_PASSES; i++)
            if (kv.second.boundary[i]) cudaFree(kv.second.boundary[i]);
        cudaFree(kv.second.pow_lo);
        cudaFree(kv.second.pow_hi);
        cudaFree(kv.second.coset_lo);
        cudaFree(kv.second.coset_hi);
    }
    g_ntt.domains.clear();
}

static bool ntt_wc_applicable_host(const NttPlan& plan, uint32_t i) { return plan.log_len[i] == WC_LOG_LEN && plan.log_cw[i] == WC_LOG_CW; }

// fused exchange of the LAST pass (multi-GPU four-step transform): see NttPassParams::xchg in ntt_core.cuh
struct NttXchg {
    void* const* dst_ptrs;          // host array of `world` device pointers
    uint32_t world, rank, log_n_total;
    int twiddle;
    unsigned long long row_base;
};
static b200_error_t ntt_run_device_x(void* d_inout, uint32_t log_n, size_t batch, size_t batch_stride, int direction, int coset,
                                     cudaStream_t stream, const NttXchg* xc);
b200_error_t ntt_run_device(void* d_inout, uint32_t log_n, size_t batch, size_t batch_stride, int direction,
                            int coset, cudaStream_t stream) {
    return ntt_run_device_x(d_inout, log_n, batch, batch_stride, direction, coset, stream, nullptr);
}
static b200_error_t ntt_run_device_x(void* d_inout, uint32_t log_n, size_t batch, size_t batch_stride, int direction, int coset,
                                     cudaStream_t stream, const NttXchg* xc) {
    if (log_n > NTT_MAX_LOG_N) return b200_err(B200_ERR_TOO_LARGE, "ntt: log_n > 28 is not supported");
    if (direction != 0 && direction != 1) return b200_err(B200_ERR_INVALID_ARG, "ntt: direction must be 0 or 1");
    if (batch == 0) return b200_ok();
    if (!d_inout) return b200_err(B200_ERR_INVALID_ARG, "ntt: null data pointer");
    const size_t n = (size_t)1 << log_n;
    if (batch > 1 && batch_stride < n) return b200_err(B200_ERR_INVALID_ARG, "ntt: batch stride smaller than the domain");
    if (batch > 65535) return b200_err(B200_ERR_TOO_LARGE, "ntt: batch > 65535");
    if (log_n == 0 && !direction) return b200_ok();      // size-1 domain: identity (n^-1 = 1, g^0 = 1)
    if (log_n == 0) return b200_ok();

    NttPlan plan;
    // tile size: 1024 elements (32 KB, six 128-thread CTAs per SM) unless 2048-element tiles save a whole pass
    // (2^26: 9 + 9 + 8 in three passes against four with 8-bit passes; measured 16.3 vs 18.5 ms)
    uint32_t tile_log = (uint32_t)b200_config().ntt_tile_log;
    if (tile_log == 0) tile_log = (log_n + 8) / 9 < (log_n + 7) / 8 ? 11 : 10;
    if (!ntt_make_plan(log_n, &plan, tile_log, b200_config().ntt_plan)) return b200_err(B200_ERR_INVALID_ARG, "ntt: no valid pass plan");
    NttDomainTables tabs;
    const uint4* tile_tw = nullptr;
    B200_TRY(get_tables(log_n, direction, stream, &tabs, &tile_tw, &plan));
    NttDomainTables xtabs;
    if (xc) {
        // last pass in "adjacent rows" tiles: 2^cw rows of one sub-problem each (1024-element tiles, at most 16 rows)
        const uint32_t ll = plan.log_len[plan.npasses - 1];
        uint32_t cw = ll >= 10 ? 0 : 10 - ll;
        if (cw > 4) cw = 4;
        while (cw && (batch & ((1u << cw) - 1))) cw--;
        plan.log_cw[plan.npasses - 1] = cw;
        if (xc->twiddle) {
            const uint4* unused = nullptr;
            B200_TRY(get_tables(xc->log_n_total, direction, stream, &xtabs, &unused));
        }
    }

    DevBuf scratch;
    if (plan.npasses > 1) CUDA_TRY(scratch.alloc(((batch - 1) * batch_stride + n) * 32, stream));

    for (uint32_t i = 0; i < plan.npasses; i++) {
        NttPassParams p;
        memset(&p, 0, sizeof(p));
        const bool first = (i == 0), last = (i + 1 == plan.npasses);
        p.src = reinterpret_cast<const uint4*>(first ? d_inout : scratch.p);
        p.dst = reinterpret_cast<uint4*>(last ? d_inout : scratch.p);
        p.tile_tw = tile_tw;
        p.pow_lo = tabs.pow_lo;
        p.pow_hi = tabs.pow_hi;
        p.boundary_tw = tabs.boundary[i];
        p.coset_lo = tabs.coset_lo;
        p.coset_hi = tabs.coset_hi;
        p.size_inv = tabs.size_inv;
        p.batch_stride = batch_stride;
        p.log_n = log_n;
        p.pass = i;
        p.npasses = plan.npasses;
        for (uint32_t k = 0; k < NTT_MAX_PASSES; k++) p.log_len[k] = plan.log_len[k];
        p.log_cw = plan.log_cw[i];
        p.coset_pre = (first && coset && direction == 0) ? 1 : 0;
        p.scale_post = (last && direction == 1) ? 1 : 0;
        p.coset_post = (last && coset && direction == 1) ? 1 : 0;
        p.radix4 = b200_config().ntt_radix4 ? 1 : 0;
        const bool xpass = xc && last;
        if (xpass) {
            p.xchg = 1;
            p.x_world = xc->world;
            p.x_rank = xc->rank;
            p.x_twiddle = xc->twiddle ? 1 : 0;
            p.x_log_n = xc->log_n_total;
            p.x_r_total = batch;
            p.x_row_base = xc->row_base;
            p.x_pow_lo = xtabs.pow_lo;
            p.x_pow_hi = xtabs.pow_hi;
            for (uint32_t d = 0; d < xc->world; d++) p.x_dst[d] = reinterpret_cast<uint4*>(xc->dst_ptrs[d]);
        }
        const uint32_t pass_tile_log = plan.log_len[i] + plan.log_cw[i];
        const uint32_t tile_elems = 1u << pass_tile_log;
        uint32_t threads = tile_elems / 2;
        if (threads > 256) threads = 256;
        if (threads < 32) threads = 32;
        static const char* const kPassName[NTT_MAX_PASSES] = {"ntt_pass0", "ntt_pass1", "ntt_pass2", "ntt_pass3"};
        STAGE(kPassName[i], stream);
        const uint32_t tiles_per_poly = 1u << (log_n - pass_tile_log);
        const unsigned long long total_tiles = (unsigned long long)tiles_per_poly * batch;
        if (b200_config().ntt_variant == 1 && threads == 256 && total_tiles < (1ull << 31) && !xpass) {
            // bulk-copy (TMA) variant: persistent CTAs over all tiles of the batch, two tile buffers
            const uint32_t buf_elems = ntt_bulk_tile_elems(plan.log_len[i], plan.log_cw[i], last ? 1u : 0u);
            const size_t smem = NTT_BULK_HEADER_U4 * 16 + ((size_t)1 << plan.log_len[i]) * 16 + 2 * (size_t)buf_elems * 32;
            int per_sm = 0;
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ntt_pass_bulk_kernel, 256, smem));
            if (per_sm >= 1) {
                unsigned long long ctas = (unsigned long long)g_ntt.sm_count * per_sm;
                if (ctas > total_tiles) ctas = total_tiles;
                ntt_pass_bulk_kernel<<<(unsigned)ctas, 256, smem, stream>>>(p, tiles_per_poly, (uint32_t)total_tiles);
                KERNEL_CHECK();
                continue;
            }
        }
        dim3 grid(tiles_per_poly, (unsigned)batch);
        if (xpass) {
            // one tile = one sub-problem of 2^cw adjacent rows; generic kernels (run-time flags), 8 elements per thread
            grid = dim3(1u << (log_n - plan.log_len[i]), (unsigned)(batch >> plan.log_cw[i]));
            const size_t xsmem = (size_t)tile_elems * 32 + ((size_t)1 << plan.log_len[i]) * 16;
            const bool xshaped = b200_config().ntt_variant == 0 && b200_config().ntt_radix4;
            if (xshaped && plan.log_len[i] == 6 && plan.log_cw[i] == 4) ntt_launch_shaped<6, 4>(0, grid, xsmem, stream, p);
            else if (xshaped && plan.log_len[i] == 7 && plan.log_cw[i] == 3) ntt_launch_shaped<7, 3>(0, grid, xsmem, stream, p);
            else if (xshaped && plan.log_len[i] == 8 && plan.log_cw[i] == 2) ntt_launch_shaped<8, 2>(0, grid, xsmem, stream, p);
            else if (tile_elems == 1024) ntt_pass_kernel_128<<<grid, 128, xsmem, stream>>>(p);
            else ntt_pass_kernel<<<grid, threads, xsmem, stream>>>(p);
            KERNEL_CHECK();
            continue;
        }
        // Warp-column kernel: built, bit-exact, measured and NOT the default -- 4.12 ms against 3.43 ms at 2^24 (ncu
        // profiles/r02_ncu_ntt_wc.txt: 30 k instructions of straight-line code per tile, `no_instruction` stalls 2.9 - 3.6
        // per issue, 16 warps per SM at 128 registers, multiplier pipe 66 - 70 % busy); option ntt_variant = 4 selects it.
        if (b200_config().ntt_variant == 4 && ntt_wc_applicable_host(plan, i)) {
            ntt_pass_wc_kernel<<<grid, 128, NTT_WC_SMEM, stream>>>(p);
            KERNEL_CHECK();
            continue;
        }
        const size_t smem = (size_t)tile_elems * 32 + ((size_t)1 << plan.log_len[i]) * 16;   // tile + L/2 twiddles
        // 1024-element tiles (the default): 128-thread CTAs, six per SM (measured 3.47 vs 3.57 ms at 2^24, 2.98 vs 3.26 ms at
        // 2^20 x 16 against 256-thread CTAs on 2048-element tiles); ntt_variant = 3 forces the 256-thread CTAs
        // shape-specialised kernels for the 1024-element tiles of the default plans (ntt_variant 2 / 3: generic kernels)
        const bool shaped = b200_config().ntt_variant == 0 && b200_config().ntt_radix4;
        const int store = (p.coset_pre || p.scale_post || p.coset_post) ? 0 : last ? (plan.npasses > 1 ? 2 : 0) : (p.boundary_tw ? 1 : 0);
        if (shaped && plan.log_len[i] == 8 && plan.log_cw[i] == 2)
            ntt_launch_shaped<8, 2>(store, grid, smem, stream, p);
        else if (shaped && plan.log_len[i] == 7 && plan.log_cw[i] == 3)
            ntt_launch_shaped<7, 3>(store, grid, smem, stream, p);
        else if (shaped && plan.log_len[i] == 6 && plan.log_cw[i] == 4)
            ntt_launch_shaped<6, 4>(store, grid, smem, stream, p);
        else if (shaped && plan.log_len[i] == 9 && plan.log_cw[i] == 2)
            ntt_launch_shaped<9, 2>(store, grid, smem, stream, p);
        else if (shaped && plan.log_len[i] == 8 && plan.log_cw[i] == 3)
            ntt_launch_shaped<8, 3>(store, grid, smem, stream, p);
        else if (tile_elems == 1024 && b200_config().ntt_variant != 3)
            ntt_pass_kernel_128<<<grid, 128, smem, stream>>>(p);
        else
            ntt_pass_kernel<<<grid, threads, smem, stream>>>(p);
        KERNEL_CHECK();
    }
    STAGE_END(stream);
    return b200_ok();
}

// Batch of row transforms whose last pass IS the exchange (see NttPassParams::xchg): `rows` rows of 2^log_len elements at
// d_rows (row-major, not modified) are transformed and every output (row, col) is stored -- times w_N^((row_base + row) * col)
// when `twiddle` -- at its place in the transposed slab of the rank that owns column col.
extern "C" b200_error_t b200_ntt_rows_exchange_device(const void* d_rows, void* const* dst_ptrs, uint32_t world, uint32_t rank,
                                                      unsigned long long rows, uint32_t log_len, uint32_t log_n_total,
                                                      int direction, int twiddle, unsigned long long row_base, void* stream) {
    B200_TRY(b200_require_device());
    if (world == 0 || world > NTT_XCHG_MAX_WORLD || rank >= world || log_n_total > NTT_MAX_LOG_N || log_len == 0 || log_len > log_n_total)
        return b200_err(B200_ERR_INVALID_ARG, "ntt_rows_exchange: bad argument");
    if (rows == 0) return b200_ok();
    if (!d_rows || !dst_ptrs) return b200_err(B200_ERR_INVALID_ARG, "ntt_rows_exchange: null pointer");
    if (((1ull << log_len) % world) != 0) return b200_err(B200_ERR_INVALID_ARG, "ntt_rows_exchange: row length must divide by the world size");
    for (uint32_t d = 0; d < world; d++)
        if (!dst_ptrs[d]) return b200_err(B200_ERR_INVALID_ARG, "ntt_rows_exchange: null destination");
    NttXchg xc;
    xc.dst_ptrs = dst_ptrs;
    xc.world = world;
    xc.rank = rank;
    xc.log_n_total = log_n_total;
    xc.twiddle = twiddle;
    xc.row_base = row_base;
    return ntt_run_device_x(const_cast<void*>(d_rows), log_len, (size_t)rows, (size_t)1 << log_len, direction, 0, (cudaStream_t)stream, &xc);
}

