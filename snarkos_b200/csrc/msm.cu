// msm.cu -- kernels + launcher for VariableBase::msm over BLS12-377 G1 on B200.
//
// Pipeline (all on the device, one stream):
//   pack      : G1Affine images (stride 104) -> 96-byte (x | y), infinity -> (0, 0)
//   count     : signed c-bit digits of every scalar; histogram of (window, bucket)        [atomics in L2]
//   scan      : exclusive prefix sum of the histogram -> start offset of every bucket
//   scatter   : point index (+ sign bit) of every non-zero digit into its bucket's slot      [counting sort]
//   accumulate: buckets cut into tasks of bounded length; one thread per task, XYZZ mixed additions over its
//               slice of the sorted point list; pairwise rounds combine the partial sums of split buckets
//   reduce    : per window, segmented running sum  sum_b (b + 1) * B_b ; block tree per window
//   fold      : Horner over the windows with c doublings each -> Jacobian result
// HBM layout: packed bases n * 96 B | entries n * nwin * 4 B | offsets (nwin * 2^(c-1) + 1) * 4 B |
//             buckets nwin * 2^(c-1) * 192 B (XYZZ) | segment sums | window sums.
#include <cstdlib>

#include "common.cuh"
#include "msm_core.cuh"

#define MSM_ACC_THREADS 128
#define MSM_RED_THREADS 64
#define MSM_SEG_LEN 32u
#define MSM_TREE_THREADS 128

// ---------------------------------------------------------------------------------------------
// pack / count / scatter
// ---------------------------------------------------------------------------------------------
__global__ void msm_pack_kernel(g1_packed_t* __restrict__ out, const uint8_t* __restrict__ pts, size_t n,
                                size_t stride) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* src = pts + i * stride;
    // the Rust side guarantees 8-byte alignment of G1Affine; read as 12 x u64
    const unsigned long long* s64 = reinterpret_cast<const unsigned long long*>(src);
    unsigned long long v[12];
#pragma unroll
    for (int k = 0; k < 12; k++) v[k] = s64[k];
    const bool inf = src[96] != 0;
    g1_packed_t p;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        unsigned long long a = inf ? 0ull : v[2 * k], b = inf ? 0ull : v[2 * k + 1];
        p.w[k] = make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b, (uint32_t)(b >> 32));
    }
    out[i] = p;
}

__global__ void msm_count_kernel(uint32_t* __restrict__ counts, const uint4* __restrict__ scalars, size_t n,
                                 MsmShape sh) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 a = scalars[2 * i], b = scalars[2 * i + 1];
    uint32_t s[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t carry = 0;
    for (uint32_t w = 0; w < sh.nwin; w++) {
        uint32_t neg;
        uint32_t d = msm_signed_digit(s, w, sh.c, carry, neg);
        if (d) atomicAdd(&counts[(size_t)w * sh.nbuckets + (d - 1)], 1u);
    }
}

__global__ void msm_scatter_kernel(uint32_t* __restrict__ entries, uint32_t* __restrict__ cursor,
                                   const uint4* __restrict__ scalars, size_t n, MsmShape sh) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 a = scalars[2 * i], b = scalars[2 * i + 1];
    uint32_t s[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t carry = 0;
    for (uint32_t w = 0; w < sh.nwin; w++) {
        uint32_t neg;
        uint32_t d = msm_signed_digit(s, w, sh.c, carry, neg);
        if (d) {
            uint32_t pos = atomicAdd(&cursor[(size_t)w * sh.nbuckets + (d - 1)], 1u);
            entries[pos] = (uint32_t)i | (neg << 31);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// exclusive scan of `count` u32 values (count + 1 outputs; out[count] = total), 3 small kernels
// ---------------------------------------------------------------------------------------------
#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_CHUNK (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= (uint32_t)d) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int d = 1; d < SCAN_THREADS / 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= (uint32_t)d) s += y;
        }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = s;      // inclusive warp totals
    }
    __syncthreads();
    uint32_t base = wid ? warp_sums[wid - 1] : 0;
    *total = warp_sums[SCAN_THREADS / 32 - 1];
    uint32_t r = base + x - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums_kernel(const uint32_t* __restrict__ in,
                                                                      uint32_t* __restrict__ block_sums,
                                                                      size_t count) {
    size_t base = (size_t)blockIdx.x * SCAN_CHUNK + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
        if (base + k < count) s += in[base + k];
    uint32_t total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}
__global__ void scan_sums_kernel(uint32_t* block_sums, uint32_t nblocks) {
    // nblocks <= a few thousand: one warp, chunked
    uint32_t carry = 0;
    for (uint32_t base = 0; base < nblocks; base += 32) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < nblocks ? block_sums[i] : 0;
        uint32_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
            if (threadIdx.x >= (uint32_t)d) x += y;
        }
        if (i < nblocks) block_sums[i] = carry + x - v;
        carry += __shfl_sync(0xffffffffu, x, 31);
    }
}
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const uint32_t* __restrict__ in,
                                                                  const uint32_t* __restrict__ block_sums,
                                                                  uint32_t* __restrict__ out, size_t count) {
    size_t base = (size_t)blockIdx.x * SCAN_CHUNK + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = (base + k < count) ? in[base + k] : 0;
        s += v[k];
    }
    uint32_t total;
    uint32_t off = block_exclusive_scan(s, &total) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k <= count) out[base + k] = off;     // note <=: writes out[count] = grand total
        off += v[k];
    }
}

// ---------------------------------------------------------------------------------------------
// bucket accumulation, load balanced.  A bucket with more than `task_len` entries is cut into
// ceil(cnt / task_len) equal tasks; one thread per task sums its slice of the sorted point list in XYZZ
// (next point fetched while the current mixed addition runs: the add is ~3.3k instructions, a gather from
// HBM a few hundred cycles).  Single-task buckets are written straight to `buckets`; the partial sums of
// multi-task buckets are combined by log2(max tasks) pairwise rounds.  This bounds the work of any thread
// by task_len additions whatever the scalar distribution (top window of the signed-digit split, repeated
// scalars, tiny scalars ...), where one thread per bucket degenerates to a serial loop over n points.
// ---------------------------------------------------------------------------------------------
__global__ void msm_task_count_kernel(uint32_t* __restrict__ ntask, const uint32_t* __restrict__ offsets,
                                      uint32_t nbuckets_total, uint32_t task_len) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > nbuckets_total) return;
    uint32_t cnt = k < nbuckets_total ? offsets[k + 1] - offsets[k] : 0;
    ntask[k] = (cnt + task_len - 1) / task_len;
}

__global__ void __launch_bounds__(MSM_ACC_THREADS) msm_accumulate_kernel(g1_xyzz_mem_t* __restrict__ buckets,
                                                                        g1_xyzz_mem_t* __restrict__ partials,
                                                                        uint32_t* __restrict__ task_bucket,
                                                                        const g1_packed_t* __restrict__ pts,
                                                                        const uint32_t* __restrict__ entries,
                                                                        const uint32_t* __restrict__ offsets,
                                                                        const uint32_t* __restrict__ task_off,
                                                                        uint32_t nbuckets_total) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= task_off[nbuckets_total]) return;
    // bucket of task t: the k with task_off[k] <= t < task_off[k + 1]
    uint32_t lo = 0, hi = nbuckets_total;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (task_off[mid] <= t) lo = mid; else hi = mid;
    }
    const uint32_t k = lo;
    const uint32_t nt = task_off[k + 1] - task_off[k], j = t - task_off[k];
    const uint32_t start = offsets[k], cnt = offsets[k + 1] - start;
    uint32_t e = start + (uint32_t)(((unsigned long long)cnt * j) / nt);
    const uint32_t end = start + (uint32_t)(((unsigned long long)cnt * (j + 1)) / nt);
    g1_xyzz_t acc = g1_xyzz_infinity();
    if (e < end) {
        uint32_t cur_id = entries[e];
        g1_packed_t cur = pts[cur_id & 0x7fffffffu];
        for (;;) {
            ++e;
            uint32_t nxt_id = 0;
            g1_packed_t nxt;
            const bool more = e < end;
            if (more) {
                nxt_id = entries[e];
                nxt = pts[nxt_id & 0x7fffffffu];
            }
            g1_affine_t a = g1_unpack(cur);
            if (cur_id >> 31) a.y = fp_neg(a.y);
            g1_madd(acc, a);
            if (!more) break;
            cur = nxt;
            cur_id = nxt_id;
        }
    }
    task_bucket[t] = k;
    g1_xyzz_store(nt == 1 ? buckets + k : partials + t, acc);
}

// one pairwise round over the partial sums of multi-task buckets: P[j] += P[j + stride] for j = 0 mod 2*stride
__global__ void __launch_bounds__(MSM_ACC_THREADS) msm_combine_round_kernel(g1_xyzz_mem_t* __restrict__ partials,
                                                                           const uint32_t* __restrict__ task_bucket,
                                                                           const uint32_t* __restrict__ task_off,
                                                                           uint32_t nbuckets_total, uint32_t stride) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= task_off[nbuckets_total]) return;
    const uint32_t k = task_bucket[t];
    const uint32_t base = task_off[k], nt = task_off[k + 1] - base;
    if (nt == 1) return;
    const uint32_t j = t - base;
    if ((j & (2 * stride - 1)) != 0 || j + stride >= nt) return;
    g1_xyzz_t a = g1_xyzz_load(partials + t);
    g1_xyzz_t b = g1_xyzz_load(partials + t + stride);
    g1_add(a, b);
    g1_xyzz_store(partials + t, a);
}

__global__ void msm_combine_final_kernel(g1_xyzz_mem_t* __restrict__ buckets, const g1_xyzz_mem_t* __restrict__ partials,
                                         const uint32_t* __restrict__ task_off, uint32_t nbuckets_total) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nbuckets_total) return;
    const uint32_t base = task_off[k];
    if (task_off[k + 1] - base > 1) buckets[k] = partials[base];
}

// ---------------------------------------------------------------------------------------------
// bucket reduction
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MSM_RED_THREADS) msm_reduce_segments_kernel(g1_xyzz_mem_t* __restrict__ segs,
                                                                             const g1_xyzz_mem_t* __restrict__ buckets,
                                                                             MsmShape sh, uint32_t seg_len,
                                                                             uint32_t segs_per_win) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= segs_per_win * sh.nwin) return;
    uint32_t w = t / segs_per_win, s = t % segs_per_win;
    uint32_t s0 = s * seg_len;
    uint32_t len = sh.nbuckets - s0 < seg_len ? sh.nbuckets - s0 : seg_len;
    g1_xyzz_t r = msm_reduce_segment(buckets + (size_t)w * sh.nbuckets, s0, len);
    g1_xyzz_store(segs + t, r);
}

// one block per window: sum of its segment results
__global__ void __launch_bounds__(MSM_TREE_THREADS) msm_window_sum_kernel(g1_xyzz_mem_t* __restrict__ wsum,
                                                                         const g1_xyzz_mem_t* __restrict__ segs,
                                                                         uint32_t segs_per_win) {
    __shared__ g1_xyzz_mem_t sh[MSM_TREE_THREADS];
    const uint32_t w = blockIdx.x, tid = threadIdx.x;
    g1_xyzz_t acc = g1_xyzz_infinity();
    for (uint32_t s = tid; s < segs_per_win; s += MSM_TREE_THREADS) {
        g1_xyzz_t v = g1_xyzz_load(segs + (size_t)w * segs_per_win + s);
        g1_add(acc, v);
    }
    g1_xyzz_store(&sh[tid], acc);
    __syncthreads();
    for (uint32_t d = MSM_TREE_THREADS / 2; d >= 1; d >>= 1) {
        if (tid < d) {
            g1_xyzz_t a = g1_xyzz_load(&sh[tid]);
            g1_xyzz_t b = g1_xyzz_load(&sh[tid + d]);
            g1_add(a, b);
            g1_xyzz_store(&sh[tid], a);
        }
        __syncthreads();
    }
    if (tid == 0) wsum[w] = sh[0];
}

__device__ __forceinline__ void store_jacobian(uint4* out, const g1_xyzz_t& p) {
    fq_t X, Y, Z;
    g1_xyzz_to_jacobian(p, X, Y, Z);
    fq_to_u4x3(X, out + 0);
    fq_to_u4x3(Y, out + 3);
    fq_to_u4x3(Z, out + 6);
}

__global__ void msm_fold_kernel(uint4* __restrict__ out_jac, const g1_xyzz_mem_t* __restrict__ wsum, MsmShape sh) {
    if (blockIdx.x || threadIdx.x) return;
    g1_xyzz_t total = msm_fold_windows(wsum, sh.nwin, sh.c);
    store_jacobian(out_jac, total);
}

__global__ void msm_write_infinity_kernel(uint4* out_jac) {
    if (blockIdx.x || threadIdx.x) return;
    store_jacobian(out_jac, g1_xyzz_infinity());
}

// sum of `count` Jacobian points (multi-GPU partial sums)
__global__ void g1_sum_jacobian_kernel(uint4* __restrict__ out_jac, const uint4* __restrict__ in, uint32_t count) {
    if (blockIdx.x || threadIdx.x) return;
    g1_xyzz_t total = g1_xyzz_infinity();
    for (uint32_t i = 0; i < count; i++) {
        const uint4* p = in + 9 * (size_t)i;
        fq_t X = fq_from_u4x3(p), Y = fq_from_u4x3(p + 3), Z = fq_from_u4x3(p + 6);
        g1_xyzz_t q;
        q.X = X;
        q.Y = Y;
        q.ZZ = fp_sqr(Z);
        q.ZZZ = fp_mul(q.ZZ, Z);
        g1_add(total, q);
    }
    store_jacobian(out_jac, total);
}

// ---------------------------------------------------------------------------------------------
// launcher
// ---------------------------------------------------------------------------------------------
extern "C" uint32_t b200_msm_window_bits(size_t n) {
    if (const char* e = getenv("B200_MSM_C")) {
        int c = atoi(e);
        if (c >= 2 && c <= 22) return (uint32_t)c;
    }
    uint32_t lg = 0;
    while (((size_t)1 << (lg + 1)) <= n) lg++;
    int c = (int)lg - 4;
    if (c < 4) c = 4;
    if (c > 21) c = 21;
    return (uint32_t)c;
}

b200_error_t msm_pack_bases_device(void* d_packed, const void* d_points, size_t n, size_t stride,
                                   cudaStream_t stream) {
    if (n == 0) return b200_ok();
    if (stride < 97 || (stride & 7)) return b200_err(B200_ERR_INVALID_ARG, "msm: affine stride must be >= 104 and 8-byte aligned");
    if (reinterpret_cast<uintptr_t>(d_points) & 7) return b200_err(B200_ERR_INVALID_ARG, "msm: points must be 8-byte aligned");
    msm_pack_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(reinterpret_cast<g1_packed_t*>(d_packed),
                                                                      reinterpret_cast<const uint8_t*>(d_points), n, stride);
    KERNEL_CHECK();
    return b200_ok();
}

static b200_error_t exclusive_scan(uint32_t* d_out, const uint32_t* d_in, size_t count, cudaStream_t stream) {
    // scans count + 1 slots so that out[count] = total
    size_t nblocks = (count + 1 + SCAN_CHUNK - 1) / SCAN_CHUNK;
    DevBuf sums;
    CUDA_TRY(sums.alloc(nblocks * 4, stream));
    scan_block_sums_kernel<<<(unsigned)nblocks, SCAN_THREADS, 0, stream>>>(d_in, sums.as<uint32_t>(), count);
    KERNEL_CHECK();
    scan_sums_kernel<<<1, 32, 0, stream>>>(sums.as<uint32_t>(), (uint32_t)nblocks);
    KERNEL_CHECK();
    scan_apply_kernel<<<(unsigned)nblocks, SCAN_THREADS, 0, stream>>>(d_in, sums.as<uint32_t>(), d_out, count);
    KERNEL_CHECK();
    return b200_ok();
}

b200_error_t msm_run_device(void* d_out, const void* d_points, size_t n, const void* d_scalars, size_t stride,
                            const void* d_packed, cudaStream_t stream) {
    if (!d_out) return b200_err(B200_ERR_INVALID_ARG, "msm: null output pointer");
    if (n == 0) {
        msm_write_infinity_kernel<<<1, 1, 0, stream>>>(reinterpret_cast<uint4*>(d_out));
        KERNEL_CHECK();
        return b200_ok();
    }
    if (!d_scalars || (!d_points && !d_packed)) return b200_err(B200_ERR_INVALID_ARG, "msm: null input pointer");
    if (n >= ((size_t)1 << 28)) return b200_err(B200_ERR_TOO_LARGE, "msm: more than 2^28 - 1 points per call");
    if (reinterpret_cast<uintptr_t>(d_scalars) & 15) return b200_err(B200_ERR_INVALID_ARG, "msm: scalars must be 16-byte aligned on the device");

    const MsmShape sh = msm_shape(b200_msm_window_bits(n));
    const size_t K = (size_t)sh.nwin * sh.nbuckets;
    if ((size_t)n * sh.nwin >= ((size_t)1 << 32)) return b200_err(B200_ERR_TOO_LARGE, "msm: n * windows overflows 32-bit offsets");

    DevBuf packed, counts, offsets, cursor, entries, buckets, segs, wsum, ntask, task_off, task_bucket, partials;
    const g1_packed_t* pts = reinterpret_cast<const g1_packed_t*>(d_packed);
    if (!pts) {
        STAGE("msm_pack", stream);
        CUDA_TRY(packed.alloc(n * sizeof(g1_packed_t), stream));
        B200_TRY(msm_pack_bases_device(packed.p, d_points, n, stride, stream));
        pts = packed.as<g1_packed_t>();
    }
    CUDA_TRY(counts.alloc((K + 1) * 4, stream));
    CUDA_TRY(offsets.alloc((K + 1) * 4, stream));
    CUDA_TRY(cursor.alloc((K + 1) * 4, stream));
    CUDA_TRY(entries.alloc(n * sh.nwin * 4, stream));
    CUDA_TRY(buckets.alloc(K * sizeof(g1_xyzz_mem_t), stream));
    const uint32_t seg_len = sh.nbuckets < MSM_SEG_LEN ? sh.nbuckets : MSM_SEG_LEN;
    const uint32_t segs_per_win = (sh.nbuckets + seg_len - 1) / seg_len;
    CUDA_TRY(segs.alloc((size_t)segs_per_win * sh.nwin * sizeof(g1_xyzz_mem_t), stream));
    CUDA_TRY(wsum.alloc((size_t)sh.nwin * sizeof(g1_xyzz_mem_t), stream));

    STAGE("msm_count", stream);
    CUDA_TRY(cudaMemsetAsync(counts.p, 0, (K + 1) * 4, stream));
    const unsigned nblk = (unsigned)((n + 255) / 256);
    msm_count_kernel<<<nblk, 256, 0, stream>>>(counts.as<uint32_t>(), reinterpret_cast<const uint4*>(d_scalars), n, sh);
    KERNEL_CHECK();
    STAGE("msm_scan", stream);
    B200_TRY(exclusive_scan(offsets.as<uint32_t>(), counts.as<uint32_t>(), K, stream));
    CUDA_TRY(cudaMemcpyAsync(cursor.p, offsets.p, (K + 1) * 4, cudaMemcpyDeviceToDevice, stream));
    STAGE("msm_scatter", stream);
    msm_scatter_kernel<<<nblk, 256, 0, stream>>>(entries.as<uint32_t>(), cursor.as<uint32_t>(),
                                                 reinterpret_cast<const uint4*>(d_scalars), n, sh);
    KERNEL_CHECK();
    // ---- load-balanced accumulation ----
    STAGE("msm_tasks", stream);
    // task length: twice the mean bucket load, so that ordinary buckets are one task and only heavy ones split
    size_t mean = n / sh.nbuckets;
    uint32_t task_len = (uint32_t)(2 * mean);
    if (const char* e = getenv("B200_MSM_TASK_LEN")) task_len = (uint32_t)atoi(e);
    if (task_len < 32) task_len = 32;
    // upper bound on the number of tasks: sum_k ceil(cnt_k / L) <= E / L + (non-empty buckets)
    const size_t E = n * sh.nwin;
    const size_t t_max = E / task_len + (K < E ? K : E) + 1;
    CUDA_TRY(ntask.alloc((K + 1) * 4, stream));
    CUDA_TRY(task_off.alloc((K + 2) * 4, stream));
    CUDA_TRY(task_bucket.alloc(t_max * 4, stream));
    CUDA_TRY(partials.alloc(t_max * sizeof(g1_xyzz_mem_t), stream));
    msm_task_count_kernel<<<(unsigned)((K + 1 + 255) / 256), 256, 0, stream>>>(ntask.as<uint32_t>(), offsets.as<uint32_t>(),
                                                                              (uint32_t)K, task_len);
    KERNEL_CHECK();
    B200_TRY(exclusive_scan(task_off.as<uint32_t>(), ntask.as<uint32_t>(), K, stream));
    CUDA_TRY(cudaMemsetAsync(buckets.p, 0, K * sizeof(g1_xyzz_mem_t), stream));      // ZZ = 0: empty buckets are infinity
    STAGE("msm_accumulate", stream);
    const unsigned tblocks = (unsigned)((t_max + MSM_ACC_THREADS - 1) / MSM_ACC_THREADS);
    msm_accumulate_kernel<<<tblocks, MSM_ACC_THREADS, 0, stream>>>(
        buckets.as<g1_xyzz_mem_t>(), partials.as<g1_xyzz_mem_t>(), task_bucket.as<uint32_t>(), pts, entries.as<uint32_t>(),
        offsets.as<uint32_t>(), task_off.as<uint32_t>(), (uint32_t)K);
    KERNEL_CHECK();
    STAGE("msm_combine", stream);
    const size_t max_tasks_per_bucket = (n + task_len - 1) / task_len;
    for (uint32_t stride = 1; stride < max_tasks_per_bucket; stride <<= 1) {
        msm_combine_round_kernel<<<tblocks, MSM_ACC_THREADS, 0, stream>>>(partials.as<g1_xyzz_mem_t>(), task_bucket.as<uint32_t>(),
                                                                         task_off.as<uint32_t>(), (uint32_t)K, stride);
        KERNEL_CHECK();
    }
    msm_combine_final_kernel<<<(unsigned)((K + 255) / 256), 256, 0, stream>>>(buckets.as<g1_xyzz_mem_t>(), partials.as<g1_xyzz_mem_t>(),
                                                                             task_off.as<uint32_t>(), (uint32_t)K);
    KERNEL_CHECK();
    STAGE("msm_reduce_segments", stream);
    const uint32_t nseg_threads = segs_per_win * sh.nwin;
    msm_reduce_segments_kernel<<<(nseg_threads + MSM_RED_THREADS - 1) / MSM_RED_THREADS, MSM_RED_THREADS, 0, stream>>>(
        segs.as<g1_xyzz_mem_t>(), buckets.as<g1_xyzz_mem_t>(), sh, seg_len, segs_per_win);
    KERNEL_CHECK();
    STAGE("msm_window_sum", stream);
    msm_window_sum_kernel<<<sh.nwin, MSM_TREE_THREADS, 0, stream>>>(wsum.as<g1_xyzz_mem_t>(), segs.as<g1_xyzz_mem_t>(),
                                                                   segs_per_win);
    KERNEL_CHECK();
    STAGE("msm_fold", stream);
    msm_fold_kernel<<<1, 32, 0, stream>>>(reinterpret_cast<uint4*>(d_out), wsum.as<g1_xyzz_mem_t>(), sh);
    KERNEL_CHECK();
    STAGE_END(stream);
    return b200_ok();
}

extern "C" b200_error_t b200_g1_sum_jacobian_device(void* d_out, const void* d_in, size_t count, void* stream) {
    B200_TRY(b200_require_device());
    if (!d_out || (count && !d_in)) return b200_err(B200_ERR_INVALID_ARG, "g1_sum: null pointer");
    g1_sum_jacobian_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint4*>(d_out),
                                                              reinterpret_cast<const uint4*>(d_in), (uint32_t)count);
    KERNEL_CHECK();
    return b200_ok();
}
