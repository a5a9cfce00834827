// msm.cu -- kernels + launcher for VariableBase::msm over BLS12-377 G1 on B200.
//
// Pipeline (all on the device, one stream):
//   pack      : G1Affine images (stride 104) -> 128-byte records (x | y | pad), infinity -> (0, 0)
//   count     : signed c-bit digits of every scalar; histogram of (window, bucket)        [atomics in L2]
//   scan      : exclusive prefix sum of the histogram -> start offset of every bucket
//   scatter   : point index (+ sign bit) of every non-zero digit into its bucket's slot      [counting sort]
//   accumulate: the sorted entry stream cut into equal chunks, one thread per chunk, XYZZ mixed additions; runs that
//               cross a chunk boundary are stitched by pairwise combine rounds
//   reduce    : per window, segmented running sum  sum_b (b + 1) * B_b ; block tree per window
//   fold      : Horner over the windows with c doublings each -> Jacobian result
// HBM layout: packed bases n * 128 B | entries n * nwin * 4 B | offsets (nwin * 2^(c-1) + 1) * 4 B |
//             buckets nwin * 2^(c-1) * 192 B (XYZZ) | segment sums | window sums.
#include <algorithm>
#include <cstdlib>
#include <memory>

#include <cooperative_groups.h>

#include "common.cuh"
#include "msm_core.cuh"
#include "msm_affine.cuh"
#include "msm_glv.cuh"
#include "ec_coop.cuh"

#define MSM_ACC_THREADS 128
#define MSM_RED_THREADS 64
#define MSM_SEG_LEN 32u
#define MSM_TREE_THREADS 128

// ---------------------------------------------------------------------------------------------
// pack / count / scatter
// ---------------------------------------------------------------------------------------------
__global__ void msm_pack_kernel(uint4* __restrict__ out, const uint8_t* __restrict__ pts, size_t n,
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
    g1_store_packed(out + i * G1_BASE_U4, p);
}

// GLV form of the inputs (msm_glv.cuh): point i becomes records 2i = P_i and 2i + 1 = phi(P_i) = (beta x, y), scalar i
// becomes the 128-bit halves 2i = k mod lambda and 2i + 1 = k div lambda.
__global__ void msm_pack_glv_kernel(uint4* __restrict__ out, const uint8_t* __restrict__ pts, size_t n, size_t stride) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* src = pts + i * stride;
    const unsigned long long* s64 = reinterpret_cast<const unsigned long long*>(src);
    const bool inf = src[96] != 0;
    g1_packed_t p;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        unsigned long long a = inf ? 0ull : s64[2 * k], b = inf ? 0ull : s64[2 * k + 1];
        p.w[k] = make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b, (uint32_t)(b >> 32));
    }
    g1_store_packed(out + 2 * i * G1_BASE_U4, p);
    g1_affine_t a = g1_unpack(p);
    a.x = fp_mul(a.x, msm_glv_beta());                     // infinity (0, 0) stays (0, 0)
    fq_to_u4x3(a.x, p.w);
    g1_store_packed(out + (2 * i + 1) * G1_BASE_U4, p);
}
__global__ void msm_glv_split_kernel(uint4* __restrict__ half, const uint4* __restrict__ scalars, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 a = scalars[2 * i], b = scalars[2 * i + 1];
    uint32_t k[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t k1[4], k2[4];
    msm_scalar_reduce(k);
    msm_glv_split(k, k1, k2);
    half[2 * i] = make_uint4(k1[0], k1[1], k1[2], k1[3]);
    half[2 * i + 1] = make_uint4(k2[0], k2[1], k2[2], k2[3]);
}

// Warp-aggregated histogram update.  Structural hot spots (the top window of the signed-digit split has only a
// few possible digits; repeated or tiny scalars) otherwise serialise millions of atomics on one address (measured:
// 8 ms instead of 3 ms at 2^24, c = 18).  __match_any_sync would find every group but costs ~500 cycles when the
// 32 keys are distinct (the common case; measured +3 ms), so only two leader rounds are run: the key of the lowest
// pending lane is broadcast, every lane holding it joins that leader's single atomicAdd and receives its rank; lanes
// still pending after two rounds (all-distinct keys) fall back to one atomic each.  Returns the reserved slot.
__device__ __forceinline__ uint32_t warp_aggregated_inc(uint32_t* counters, size_t key, bool active) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t pending = __ballot_sync(0xffffffffu, active);
    uint32_t pos = 0;
    bool mine = active;
#pragma unroll
    for (int round = 0; round < 2; round++) {
        if (pending == 0) break;                                   // warp-uniform
        const uint32_t leader = __ffs(pending) - 1;
        const unsigned long long lkey = __shfl_sync(0xffffffffu, (unsigned long long)key, leader);
        const uint32_t group = __ballot_sync(0xffffffffu, mine && (unsigned long long)key == lkey);
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(&counters[key], (uint32_t)__popc(group));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (mine && ((group >> lane) & 1u)) {
            pos = base + __popc(group & ((1u << lane) - 1));
            mine = false;
        }
        pending &= ~group;
    }
    if (mine) pos = atomicAdd(&counters[key], 1u);
    return pos;
}

// Batched calls (many independent small MSMs in one launch set, e.g. the verifier's linear combinations of a whole
// block of transactions): point i belongs to MSM m with seg_off[m] <= i < seg_off[m + 1]; its buckets live in the
// "virtual windows" m * nwin + w, so everything after the digit kernels is unchanged.
__device__ __forceinline__ uint32_t msm_segment_of(const unsigned long long* seg_off, uint32_t nmsm, size_t i) {
    if (!seg_off) return 0;
    uint32_t lo = 0, hi = nmsm;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (seg_off[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// scalar i as 8 limbs: a 256-bit canonical scalar (two uint4) or, after the GLV split, one 128-bit half
__device__ __forceinline__ void msm_load_scalar(uint32_t* s, const uint4* __restrict__ scalars, size_t i, uint32_t glv) {
    if (glv) {
        const uint4 a = scalars[i];
        s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w;
    } else {
        const uint4 a = scalars[2 * i], b = scalars[2 * i + 1];
        s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w; s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
        msm_scalar_reduce(s);
    }
}

// Tabulated bases (n_reg != 0): the caller holds 2^(c*w) * P_i for every window w at index w * n_reg + i, so all
// windows share ONE set of 2^(c-1) buckets -- the key is the bucket alone and the entry points into the table.
__global__ void msm_count_kernel(uint32_t* __restrict__ counts, const uint4* __restrict__ scalars, size_t n,
                                 MsmShape sh, const unsigned long long* __restrict__ seg_off, uint32_t nmsm,
                                 size_t n_reg, uint32_t glv) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n;
    // bucket set of MSM m: its own nwin windows, or ONE set when the bases are tabulated (all windows share it)
    const size_t vbase = valid ? (size_t)msm_segment_of(seg_off, nmsm, glv ? i >> 1 : i) * (n_reg ? 1 : sh.nwin) : 0;
    uint32_t s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (valid) msm_load_scalar(s, scalars, i, glv);
    uint32_t carry = 0;
    for (uint32_t w = 0; w < sh.nwin; w++) {
        uint32_t neg;
        uint32_t d = msm_signed_digit(s, w, sh.c, carry, neg);
        warp_aggregated_inc(counts, (n_reg ? vbase : vbase + w) * sh.nbuckets + (d ? d - 1 : 0), valid && d != 0);
    }
}

// scatter is launched window-major (blockIdx.y = window): at any moment the chip works on one or two windows, so
// the live cursor slice and that window's slice of `entries` (n * 4 B) stay in L2 and the 4-byte scattered writes
// merge there instead of each dirtying a 32-byte HBM sector (measured 8.9 -> 3.7 ms at 2^24, c = 20).  Every
// thread re-derives the carry chain of the lower windows (a few shifts per window).
__global__ void msm_scatter_kernel(uint32_t* __restrict__ entries, uint32_t* __restrict__ cursor,
                                   const uint4* __restrict__ scalars, size_t n, MsmShape sh,
                                   const unsigned long long* __restrict__ seg_off, uint32_t nmsm, size_t n_reg,
                                   uint32_t glv, uint32_t shared_bases) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n;
    const uint32_t w = blockIdx.y;
    const uint32_t seg = valid ? msm_segment_of(seg_off, nmsm, glv ? i >> 1 : i) : 0;
    const size_t vbase = (size_t)seg * (n_reg ? 1 : sh.nwin);
    // shared bases (k polynomials committed against ONE resident set): scalar i of MSM m multiplies base i - seg_off[m]
    size_t pi = i;
    if (valid && shared_bases && seg_off) pi = glv ? ((i >> 1) - (size_t)seg_off[seg]) * 2 + (i & 1) : i - (size_t)seg_off[seg];
    uint32_t s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (valid) msm_load_scalar(s, scalars, i, glv);
    uint32_t carry = 0, neg = 0, d = 0;
    for (uint32_t ww = 0; ww <= w; ww++) d = msm_signed_digit(s, ww, sh.c, carry, neg);
    const bool active = valid && d != 0;
    uint32_t pos = warp_aggregated_inc(cursor, (n_reg ? vbase : vbase + w) * sh.nbuckets + (d ? d - 1 : 0), active);
    if (active) entries[pos] = (uint32_t)(n_reg ? (size_t)w * n_reg + pi : pi) | (neg << 31);
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
// bucket accumulation, equal work per lane.  The sorted entry stream is cut into chunks of `chunk` entries
// regardless of bucket boundaries; one thread per chunk walks it with XYZZ mixed additions (the next point is
// fetched while the current addition runs: the add is ~3.3k instructions, a gather from HBM a few hundred
// cycles).  Every lane of a warp does the same number of additions, whatever the bucket-size distribution, so
// the window width can grow (fewer windows = fewer additions) without losing SIMT efficiency, and no scalar
// distribution (top window of the signed-digit split, repeated or tiny scalars) can pile n points on a thread.
// A run of equal-bucket entries inside a chunk is flushed when the bucket ends:
//   bucket entirely inside the chunk          -> buckets[k]
//   run that continues a bucket begun earlier -> heads[t]   (first run of thread t)
//   run that starts a bucket continuing later -> tails[t]   (last run of thread t)
// and a bucket spanning threads t_a .. t_b is tails[t_a] + heads[t_a+1] + ... + heads[t_b]: the heads are summed
// by pairwise rounds (log2 of the longest span), the tail is added last.
// ---------------------------------------------------------------------------------------------
#define MSM_NONE 0xffffffffu

// DIRECT: the entry stream IS the point list (output of the batched-affine rounds), no index / sign indirection.
// ADD: `buckets` holds the sums of earlier point ranges (streamed host path) and this range is added to them: a bucket's
// walk starts from its stored sum instead of infinity -- the merge costs no product at all (a separate pass over the
// bucket arrays is one general XYZZ addition, 14 products, per bucket and range); buckets without entries stay as they are.
template <bool DIRECT, bool ADD>
__global__ void __launch_bounds__(MSM_ACC_THREADS, 3) msm_accumulate_kernel(g1_xyzz_mem_t* __restrict__ buckets,
                                                                        g1_xyzz_mem_t* __restrict__ heads,
                                                                        g1_xyzz_mem_t* __restrict__ tails,
                                                                        uint32_t* __restrict__ head_bucket,
                                                                        uint32_t* __restrict__ tail_bucket,
                                                                        const uint4* __restrict__ pts,
                                                                        const uint4* __restrict__ pts_y,
                                                                        const uint32_t* __restrict__ entries,
                                                                        const uint32_t* __restrict__ offsets,
                                                                        uint32_t* __restrict__ max_heads,
                                                                        uint32_t nbuckets_total, uint32_t chunk) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = offsets[nbuckets_total];
    const unsigned long long e0 = (unsigned long long)t * chunk;
    if (e0 >= total) return;
    uint32_t e = (uint32_t)e0;
    const uint32_t end = (e0 + chunk < total) ? (uint32_t)(e0 + chunk) : total;
    // bucket of the first entry: the k with offsets[k] <= e < offsets[k + 1]
    uint32_t lo = 0, hi = nbuckets_total;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= e) lo = mid; else hi = mid;
    }
    uint32_t k = lo;
    uint32_t bucket_end = offsets[k + 1];
    bool from_start = (offsets[k] == e);
    g1_xyzz_t acc = g1_xyzz_infinity();
    if (ADD && from_start) acc = g1_xyzz_load(buckets + k);
    uint32_t cur_id = DIRECT ? e : entries[e];
    g1_packed_t cur = DIRECT ? g1_load_planes(pts + 3 * (size_t)cur_id, pts_y + 3 * (size_t)cur_id) : g1_load_packed(pts + (size_t)(cur_id & 0x7fffffffu) * G1_BASE_U4);
    for (;;) {
        const uint32_t nxt_e = e + 1;
        uint32_t nxt_id = 0;
        g1_packed_t nxt;
        if (nxt_e < end) {
            nxt_id = DIRECT ? nxt_e : entries[nxt_e];
            nxt = DIRECT ? g1_load_planes(pts + 3 * (size_t)nxt_id, pts_y + 3 * (size_t)nxt_id) : g1_load_packed(pts + (size_t)(nxt_id & 0x7fffffffu) * G1_BASE_U4);
        }
        g1_affine_t a = g1_unpack(cur);
        if (!DIRECT && (cur_id >> 31)) a.y = fp_neg(a.y);
        g1_madd(acc, a);
        e = nxt_e;
        const bool at_bucket_end = (e == bucket_end), at_chunk_end = (e == end);
        if (at_bucket_end || at_chunk_end) {
            if (!from_start) {
                g1_xyzz_store(heads + t, acc);
                head_bucket[t] = k;
                // longest run of heads of any bucket (plain read as a filter, atomic only when it grows)
                const uint32_t nh = (bucket_end - 1) / chunk - offsets[k] / chunk;
                if (nh > *(volatile uint32_t*)max_heads) atomicMax(max_heads, nh);
            } else if (at_bucket_end) {
                g1_xyzz_store(buckets + k, acc);
            } else {
                g1_xyzz_store(tails + t, acc);
                tail_bucket[t] = k;
            }
            if (at_chunk_end) break;
            do { ++k; } while (offsets[k + 1] == e);          // next non-empty bucket starts exactly at e
            bucket_end = offsets[k + 1];
            from_start = true;
            if (ADD) acc = g1_xyzz_load(buckets + k);
            else acc = g1_xyzz_infinity();
        }
        cur = nxt;
        cur_id = nxt_id;
    }
}

// one pairwise round over the heads of every spanning bucket: heads[j] += heads[j + stride], j = 0 mod 2*stride
__device__ __forceinline__ void msm_combine_heads_step(g1_xyzz_mem_t* __restrict__ heads, const uint32_t* __restrict__ head_bucket,
                                                       const uint32_t* __restrict__ offsets, uint32_t t, uint32_t nthreads_total,
                                                       uint32_t chunk, uint32_t stride) {
    if (t >= nthreads_total) return;
    const uint32_t k = head_bucket[t];
    if (k == MSM_NONE) return;
    const uint32_t ta = offsets[k] / chunk, tb = (offsets[k + 1] - 1) / chunk;
    const uint32_t nh = tb - ta, j = t - (ta + 1);
    if ((j & (2 * stride - 1)) != 0 || j + stride >= nh) return;
    g1_xyzz_t a = g1_xyzz_load(heads + t);
    g1_xyzz_t b = g1_xyzz_load(heads + t + stride);
    g1_add(a, b);
    g1_xyzz_store(heads + t, a);
}
// bucket = tail of the thread where it starts + the (already summed) heads of the following threads
__device__ __forceinline__ void msm_combine_tails_step(g1_xyzz_mem_t* __restrict__ buckets, const g1_xyzz_mem_t* __restrict__ heads,
                                                       const g1_xyzz_mem_t* __restrict__ tails, const uint32_t* __restrict__ tail_bucket,
                                                       uint32_t t, uint32_t nthreads_total) {
    if (t >= nthreads_total) return;
    const uint32_t k = tail_bucket[t];
    if (k == MSM_NONE) return;
    g1_xyzz_t a = g1_xyzz_load(tails + t);
    g1_xyzz_t b = g1_xyzz_load(heads + t + 1);
    g1_add(a, b);
    g1_xyzz_store(buckets + k, a);
}

__global__ void __launch_bounds__(MSM_ACC_THREADS) msm_combine_heads_kernel(g1_xyzz_mem_t* __restrict__ heads,
                                                                           const uint32_t* __restrict__ head_bucket,
                                                                           const uint32_t* __restrict__ offsets,
                                                                           const uint32_t* __restrict__ max_heads,
                                                                           uint32_t nthreads_total, uint32_t chunk,
                                                                           uint32_t stride) {
    if (stride >= *max_heads) return;                          // no bucket has that many heads: whole round is a no-op
    msm_combine_heads_step(heads, head_bucket, offsets, blockIdx.x * blockDim.x + threadIdx.x, nthreads_total, chunk, stride);
}

__global__ void __launch_bounds__(MSM_ACC_THREADS) msm_combine_tails_kernel(g1_xyzz_mem_t* __restrict__ buckets,
                                                                           const g1_xyzz_mem_t* __restrict__ heads,
                                                                           const g1_xyzz_mem_t* __restrict__ tails,
                                                                           const uint32_t* __restrict__ tail_bucket,
                                                                           uint32_t nthreads_total) {
    msm_combine_tails_step(buckets, heads, tails, tail_bucket, blockIdx.x * blockDim.x + threadIdx.x, nthreads_total);
}

// All rounds and the tails in ONE cooperative launch, for grids that are resident as a whole (small calls): the number
// of rounds is known on the device only (max_heads), so the launcher above issues log2(longest possible span) launches
// of which all but the first one or two return at once -- 16 launches, 0.12 ms of a 1.2 ms commit at 2^16.
__global__ void __launch_bounds__(MSM_ACC_THREADS) msm_combine_coop_kernel(g1_xyzz_mem_t* __restrict__ buckets,
                                                                          g1_xyzz_mem_t* __restrict__ heads,
                                                                          const g1_xyzz_mem_t* __restrict__ tails,
                                                                          const uint32_t* __restrict__ head_bucket,
                                                                          const uint32_t* __restrict__ tail_bucket,
                                                                          const uint32_t* __restrict__ offsets,
                                                                          const uint32_t* __restrict__ max_heads,
                                                                          uint32_t nthreads_total, uint32_t chunk) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t mh = *max_heads;                            // final: the accumulation kernel has completed
    for (uint32_t stride = 1; stride < mh; stride <<= 1) {
        msm_combine_heads_step(heads, head_bucket, offsets, t, nthreads_total, chunk, stride);
        grid.sync();
    }
    msm_combine_tails_step(buckets, heads, tails, tail_bucket, t, nthreads_total);
}

// ---------------------------------------------------------------------------------------------
// batched-affine pair rounds (msm_affine.cuh): one thread per MSM_PAIRS_PER_THREAD outputs
// ---------------------------------------------------------------------------------------------
#define MSM_PAIR_THREADS 128

__global__ void msm_half_counts_kernel(uint32_t* __restrict__ cnt, const uint32_t* __restrict__ off, uint32_t K) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < K) cnt[k] = (off[k + 1] - off[k] + 1) >> 1;
}

// both kernels work on the thread range [t0, t1) of the round: a round is issued in slices so that the additions of
// one slice overlap the denominators of the next
__global__ void __launch_bounds__(MSM_PAIR_THREADS) msm_pair_denoms_kernel(PairRound rd, uint4* __restrict__ pre,
                                                                          uint4* __restrict__ partial, uint32_t t0, uint32_t t1) {
    const uint32_t t = t0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (t < t1) pair_denoms_thread(rd, t, pre, partial);
}

// Four 128-thread blocks per SM at 128 registers (16 warps).  Measured alternatives: 5 blocks (96 registers, 124 bytes
// of spills, 20 warps) is SLOWER -- 23.56 / 21.83 ms against 22.75 / 21.00 ms for round 0 / rounds 1..4 at 2^24; 3 blocks
// change nothing (the kernel wants 132 registers).  The 19 % it is short of the multiplier peak are fixed-latency waits
// of the IMAD.WIDE carry chains with four warps per scheduler (ncu: `wait` 3.9, `long_scoreboard` 1.6 - 2.8).
#ifndef MSM_PAIR_ADD_MIN_BLOCKS
#define MSM_PAIR_ADD_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(MSM_PAIR_THREADS, MSM_PAIR_ADD_MIN_BLOCKS) msm_pair_add_kernel(PairRound rd, const uint4* __restrict__ pre,
                                                                       const uint4* __restrict__ partial_inv,
                                                                       uint4* __restrict__ out_x, uint4* __restrict__ out_y,
                                                                       uint32_t t0, uint32_t t1) {
    const uint32_t t = t0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (t < t1) pair_add_thread(rd, t, pre, partial_inv, out_x, out_y);
}

__global__ void __launch_bounds__(128) fq_inv_up_kernel(uint4* __restrict__ partial, const uint4* __restrict__ data, size_t n) {
    fq_inv_up_thread(partial, data, n, (size_t)blockIdx.x * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(128) fq_inv_down_kernel(uint4* __restrict__ data, const uint4* __restrict__ partial_inv, size_t n) {
    fq_inv_down_thread(data, partial_inv, n, (size_t)blockIdx.x * blockDim.x + threadIdx.x);
}
// Tail of the recursion: what counts is the latency of one inversion.  Binary extended Euclid (fp_inv_gcd: ALU pipe,
// ~550 iterations of ~130 instructions) instead of the Fermat ladder (~570 dependent Fq products, 0.52 ms per round
// measured); one value per thread, one warp per block, so that every warp has an SM sub-partition to itself.
#define FQ_INV_TAIL_MAX 16384u
__global__ void __launch_bounds__(32) fq_inv_tail_kernel(uint4* __restrict__ data, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) fq_to_u4x3(fp_inv_gcd(pair_load_fq(data + 3 * i)), data + 3 * i);
}

// in-place inversion of n Fq values, none of them zero
static b200_error_t fq_batch_inverse_nonzero(uint4* d_data, size_t n, cudaStream_t s) {
    if (n == 0) return b200_ok();
    if (n <= FQ_INV_TAIL_MAX) {
        fq_inv_tail_kernel<<<(unsigned)((n + 31) / 32), 32, 0, s>>>(d_data, n);
        KERNEL_CHECK();
        return b200_ok();
    }
    const size_t nt = (n + MSM_INV_CHUNK - 1) / MSM_INV_CHUNK;
    DevBuf partial;
    CUDA_TRY(partial.alloc(nt * 48, s));
    fq_inv_up_kernel<<<(unsigned)((nt + 127) / 128), 128, 0, s>>>(partial.as<uint4>(), d_data, n);
    KERNEL_CHECK();
    B200_TRY(fq_batch_inverse_nonzero(partial.as<uint4>(), nt, s));
    fq_inv_down_kernel<<<(unsigned)((nt + 127) / 128), 128, 0, s>>>(d_data, partial.as<uint4>(), n);
    KERNEL_CHECK();
    return b200_ok();
}

// ---------------------------------------------------------------------------------------------
// bucket reduction
// ---------------------------------------------------------------------------------------------
// (An out-of-line g1_add here -- 168 registers instead of 255 -- was measured slower: 14.0 vs 11.5 ms at 2^24, c = 20.)
// (A compact form -- the exceptional doubling out of line, all additions of the running-sum loop through one call site
// picked with selects -- was also built: 27 k instead of 43 k instructions, but still 255 registers and a 408-byte
// stack frame for the call; not pursued.)
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

// The same segment sums with a quad of lanes per segment (ec_coop.cuh), for calls with so few segments that the
// kernel is one dependent chain per thread on a mostly idle chip (a 2^16-point commit: 4096 segments, 16 additions and a
// ~24-step double-and-add each: 0.69 ms with one lane per segment).
__global__ void __launch_bounds__(MSM_RED_THREADS) msm_reduce_segments_quad_kernel(g1_xyzz_mem_t* __restrict__ segs,
                                                                                  const g1_xyzz_mem_t* __restrict__ buckets,
                                                                                  MsmShape sh, uint32_t seg_len,
                                                                                  uint32_t segs_per_win) {
    const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    if (t >= segs_per_win * sh.nwin) return;
    const Quad Q = quad_here();
    const uint32_t w = t / segs_per_win, s = t % segs_per_win;
    const uint32_t s0 = s * seg_len;
    const uint32_t len = sh.nbuckets - s0 < seg_len ? sh.nbuckets - s0 : seg_len;
    const g1_xyzz_mem_t* b = buckets + (size_t)w * sh.nbuckets + s0;
    g1_xyzz_t running = g1_xyzz_infinity(), acc = g1_xyzz_infinity();
    for (uint32_t i = len; i-- > 0;) {
        const g1_xyzz_t v = g1_xyzz_load(b + i);
        g1_add_quad(Q, running, v);
        g1_add_quad(Q, acc, running);
    }
    if (s0) {                                                  // + s0 * running, left-to-right double-and-add
        g1_xyzz_t m = g1_xyzz_infinity();
        bool started = false;
        for (int i = 31 - __clz(s0); i >= 0; i--) {
            if (started) g1_dbl_quad(Q, m);
            if ((s0 >> i) & 1u) {
                if (started) g1_add_quad(Q, m, running);
                else { m = running; started = true; }
            }
        }
        g1_add_quad(Q, acc, m);
    }
    if (Q.q == 0) g1_xyzz_store(segs + t, acc);
}

// Sum of the segment results of every window, in up to two levels: blockIdx.x = window, blockIdx.y = slice of that
// window's `count` inputs (each block adds its slice: strided per-thread sums, then a shared-memory tree) and writes
// out[window * gridDim.y + blockIdx.y].  The launcher runs it once with many slices per window and once more over the
// slice sums, so a call with few windows and millions of buckets (tabulated bases: ONE window of 2^23 buckets) does
// not serialise a hundred thousand additions on 128 threads (measured 13.9 ms -> 0.2 ms).
__global__ void __launch_bounds__(MSM_TREE_THREADS) msm_window_sum_kernel(g1_xyzz_mem_t* __restrict__ out,
                                                                         const g1_xyzz_mem_t* __restrict__ in,
                                                                         uint32_t count, uint32_t per_block) {
    __shared__ g1_xyzz_mem_t sh[MSM_TREE_THREADS];
    const uint32_t w = blockIdx.x, tid = threadIdx.x;
    const uint32_t lo = blockIdx.y * per_block;
    const uint32_t hi = lo + per_block < count ? lo + per_block : count;
    g1_xyzz_t acc = g1_xyzz_infinity();
    for (uint32_t s = lo + tid; s < hi; s += MSM_TREE_THREADS) {
        g1_xyzz_t v = g1_xyzz_load(in + (size_t)w * count + s);
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
    if (tid == 0) out[(size_t)w * gridDim.y + blockIdx.y] = sh[0];
}

// The same sums for calls with so few inputs that the tree is one dependent chain on a mostly idle chip (a 2^16-point
// commit: 4096 segment sums, 0.29 ms of its 1.29 ms in the two levels above -- ncu, profiles/r02_ncu_commit_2p16.txt):
// a QUAD of lanes per running sum (ec_coop.cuh: 4 product latencies per addition instead of 14), 32 quads per block,
// short slices (the launcher uses 128 inputs per block: 4 serial additions + 5 tree levels).
__global__ void __launch_bounds__(MSM_TREE_THREADS) msm_window_sum_quad_kernel(g1_xyzz_mem_t* __restrict__ out,
                                                                              const g1_xyzz_mem_t* __restrict__ in,
                                                                              uint32_t count, uint32_t per_block) {
    __shared__ g1_xyzz_mem_t sh[MSM_TREE_THREADS / 4];
    const Quad Q = quad_here();
    const uint32_t w = blockIdx.x, qid = threadIdx.x >> 2, nquads = MSM_TREE_THREADS / 4;
    const uint32_t lo = blockIdx.y * per_block;
    const uint32_t hi = lo + per_block < count ? lo + per_block : count;
    g1_xyzz_t acc = g1_xyzz_infinity();
    for (uint32_t s = lo + qid; s < hi; s += nquads) {           // trip count is quad-uniform
        const g1_xyzz_t v = g1_xyzz_load(in + (size_t)w * count + s);
        g1_add_quad(Q, acc, v);
    }
    if (Q.q == 0) g1_xyzz_store(&sh[qid], acc);
    __syncthreads();
    for (uint32_t d = nquads / 2; d >= 1; d >>= 1) {
        if (qid < d) {
            g1_xyzz_t a = g1_xyzz_load(&sh[qid]);
            const g1_xyzz_t b = g1_xyzz_load(&sh[qid + d]);
            g1_add_quad(Q, a, b);
            if (Q.q == 0) g1_xyzz_store(&sh[qid], a);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[(size_t)w * gridDim.y + blockIdx.y] = sh[0];
}

__device__ __forceinline__ void store_jacobian(uint4* out, const g1_xyzz_t& p) {
    fq_t X, Y, Z;
    g1_xyzz_to_jacobian(p, X, Y, Z);
    fq_to_u4x3(X, out + 0);
    fq_to_u4x3(Y, out + 3);
    fq_to_u4x3(Z, out + 6);
}

// Horner over the windows, high window first: total = sum_w 2^(c*w) * W_w.  One dependent chain per MSM, so it is run
// by a QUAD of lanes per MSM (ec_coop.cuh: the independent products of every doubling / addition on different lanes):
// 3 product latencies per doubling instead of 9, 4 per addition instead of 14.
__global__ void __launch_bounds__(32) msm_fold_kernel(uint4* __restrict__ out_jac, const g1_xyzz_mem_t* __restrict__ wsum,
                                                      MsmShape sh, uint32_t nmsm) {
    const uint32_t m = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    if (m >= nmsm) return;                                    // whole quads leave together
    const Quad Q = quad_here();
    const g1_xyzz_mem_t* w = wsum + (size_t)m * sh.nwin;
    g1_xyzz_t total = g1_xyzz_infinity();
    for (uint32_t i = sh.nwin; i-- > 0;) {
        for (uint32_t k = 0; k < sh.c; k++) g1_dbl_quad(Q, total);
        const g1_xyzz_t s = g1_xyzz_load(w + i);
        g1_add_quad(Q, total, s);
    }
    if (Q.q == 0) store_jacobian(out_jac + 9 * (size_t)m, total);
}

__global__ void msm_write_infinity_kernel(uint4* out_jac, uint32_t count) {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= count) return;
    store_jacobian(out_jac + 9 * (size_t)m, g1_xyzz_infinity());
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
// window table of a resident base set: table[w * n + i] = 2^(c*w) * P_i as packed affine points.  With 180 GB of HBM a
// 2^24-point SRS times 11 windows (17.7 GB) fits comfortably, and an MSM against the table needs no per-window bucket
// sets, no window fold and a single bucket reduction.  One thread per point: c doublings per window in XYZZ, then ONE
// field inversion for all windows of the point (Montgomery's trick over the ZZZ coordinates).
// ---------------------------------------------------------------------------------------------
#define MSM_TABLE_MAX_WINDOWS 17

__global__ void __launch_bounds__(64) msm_window_table_kernel(uint4* __restrict__ table, size_t n, uint32_t c, uint32_t nwin) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    g1_xyzz_t q[MSM_TABLE_MAX_WINDOWS];
    fq_t pre[MSM_TABLE_MAX_WINDOWS];
    g1_xyzz_t cur = g1_xyzz_from_affine(g1_unpack(g1_load_packed(table + i * G1_BASE_U4)));
    fq_t acc = fp_one<FqP>();
    for (uint32_t w = 1; w < nwin; w++) {
        for (uint32_t k = 0; k < c; k++) g1_dbl(cur);
        q[w] = cur;
        pre[w] = acc;                                             // product of the non-zero ZZZ before w
        if (!g1_xyzz_is_infinity(cur)) acc = fp_mul(acc, cur.ZZZ);
    }
    fq_t inv = fp_inv(acc);
    for (uint32_t w = nwin; w-- > 1;) {
        g1_packed_t out;
        if (g1_xyzz_is_infinity(q[w])) {
#pragma unroll
            for (int k = 0; k < 6; k++) out.w[k] = make_uint4(0, 0, 0, 0);
        } else {
            fq_t i3 = fp_mul(inv, pre[w]);                        // 1 / ZZZ_w
            inv = fp_mul(inv, q[w].ZZZ);
            fq_t i2 = fp_mul(fp_sqr(i3), fp_sqr(q[w].ZZ));        // 1 / ZZ_w  (ZZ^3 = ZZZ^2)
            fq_t x = fp_mul(q[w].X, i2), y = fp_mul(q[w].Y, i3);
            fq_to_u4x3(x, out.w);
            fq_to_u4x3(y, out.w + 3);
        }
        g1_store_packed(table + ((size_t)w * n + i) * G1_BASE_U4, out);
    }
}

b200_error_t msm_build_window_table_device(void* d_table, size_t n, uint32_t c, cudaStream_t stream) {
    const uint32_t nwin = msm_shape(c).nwin;
    if (nwin > MSM_TABLE_MAX_WINDOWS) return b200_err(B200_ERR_INVALID_ARG, "msm: window table needs c >= 16");
    if (n == 0) return b200_ok();
    msm_window_table_kernel<<<(unsigned)((n + 63) / 64), 64, 0, stream>>>(reinterpret_cast<uint4*>(d_table), n, c, nwin);
    KERNEL_CHECK();
    return b200_ok();
}

// ---------------------------------------------------------------------------------------------
// launcher
// ---------------------------------------------------------------------------------------------
extern "C" uint32_t b200_msm_window_bits(size_t n) {
    {
        const int c = b200_config().msm_c;
        if (c >= 2 && c <= 22) return (uint32_t)c;
    }
    uint32_t lg = 0;
    while (((size_t)1 << (lg + 1)) <= n) lg++;
    // measured sweeps with the batched-affine rounds on (gpurun_out/affine_sweep10.log -> profiles/r01_msm_affine_sweep.txt):
    // 2^20 -> 15, 2^22 -> 17, 2^24 -> 18 (82.7 ms; 17: 83.2, 20: 85.4).  c = 19 is avoided: 253 = 13 * 19 + 6 leaves a
    // 6-bit top window whose 64 buckets take n atomics each (count 4.6 ms instead of 1.3).
    int c = (int)lg - 5;
    if (c < 4) c = 4;
    if (c > 18) c = lg >= 25 ? 20 : 18;
    return (uint32_t)c;
}

b200_error_t msm_pack_bases_device(void* d_packed, const void* d_points, size_t n, size_t stride,
                                   cudaStream_t stream) {
    if (n == 0) return b200_ok();
    if (stride < 97 || (stride & 7)) return b200_err(B200_ERR_INVALID_ARG, "msm: affine stride must be >= 104 and 8-byte aligned");
    if (reinterpret_cast<uintptr_t>(d_points) & 7) return b200_err(B200_ERR_INVALID_ARG, "msm: points must be 8-byte aligned");
    msm_pack_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(reinterpret_cast<uint4*>(d_packed),
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
                            const void* d_packed, cudaStream_t stream, bool packed_glv) {
    return msm_run_batch_device(d_out, d_points, n, d_scalars, stride, d_packed, nullptr, 1, stream, 0, 0, packed_glv);
}

bool msm_glv_enabled() { return b200_config().msm_glv; }

// resident sets stored in GLV form: 2n records (P_i, phi(P_i))
b200_error_t msm_pack_bases_glv_device(void* d_packed_2n, const void* d_points, size_t n, size_t stride, cudaStream_t stream) {
    if (n == 0) return b200_ok();
    if (stride < 97 || (stride & 7)) return b200_err(B200_ERR_INVALID_ARG, "msm: affine stride must be >= 104 and 8-byte aligned");
    if (reinterpret_cast<uintptr_t>(d_points) & 7) return b200_err(B200_ERR_INVALID_ARG, "msm: points must be 8-byte aligned");
    msm_pack_glv_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(reinterpret_cast<uint4*>(d_packed_2n),
                                                                          reinterpret_cast<const uint8_t*>(d_points), n, stride);
    KERNEL_CHECK();
    return b200_ok();
}

b200_error_t msm_run_tabulated_device(void* d_out, size_t n, const void* d_scalars, const void* d_table, size_t n_reg,
                                      uint32_t c, cudaStream_t stream) {
    return msm_run_batch_device(d_out, nullptr, n, d_scalars, 0, d_table, nullptr, 1, stream, n_reg, c);
}

// ---------------------------------------------------------------------------------------------
// The pipeline in two halves, so that the host-buffer path can stream point ranges: `msm_front` turns one range of
// (points, scalars) into a full bucket array (pack .. combine), `msm_back` turns a bucket array into the result
// (reduce .. fold); bucket arrays of several ranges computed with the same window width add up bucket by bucket.
// ---------------------------------------------------------------------------------------------
struct MsmPlan {
    bool shared_bases = false;   // batched call whose MSMs all index ONE base set from 0 (k commits against one SRS)
    MsmShape sh;             // per-MSM shape (c, windows, buckets per window)
    MsmShape vsh;            // shape seen after the digit stage: nwin = virtual windows (nmsm * nwin, or 1 when tabulated)
    size_t K;                // total buckets
    uint32_t nmsm;
    size_t n_reg;            // != 0: tabulated bases, table row length
    uint32_t seg_len, segs_per_win;
    bool glv;                // inputs split by the endomorphism: 2n points, 127-bit scalars (msm_glv.cuh)
};

// GLV applies where the call packs its own points (VariableBase::msm proper, the streamed host path, the batched
// verifier MSMs); resident sets keep their stored form (tabulated sets have no fold to shorten anyway).
static bool msm_use_glv(const void* d_packed, size_t n_reg) {
    return !d_packed && !n_reg && b200_config().msm_glv;
}

// window width for 2n points with 127-bit scalars: the usual log2 - 5, then the narrowest window with the same
// number of windows (127 / c + 1 is a step function: 19 bits already give the 7 windows of 20 and 21)
static uint32_t msm_glv_window_bits(size_t n_eff) {
    {
        const int c = b200_config().msm_c;
        if (c >= 2 && c <= 22) return (uint32_t)c;
    }
    uint32_t lg = 0;
    while (((size_t)1 << (lg + 1)) <= n_eff) lg++;
    int c = (int)lg - 5;
    if (c < 4) c = 4;
    if (c > 19) c = 19;
    while (c > 4 && MSM_GLV_BITS / (c - 1) == MSM_GLV_BITS / c) c--;
    return (uint32_t)c;
}

// threads of msm_reduce_segments_kernel resident on the device at once (queried once per process)
static size_t msm_reduce_wave_threads() {
    static std::once_flag once;
    static size_t wave = 148 * 256;
    std::call_once(once, [] {
        int dev = 0, sms = 0, per_sm = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, msm_reduce_segments_kernel, MSM_RED_THREADS, 0) == cudaSuccess &&
            sms > 0 && per_sm > 0)
            wave = (size_t)sms * per_sm * MSM_RED_THREADS;
    });
    return wave;
}

static b200_error_t msm_make_plan(MsmPlan* pl, size_t n, uint32_t nmsm, size_t n_reg, uint32_t c_force, bool glv = false) {
    pl->glv = glv;
    if (glv) {
        n *= 2;
        pl->sh = msm_shape(c_force ? c_force : msm_glv_window_bits(nmsm > 1 ? (n + nmsm - 1) / nmsm : n), MSM_GLV_BITS);
    } else {
        pl->sh = msm_shape(c_force ? c_force : b200_msm_window_bits(nmsm > 1 ? (n + nmsm - 1) / nmsm : n));
    }
    pl->vsh = pl->sh;
    pl->vsh.nwin = n_reg ? nmsm : pl->sh.nwin * nmsm;        // tabulated: one bucket set per MSM, nothing to fold
    pl->K = (size_t)pl->vsh.nwin * pl->sh.nbuckets;
    pl->nmsm = nmsm;
    pl->n_reg = n_reg;
    if (pl->K >= ((size_t)1 << 31)) return b200_err(B200_ERR_TOO_LARGE, "msm: too many buckets (batch too large)");
    if ((size_t)n * pl->sh.nwin >= ((size_t)1 << 32)) return b200_err(B200_ERR_TOO_LARGE, "msm: n * windows overflows 32-bit offsets");
    // segment length of the running-sum reduction: short segments when there are few buckets (latency), long
    // ones when there are many (each segment pays a ~c-step double-and-add for its offset)
    // (up to 2^15 buckets -- a commit against <= 2^16 powers -- the reduction is ONE dependent chain per quad of lanes:
    // 4-bucket segments make it 8 additions + a 13-step offset instead of 16 + 12)
    // Above one wave of the reduction kernel (255 registers: 4 blocks of 64 threads per SM, 37888 threads on a B200)
    // the segments grow so that the grid stays ONE wave: the kernel is a dependent chain per thread at 8 warps per SM,
    // so a second wave costs a whole chain (2 * seg_len additions + the offset) while longer segments cost 2 additions
    // per bucket.  2^24 points (7 x 2^18 buckets): 49 buckets per segment 2.8 ms, 32 (1.5 waves) 4.3 ms, 64 3.4 ms
    // (tools/reduce_sweep.py, profiles/r02_reduce_sweep.jsonl).
    uint32_t seg_len = 8;
    const size_t wave = msm_reduce_wave_threads();
    if (pl->K / seg_len > wave) {
        seg_len = (uint32_t)((pl->K + wave - 1) / wave);
        while ((size_t)pl->vsh.nwin * ((pl->sh.nbuckets + seg_len - 1) / seg_len) > wave && seg_len < pl->sh.nbuckets) seg_len++;
    }
    if (b200_config().msm_seg_len > 0) seg_len = (uint32_t)b200_config().msm_seg_len;
    if (seg_len > pl->sh.nbuckets) seg_len = pl->sh.nbuckets;
    pl->seg_len = seg_len;
    pl->segs_per_win = (pl->sh.nbuckets + seg_len - 1) / seg_len;
    return b200_ok();
}

// events for the fork / join between the caller's stream and the helper stream of one call
struct EventPool {
    std::vector<cudaEvent_t> used;
    cudaError_t get(cudaEvent_t* ev) {
        cudaError_t e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
        if (e == cudaSuccess) used.push_back(*ev);
        return e;
    }
    ~EventPool() { for (cudaEvent_t e : used) cudaEventDestroy(e); }       // deferred by the runtime until complete
};

// list length after one pair round: sum_k ceil(m_k / 2) <= (E + #non-empty buckets) / 2
static size_t msm_halved_bound(size_t E, size_t K) { return (E + (K < E ? K : E) + 1) / 2; }

// Number of batched-affine rounds before the XYZZ finish: enough to leave lists of 2..4 points.  Small calls are
// launch-bound (every round is ~10 launches and one serial Fermat inversion), they keep the one-kernel path.
static uint32_t msm_affine_rounds(size_t E, size_t K) {
    {
        const int r = b200_config().msm_affine_rounds;
        if (r >= 0) return r > 30 ? 30u : (uint32_t)r;
    }
    // round r adds E / 2^(r+1) pairs at ~0.29 ns instead of ~0.40 ns each and costs ~0.4 ms of launches and the
    // latency-bound inversion chain: worth it above ~3.5 M pairs (measured optimum: 0 rounds at 2^18, 2 at 2^20,
    // 4 at 2^22, 5 at 2^24; the XYZZ finish is cheap on lists of 4..8)
    const size_t avg = E / K;
    uint32_t r = 0;
    while (r < 5 && (E >> (r + 1)) >= ((size_t)7 << 19) && (avg >> r) >= 4) r++;
    return r;
}

extern "C" uint32_t b200_msm_affine_rounds(size_t n) {
    const MsmShape sh = msm_shape(b200_msm_window_bits(n));
    return msm_affine_rounds(n * sh.nwin, (size_t)sh.nwin * sh.nbuckets);
}

// what a plain VariableBase::msm call of n points runs with: out = {window bits, windows, pair rounds, GLV (0 / 1)}
extern "C" void b200_msm_describe(size_t n, uint32_t* out4) {
    const bool glv = msm_use_glv(nullptr, 0);
    const size_t n_eff = glv ? 2 * n : n;
    const MsmShape sh = glv ? msm_shape(msm_glv_window_bits(n_eff), MSM_GLV_BITS) : msm_shape(b200_msm_window_bits(n));
    out4[0] = sh.c;
    out4[1] = sh.nwin;
    out4[2] = msm_affine_rounds(n_eff * sh.nwin, (size_t)sh.nwin * sh.nbuckets);
    out4[3] = glv ? 1u : 0u;
}

// blocks of msm_combine_coop_kernel that are resident at once on the current device (0: cooperative launch unsupported)
static unsigned msm_combine_coop_capacity() {
    static std::atomic<int> cached{-1};
    int v = cached.load(std::memory_order_relaxed);
    if (v >= 0) return (unsigned)v;
    int dev = 0, coop = 0, sms = 0, per_sm = 0;
    v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) == cudaSuccess && coop &&
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, msm_combine_coop_kernel, MSM_ACC_THREADS, 0) == cudaSuccess)
        v = sms * per_sm;
    else
        (void)cudaGetLastError();
    cached.store(v, std::memory_order_relaxed);
    return (unsigned)v;
}

// bucket array (K x XYZZ) of one range of points; d_buckets is overwritten
static b200_error_t msm_front(const MsmPlan& pl, g1_xyzz_mem_t* d_buckets, const void* d_points, size_t n,
                              const void* d_scalars, size_t stride, const void* d_packed,
                              const unsigned long long* d_seg_off, cudaStream_t stream, bool add_to = false) {
    const MsmShape& sh = pl.sh;
    const size_t K = pl.K;
    DevBuf packed, halves;
    const uint4* pts = reinterpret_cast<const uint4*>(d_packed);
    const uint32_t glv = pl.glv ? 1u : 0u;
    if (glv) {
        STAGE("msm_pack", stream);
        if (!pts) {                                        // resident GLV sets come packed as (P_i, phi(P_i)) already
            CUDA_TRY(packed.alloc(2 * n * (size_t)G1_BASE_BYTES, stream));
            B200_TRY(msm_pack_bases_glv_device(packed.p, d_points, n, stride, stream));
            pts = packed.as<uint4>();
        }
        CUDA_TRY(halves.alloc(2 * n * 16, stream));
        msm_glv_split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(halves.as<uint4>(), reinterpret_cast<const uint4*>(d_scalars), n);
        KERNEL_CHECK();
        d_scalars = halves.p;
        n *= 2;                                            // from here on: 2n points, 128-bit scalars
    } else if (!pts) {
        STAGE("msm_pack", stream);
        CUDA_TRY(packed.alloc(n * (size_t)G1_BASE_BYTES, stream));
        B200_TRY(msm_pack_bases_device(packed.p, d_points, n, stride, stream));
        pts = packed.as<uint4>();
    }
    DevBuf counts, offsets, cursor, entries, heads, tails, head_bucket, tail_bucket, max_heads;
    CUDA_TRY(counts.alloc((K + 1) * 4, stream));
    CUDA_TRY(offsets.alloc((K + 1) * 4, stream));
    CUDA_TRY(cursor.alloc((K + 1) * 4, stream));
    CUDA_TRY(entries.alloc(n * sh.nwin * 4, stream));

    STAGE("msm_count", stream);
    CUDA_TRY(cudaMemsetAsync(counts.p, 0, (K + 1) * 4, stream));
    const unsigned nblk = (unsigned)((n + 255) / 256);
    msm_count_kernel<<<nblk, 256, 0, stream>>>(counts.as<uint32_t>(), reinterpret_cast<const uint4*>(d_scalars), n, sh, d_seg_off, pl.nmsm, pl.n_reg, glv);
    KERNEL_CHECK();
    STAGE("msm_scan", stream);
    B200_TRY(exclusive_scan(offsets.as<uint32_t>(), counts.as<uint32_t>(), K, stream));
    CUDA_TRY(cudaMemcpyAsync(cursor.p, offsets.p, (K + 1) * 4, cudaMemcpyDeviceToDevice, stream));
    STAGE("msm_scatter", stream);
    msm_scatter_kernel<<<dim3(nblk, sh.nwin), 256, 0, stream>>>(entries.as<uint32_t>(), cursor.as<uint32_t>(),
                                                 reinterpret_cast<const uint4*>(d_scalars), n, sh, d_seg_off, pl.nmsm, pl.n_reg, glv,
                                                 pl.shared_bases ? 1u : 0u);
    KERNEL_CHECK();
    // ---- batched-affine pair rounds: halve every bucket's list `rounds` times at ~6.3 products per addition ----
    const size_t E = n * sh.nwin;                                   // upper bound on the number of entries
    const uint4* acc_pts = pts;
    const uint4* acc_pts_y = nullptr;                              // lists: plane of y coordinates (x plane = acc_pts)
    const uint32_t* acc_off = offsets.as<uint32_t>();
    size_t acc_E = E;
    bool direct = false;
    DevBuf list_a, list_b, pre, partial, half, off_a, off_b;
    uint32_t rounds = msm_affine_rounds(E, K);
    const size_t e1 = msm_halved_bound(E, K), e2 = msm_halved_bound(e1, K);
    if (rounds) {
        const size_t t1 = (e1 + MSM_PAIRS_PER_THREAD - 1) / MSM_PAIRS_PER_THREAD;
        const unsigned long long need = (unsigned long long)e1 * sizeof(g1_packed_t) + (rounds > 1 ? (unsigned long long)e2 * sizeof(g1_packed_t) : 0) +
                                        (unsigned long long)e1 * 48 + (unsigned long long)t1 * 48 + 3ull * (K + 1) * 4;
        const unsigned long long budget = b200_config().msm_list_budget;
        cudaError_t e = (budget && need > budget) ? cudaErrorMemoryAllocation : list_a.alloc(e1 * sizeof(g1_packed_t), stream);
        if (e == cudaSuccess && rounds > 1) e = list_b.alloc(e2 * sizeof(g1_packed_t), stream);
        if (e == cudaSuccess) e = pre.alloc(e1 * 48, stream);
        if (e == cudaSuccess) e = partial.alloc(t1 * 48, stream);
        if (e == cudaSuccess) e = half.alloc((K + 1) * 4, stream);
        if (e == cudaSuccess) e = off_a.alloc((K + 1) * 4, stream);
        if (e == cudaSuccess) e = off_b.alloc((K + 1) * 4, stream);
        if (e != cudaSuccess) {
            if (e != cudaErrorMemoryAllocation) return b200_cuda_err(e);
            (void)cudaGetLastError();                                 // not enough HBM for the lists: XYZZ path only
            list_a.release(); list_b.release(); pre.release(); partial.release(); half.release(); off_a.release(); off_b.release();
            rounds = 0;
            g_counters.msm_xyzz_fallbacks.fetch_add(1, std::memory_order_relaxed);     // observable: b200_get_counter
        }
    }
    // A round CAN be issued in slices of its output range (B200_MSM_SLICES > 1): denominators + inversion of a slice
    // go to the thread's high-priority helper stream, the additions to the caller's stream, so that the gather-bound
    // first half of slice i + 1 runs under the multiplier-bound second half of slice i.  Measured at 2^24: 84.4 ms
    // with 1 slice, 85.6 / 86.9 / 89.2 ms with 2 / 3 / 4 -- both halves gather at random and the memory system is what
    // they share, while every extra slice adds one latency-bound inversion chain.  Default: one slice, one stream.
    const uint32_t max_slices = b200_config().msm_slices > 1 ? (uint32_t)b200_config().msm_slices : 1u;
    cudaStream_t aux = rounds && max_slices > 1 ? b200_thread_aux_stream() : nullptr;
    EventPool events;
    for (uint32_t r = 0; r < rounds; r++) {
        STAGE(r == 0 ? "msm_pairs_round0" : "msm_pairs_rounds", stream);
        const size_t e_out = msm_halved_bound(acc_E, K);
        uint32_t* noff = (r & 1) ? off_b.as<uint32_t>() : off_a.as<uint32_t>();
        uint4* out_x = (r & 1) ? list_b.as<uint4>() : list_a.as<uint4>();
        uint4* out_y = out_x + 3 * ((r & 1) ? e2 : e1);                // plane of y coordinates behind the x plane
        msm_half_counts_kernel<<<(unsigned)((K + 255) / 256), 256, 0, stream>>>(half.as<uint32_t>(), acc_off, (uint32_t)K);
        KERNEL_CHECK();
        B200_TRY(exclusive_scan(noff, half.as<uint32_t>(), K, stream));
        PairRound rd;
        rd.src = acc_pts;
        rd.src_y = acc_pts_y;
        rd.entries = direct ? nullptr : entries.as<uint32_t>();
        rd.off = acc_off;
        rd.noff = noff;
        rd.K = (uint32_t)K;
        const uint32_t nthr = (uint32_t)((e_out + MSM_PAIRS_PER_THREAD - 1) / MSM_PAIRS_PER_THREAD);
        uint32_t nslices = aux ? nthr / (1u << 17) : 1;
        if (nslices < 1) nslices = 1;
        if (nslices > max_slices) nslices = max_slices;
        cudaStream_t front = nslices > 1 ? aux : stream;
        if (front != stream) {
            cudaEvent_t ev;
            CUDA_TRY(events.get(&ev));
            CUDA_TRY(cudaEventRecord(ev, stream));                    // offsets of this round, lists of the previous one
            CUDA_TRY(cudaStreamWaitEvent(front, ev, 0));
        }
        const uint32_t per = ((nthr + nslices - 1) / nslices + MSM_PAIR_THREADS - 1) / MSM_PAIR_THREADS * MSM_PAIR_THREADS;
        for (uint32_t t0 = 0; t0 < nthr; t0 += per) {
            const uint32_t t1 = t0 + per < nthr ? t0 + per : nthr;
            const unsigned nblk_p = (t1 - t0 + MSM_PAIR_THREADS - 1) / MSM_PAIR_THREADS;
            if (nslices == 1) STAGE(r == 0 ? "msm_pairs0_denoms" : "msm_pairs_denoms", stream);
            msm_pair_denoms_kernel<<<nblk_p, MSM_PAIR_THREADS, 0, front>>>(rd, pre.as<uint4>(), partial.as<uint4>(), t0, t1);
            KERNEL_CHECK();
            if (nslices == 1) STAGE(r == 0 ? "msm_pairs0_invert" : "msm_pairs_invert", stream);
            B200_TRY(fq_batch_inverse_nonzero(partial.as<uint4>() + 3 * (size_t)t0, t1 - t0, front));
            if (nslices == 1) STAGE(r == 0 ? "msm_pairs0_add" : "msm_pairs_add", stream);
            if (front != stream) {
                cudaEvent_t ev;
                CUDA_TRY(events.get(&ev));
                CUDA_TRY(cudaEventRecord(ev, front));
                CUDA_TRY(cudaStreamWaitEvent(stream, ev, 0));
            }
            msm_pair_add_kernel<<<nblk_p, MSM_PAIR_THREADS, 0, stream>>>(rd, pre.as<uint4>(), partial.as<uint4>(), out_x, out_y, t0, t1);
            KERNEL_CHECK();
        }
        acc_pts = out_x;
        acc_pts_y = out_y;
        acc_off = noff;
        acc_E = e_out;
        direct = true;
    }
    // ---- equal-work XYZZ accumulation of what is left ----
    STAGE("msm_accumulate", stream);
    // entries per thread of the XYZZ walk: 128 keeps the head / tail stitching negligible on large calls; small calls
    // are a latency problem (a thread's additions are one dependent chain), shorter chunks put more threads to work
    // (2^14 points: 1.29 ms at 128, 0.18 ms at 16; 2^16: 1.36 -> 0.50 ms at 32; from 2^22 entries on 128 wins: 1.91 vs 2.30 ms at 64)
    uint32_t chunk = acc_E >= ((size_t)1 << 22) ? 128 : acc_E >= ((size_t)1 << 21) ? 64 : acc_E >= ((size_t)1 << 19) ? 32 : 16;
    if (b200_config().msm_chunk > 0) chunk = (uint32_t)b200_config().msm_chunk;
    if (chunk < 8) chunk = 8;
    const size_t t_max = (acc_E + chunk - 1) / chunk;
    CUDA_TRY(heads.alloc((t_max + 1) * sizeof(g1_xyzz_mem_t), stream));
    CUDA_TRY(tails.alloc(t_max * sizeof(g1_xyzz_mem_t), stream));
    CUDA_TRY(head_bucket.alloc((t_max + 1) * 4, stream));
    CUDA_TRY(tail_bucket.alloc(t_max * 4, stream));
    CUDA_TRY(max_heads.alloc(16, stream));
    CUDA_TRY(cudaMemsetAsync(max_heads.p, 0, 16, stream));
    CUDA_TRY(cudaMemsetAsync(head_bucket.p, 0xff, (t_max + 1) * 4, stream));
    CUDA_TRY(cudaMemsetAsync(tail_bucket.p, 0xff, t_max * 4, stream));
    if (!add_to) CUDA_TRY(cudaMemsetAsync(d_buckets, 0, K * sizeof(g1_xyzz_mem_t), stream));      // ZZ = 0: empty buckets are infinity
    const unsigned tblocks = (unsigned)((t_max + MSM_ACC_THREADS - 1) / MSM_ACC_THREADS);
#define MSM_ACC_ARGS_DIRECT d_buckets, heads.as<g1_xyzz_mem_t>(), tails.as<g1_xyzz_mem_t>(), head_bucket.as<uint32_t>(), \
                            tail_bucket.as<uint32_t>(), acc_pts, acc_pts_y, nullptr, acc_off, max_heads.as<uint32_t>(), (uint32_t)K, chunk
#define MSM_ACC_ARGS_INDEXED d_buckets, heads.as<g1_xyzz_mem_t>(), tails.as<g1_xyzz_mem_t>(), head_bucket.as<uint32_t>(), \
                             tail_bucket.as<uint32_t>(), pts, nullptr, entries.as<uint32_t>(), acc_off, max_heads.as<uint32_t>(), (uint32_t)K, chunk
    if (direct && add_to) msm_accumulate_kernel<true, true><<<tblocks, MSM_ACC_THREADS, 0, stream>>>(MSM_ACC_ARGS_DIRECT);
    else if (direct) msm_accumulate_kernel<true, false><<<tblocks, MSM_ACC_THREADS, 0, stream>>>(MSM_ACC_ARGS_DIRECT);
    else if (add_to) msm_accumulate_kernel<false, true><<<tblocks, MSM_ACC_THREADS, 0, stream>>>(MSM_ACC_ARGS_INDEXED);
    else msm_accumulate_kernel<false, false><<<tblocks, MSM_ACC_THREADS, 0, stream>>>(MSM_ACC_ARGS_INDEXED);
#undef MSM_ACC_ARGS_DIRECT
#undef MSM_ACC_ARGS_INDEXED
    KERNEL_CHECK();
    STAGE("msm_combine", stream);
    // longest possible run of heads: a bucket holds at most all acc_E entries of the stream (with tabulated bases every
    // window of every point lands in ONE bucket set, so "at most n" does not hold: n * nwin equal digits are possible).
    // Rounds above the device-side maximum return at once, so the extra launches of the bound are cheap.
    const size_t max_span = (acc_E + chunk - 1) / chunk + 1;
    if (tblocks <= msm_combine_coop_capacity()) {
        g1_xyzz_mem_t* a_buckets = d_buckets;
        g1_xyzz_mem_t* a_heads = heads.as<g1_xyzz_mem_t>();
        const g1_xyzz_mem_t* a_tails = tails.as<g1_xyzz_mem_t>();
        const uint32_t* a_hb = head_bucket.as<uint32_t>();
        const uint32_t* a_tb = tail_bucket.as<uint32_t>();
        const uint32_t* a_off = acc_off;
        const uint32_t* a_mh = max_heads.as<uint32_t>();
        uint32_t a_tmax = (uint32_t)t_max, a_chunk = chunk;
        void* args[] = {&a_buckets, &a_heads, &a_tails, &a_hb, &a_tb, &a_off, &a_mh, &a_tmax, &a_chunk};
        CUDA_TRY(cudaLaunchCooperativeKernel((const void*)msm_combine_coop_kernel, dim3(tblocks), dim3(MSM_ACC_THREADS), args, 0, stream));
        B200_LAUNCH_COUNT();
        return b200_ok();
    }
    for (uint32_t stride2 = 1; stride2 < max_span; stride2 <<= 1) {
        msm_combine_heads_kernel<<<tblocks, MSM_ACC_THREADS, 0, stream>>>(heads.as<g1_xyzz_mem_t>(), head_bucket.as<uint32_t>(),
                                                                         acc_off, max_heads.as<uint32_t>(), (uint32_t)t_max, chunk, stride2);
        KERNEL_CHECK();
    }
    msm_combine_tails_kernel<<<tblocks, MSM_ACC_THREADS, 0, stream>>>(d_buckets, heads.as<g1_xyzz_mem_t>(),
                                                                     tails.as<g1_xyzz_mem_t>(), tail_bucket.as<uint32_t>(), (uint32_t)t_max);
    KERNEL_CHECK();
    return b200_ok();
}

// total[k] += part[k] for every bucket (streamed point ranges share one bucket array)
__global__ void __launch_bounds__(MSM_ACC_THREADS) msm_merge_buckets_kernel(g1_xyzz_mem_t* __restrict__ total,
                                                                           const g1_xyzz_mem_t* __restrict__ part, uint32_t K) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    g1_xyzz_t b = g1_xyzz_load(part + k);
    if (g1_xyzz_is_infinity(b)) return;
    g1_xyzz_t a = g1_xyzz_load(total + k);
    g1_add(a, b);
    g1_xyzz_store(total + k, a);
}

// result(s) from a bucket array: per-window running sums, window sums, fold
static b200_error_t msm_back(const MsmPlan& pl, void* d_out, const g1_xyzz_mem_t* d_buckets, cudaStream_t stream) {
    const MsmShape& vsh = pl.vsh;
    DevBuf segs, wsum;
    CUDA_TRY(segs.alloc((size_t)pl.segs_per_win * vsh.nwin * sizeof(g1_xyzz_mem_t), stream));
    CUDA_TRY(wsum.alloc((size_t)vsh.nwin * sizeof(g1_xyzz_mem_t), stream));
    STAGE("msm_reduce_segments", stream);
    const uint32_t nseg_threads = pl.segs_per_win * vsh.nwin;
    const bool quads = nseg_threads <= (uint32_t)b200_config().msm_reduce_quad_max;
    if (quads)                         // latency-bound: four lanes per segment
        msm_reduce_segments_quad_kernel<<<(4 * nseg_threads + MSM_RED_THREADS - 1) / MSM_RED_THREADS, MSM_RED_THREADS, 0, stream>>>(
            segs.as<g1_xyzz_mem_t>(), d_buckets, vsh, pl.seg_len, pl.segs_per_win);
    else
        msm_reduce_segments_kernel<<<(nseg_threads + MSM_RED_THREADS - 1) / MSM_RED_THREADS, MSM_RED_THREADS, 0, stream>>>(
            segs.as<g1_xyzz_mem_t>(), d_buckets, vsh, pl.seg_len, pl.segs_per_win);
    KERNEL_CHECK();
    STAGE("msm_window_sum", stream);
    if (quads && pl.segs_per_win > 1) {
        // latency-bound: quads, 128 inputs per block, then the (<= 64) slice sums of each window
        const uint32_t per_block = 128;
        const uint32_t slices = (pl.segs_per_win + per_block - 1) / per_block;
        if (slices > 1) {
            DevBuf slice_sums;
            CUDA_TRY(slice_sums.alloc((size_t)slices * vsh.nwin * sizeof(g1_xyzz_mem_t), stream));
            msm_window_sum_quad_kernel<<<dim3(vsh.nwin, slices), MSM_TREE_THREADS, 0, stream>>>(
                slice_sums.as<g1_xyzz_mem_t>(), segs.as<g1_xyzz_mem_t>(), pl.segs_per_win, per_block);
            KERNEL_CHECK();
            msm_window_sum_quad_kernel<<<dim3(vsh.nwin, 1), MSM_TREE_THREADS, 0, stream>>>(
                wsum.as<g1_xyzz_mem_t>(), slice_sums.as<g1_xyzz_mem_t>(), slices, slices);
            KERNEL_CHECK();
        } else {
            msm_window_sum_quad_kernel<<<dim3(vsh.nwin, 1), MSM_TREE_THREADS, 0, stream>>>(
                wsum.as<g1_xyzz_mem_t>(), segs.as<g1_xyzz_mem_t>(), pl.segs_per_win, pl.segs_per_win);
            KERNEL_CHECK();
        }
    } else {
        // level 1: slices of >= 4 * MSM_TREE_THREADS segment results per block; level 2: the slice sums of each window
        const uint32_t per_block = 4 * MSM_TREE_THREADS;
        const uint32_t slices = (pl.segs_per_win + per_block - 1) / per_block;
        if (slices > 1) {
            DevBuf slice_sums;
            CUDA_TRY(slice_sums.alloc((size_t)slices * vsh.nwin * sizeof(g1_xyzz_mem_t), stream));
            msm_window_sum_kernel<<<dim3(vsh.nwin, slices), MSM_TREE_THREADS, 0, stream>>>(
                slice_sums.as<g1_xyzz_mem_t>(), segs.as<g1_xyzz_mem_t>(), pl.segs_per_win, per_block);
            KERNEL_CHECK();
            msm_window_sum_kernel<<<dim3(vsh.nwin, 1), MSM_TREE_THREADS, 0, stream>>>(
                wsum.as<g1_xyzz_mem_t>(), slice_sums.as<g1_xyzz_mem_t>(), slices, slices);
            KERNEL_CHECK();
        } else {
            msm_window_sum_kernel<<<dim3(vsh.nwin, 1), MSM_TREE_THREADS, 0, stream>>>(
                wsum.as<g1_xyzz_mem_t>(), segs.as<g1_xyzz_mem_t>(), pl.segs_per_win, pl.segs_per_win);
            KERNEL_CHECK();
        }
    }
    STAGE("msm_fold", stream);
    MsmShape fold_sh = pl.sh;
    if (pl.n_reg) fold_sh.nwin = 1;                          // tabulated: the single bucket set already carries 2^(c*w)
    msm_fold_kernel<<<(pl.nmsm + 7) / 8, 32, 0, stream>>>(reinterpret_cast<uint4*>(d_out), wsum.as<g1_xyzz_mem_t>(),
                                                          fold_sh, pl.nmsm);
    KERNEL_CHECK();
    STAGE_END(stream);
    return b200_ok();
}

// nmsm independent MSMs over consecutive point ranges [seg_off[m], seg_off[m + 1]) of one (points, scalars) pair;
// d_out receives nmsm Jacobian points.  d_seg_off == nullptr means a single MSM over everything.
b200_error_t msm_run_batch_device(void* d_out, const void* d_points, size_t n, const void* d_scalars, size_t stride,
                                  const void* d_packed, const unsigned long long* d_seg_off, uint32_t nmsm,
                                  cudaStream_t stream, size_t n_reg, uint32_t c_tab, bool packed_glv, bool shared_bases) {
    if (!d_out) return b200_err(B200_ERR_INVALID_ARG, "msm: null output pointer");
    if (n_reg && (!d_packed || c_tab < 2 || (nmsm != 1 && !shared_bases) || (nmsm == 1 && n > n_reg)))
        return b200_err(B200_ERR_INVALID_ARG, "msm: tabulated bases need MSMs over a prefix of the table");
    if (shared_bases && (!d_packed || !d_seg_off)) return b200_err(B200_ERR_INVALID_ARG, "msm: shared bases need a resident set and offsets");
    if (nmsm == 0) return b200_ok();
    if (n == 0) {
        msm_write_infinity_kernel<<<(nmsm + 63) / 64, 64, 0, stream>>>(reinterpret_cast<uint4*>(d_out), nmsm);
        KERNEL_CHECK();
        return b200_ok();
    }
    if (!d_scalars || (!d_points && !d_packed)) return b200_err(B200_ERR_INVALID_ARG, "msm: null input pointer");
    if (n >= ((size_t)1 << 28)) return b200_err(B200_ERR_TOO_LARGE, "msm: more than 2^28 - 1 points per call");
    if (reinterpret_cast<uintptr_t>(d_scalars) & 15) return b200_err(B200_ERR_INVALID_ARG, "msm: scalars must be 16-byte aligned on the device");
    MsmPlan pl;
    B200_TRY(msm_make_plan(&pl, n, nmsm, n_reg, n_reg ? c_tab : 0, packed_glv || msm_use_glv(d_packed, n_reg)));
    pl.shared_bases = shared_bases;
    DevBuf buckets;
    CUDA_TRY(buckets.alloc(pl.K * sizeof(g1_xyzz_mem_t), stream));
    B200_TRY(msm_front(pl, buckets.as<g1_xyzz_mem_t>(), d_points, n, d_scalars, stride, d_packed, d_seg_off, stream));
    return msm_back(pl, d_out, buckets.as<g1_xyzz_mem_t>(), stream);
}

// Streaming form for the host-buffer path: a session fixes the window width for n_total points; every range is
// accumulated into a bucket array of its own, on whatever stream the caller gives it, and added to THAT stream's running
// total (streams are independent until the finish: the caller alternates two, so that the latency-bound chains of one
// range run under the multiplier-bound additions of the other); finish adds the totals up and reduces once.
// (Tried and dropped: keeping every range's half-reduced point lists, concatenating them bucket by bucket at the finish
// and running the last pair rounds + the XYZZ walk once -- 3.0 instead of 9.8 ms of XYZZ walks and no merges, but the
// serial tail of concatenation + rounds and the extra inversion chains cost as much: 93.3 against 90.3 ms at 2^24.)
struct MsmStream {
    MsmPlan plan;
    cudaStream_t main = nullptr;
    std::vector<cudaStream_t> others;                          // streams other than `main` that ranges ran on
    std::vector<std::pair<cudaStream_t, std::unique_ptr<DevBuf>>> totals;      // per stream: the sum of its ranges' bucket arrays
    ~MsmStream() {
        // blocks allocated on another stream go back to THAT stream's cache: its later work must not reuse them before
        // the finish (enqueued on `main`) has read them
        if (!main || others.empty()) return;
        cudaEvent_t ev;
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { cudaStreamSynchronize(main); return; }
        cudaEventRecord(ev, main);
        for (cudaStream_t o : others) cudaStreamWaitEvent(o, ev, 0);
        cudaEventDestroy(ev);
    }
};

b200_error_t msm_stream_begin(void** session, size_t n_total, cudaStream_t stream) {
    MsmStream* st = new MsmStream();
    st->main = stream;
    // one window narrower than the single-shot choice: every range pays one group addition per touched bucket when it
    // is merged, so fewer, fuller buckets win (2^24 in 2^22 ranges: c = 19)
    const bool glv = msm_use_glv(nullptr, 0);
    uint32_t c = glv ? msm_glv_window_bits(2 * n_total) : b200_msm_window_bits(n_total);
    if (!glv && !(b200_config().msm_c >= 2 && b200_config().msm_c <= 22) && c > 6) c -= 1;
    b200_error_t r = msm_make_plan(&st->plan, n_total, 1, 0, c, glv);
    if (r.code != 0) { delete st; return r; }
    *session = st;
    return b200_ok();
}

static b200_error_t msm_stream_merge(const MsmPlan& pl, g1_xyzz_mem_t* total, const g1_xyzz_mem_t* part, cudaStream_t stream) {
    STAGE("msm_merge", stream);
    const uint32_t K = (uint32_t)pl.K;
    msm_merge_buckets_kernel<<<(K + MSM_ACC_THREADS - 1) / MSM_ACC_THREADS, MSM_ACC_THREADS, 0, stream>>>(total, part, K);
    KERNEL_CHECK();
    return b200_ok();
}

// `stream` may differ from range to range; the caller orders msm_stream_finish behind all of them
b200_error_t msm_stream_add(void* session, const void* d_points, size_t n, const void* d_scalars, size_t stride,
                            cudaStream_t stream) {
    MsmStream* st = reinterpret_cast<MsmStream*>(session);
    if (n == 0) return b200_ok();
    if (stream != st->main && std::find(st->others.begin(), st->others.end(), stream) == st->others.end()) st->others.push_back(stream);
    for (auto& t : st->totals)
        if (t.first == stream) {
            // this stream's running total (no cross-stream dependency): the range's XYZZ walk starts every bucket from
            // the sum so far (msm_accumulate_kernel<.., ADD>), so there is no bucket array of its own and no merge pass
            B200_TRY(msm_front(st->plan, t.second->as<g1_xyzz_mem_t>(), d_points, n, d_scalars, stride, nullptr, nullptr, stream, true));
            STAGE_END(stream);                                // stages are per stream: the next one here may be a range away
            return b200_ok();
        }
    std::unique_ptr<DevBuf> buckets(new DevBuf());
    CUDA_TRY(buckets->alloc(st->plan.K * sizeof(g1_xyzz_mem_t), stream));
    B200_TRY(msm_front(st->plan, buckets->as<g1_xyzz_mem_t>(), d_points, n, d_scalars, stride, nullptr, nullptr, stream));
    STAGE_END(stream);
    st->totals.emplace_back(stream, std::move(buckets));
    return b200_ok();
}

b200_error_t msm_stream_finish(void* session, void* d_out, cudaStream_t stream) {
    std::unique_ptr<MsmStream> st(reinterpret_cast<MsmStream*>(session));      // stream-ordered frees on every path
    if (st->totals.empty()) {
        msm_write_infinity_kernel<<<1, 64, 0, stream>>>(reinterpret_cast<uint4*>(d_out), 1);
        KERNEL_CHECK();
        return b200_ok();
    }
    g1_xyzz_mem_t* total = st->totals[0].second->as<g1_xyzz_mem_t>();
    for (size_t i = 1; i < st->totals.size(); i++)
        B200_TRY(msm_stream_merge(st->plan, total, st->totals[i].second->as<g1_xyzz_mem_t>(), stream));
    return msm_back(st->plan, d_out, total, stream);
}


void msm_stream_abort(void* session) { delete reinterpret_cast<MsmStream*>(session); }

extern "C" b200_error_t b200_g1_sum_jacobian_device(void* d_out, const void* d_in, size_t count, void* stream) {
    B200_TRY(b200_require_device());
    if (!d_out || (count && !d_in)) return b200_err(B200_ERR_INVALID_ARG, "g1_sum: null pointer");
    g1_sum_jacobian_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint4*>(d_out),
                                                              reinterpret_cast<const uint4*>(d_in), (uint32_t)count);
    KERNEL_CHECK();
    return b200_ok();
}
