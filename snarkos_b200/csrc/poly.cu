// poly.cu -- the polynomial glue snarkVM's Varuna prover runs between (i)FFTs, kept on the device so evaluation
// vectors never leave HBM between a coset FFT, the pointwise stage and the inverse FFT:
//   Evaluations * / + / - Evaluations, scalar scaling and fused multiply-add   [UPSTREAM algorithms/src/fft/evaluations.rs]
//   batch_inversion (Montgomery's trick, zeros stay zero)                      [UPSTREAM fields/src/traits/field.rs]
//   division by a vanishing polynomial on a coset                              [UPSTREAM algorithms/src/fft/domain.rs:
//        evaluate_vanishing_polynomial / divide_by_vanishing_poly_on_coset_in_place]
// All data are Montgomery Fr (32 B), results fully reduced; every kernel is a single HBM pass.
#include <cstring>

#include "common.cuh"
#include "ntt_core.cuh"

// ---------------------------------------------------------------------------------------------
// element-wise:  op 0: a*b   1: a+b   2: a-b   3: a*b + c   4: a*b - c      (b: vector or broadcast scalar)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fr_vec_op_kernel(int op, uint4* __restrict__ out, const uint4* __restrict__ a,
                                                        const uint4* __restrict__ b, const uint4* __restrict__ c, size_t n,
                                                        int b_is_scalar) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_t x = fr_load(a, i);
    fr_t y = fr_load(b, b_is_scalar ? 0 : i);
    fr_t r;
    if (op == 0) r = fp_mul(x, y);
    else if (op == 1) r = fp_add(x, y);
    else if (op == 2) r = fp_sub(x, y);
    else {
        r = fp_mul(x, y);
        fr_t z = fr_load(c, i);
        r = (op == 3) ? fp_add(r, z) : fp_sub(r, z);
    }
    uint4 lo, hi;
    fr_to_u4(r, lo, hi);
    out[2 * i] = lo;
    out[2 * i + 1] = hi;
}

extern "C" b200_error_t b200_fr_vec_op_device(int op, void* d_out, const void* d_a, const void* d_b, const void* d_c,
                                              size_t n, int b_is_scalar, void* stream) {
    B200_TRY(b200_require_device());
    if (op < 0 || op > 4) return b200_err(B200_ERR_INVALID_ARG, "fr_vec_op: unknown op");
    if (n == 0) return b200_ok();
    if (!d_out || !d_a || !d_b || (op >= 3 && !d_c)) return b200_err(B200_ERR_INVALID_ARG, "fr_vec_op: null pointer");
    fr_vec_op_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        op, reinterpret_cast<uint4*>(d_out), reinterpret_cast<const uint4*>(d_a), reinterpret_cast<const uint4*>(d_b),
        reinterpret_cast<const uint4*>(d_c), n, b_is_scalar);
    KERNEL_CHECK();
    return b200_ok();
}

// ---------------------------------------------------------------------------------------------
// batch inversion: three-kernel Montgomery trick with a recursive middle
//   up   : every thread multiplies its chunk of BI_CHUNK elements (zeros count as 1) -> partial[t]
//   (recursively invert `partial`; below BI_SMALL elements one block does it serially-in-parallel)
//   down : every thread walks its chunk backwards turning inv(partial[t]) into the element inverses
// 3 products per element + 1/BI_CHUNK of the same again; zeros are left at zero like snarkVM's batch_inversion.
// ---------------------------------------------------------------------------------------------
#define BI_CHUNK 32
#define BI_SMALL 2048

__global__ void __launch_bounds__(128) batch_inv_up_kernel(uint4* __restrict__ partial, const uint4* __restrict__ data, size_t n) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = t * BI_CHUNK;
    if (lo >= n) return;
    size_t hi = lo + BI_CHUNK < n ? lo + BI_CHUNK : n;
    fr_t acc = fp_one<FrP>();
    for (size_t i = lo; i < hi; i++) {
        fr_t x = fr_load(data, i);
        if (!fp_is_zero(x)) acc = fp_mul(acc, x);
    }
    uint4 a, b;
    fr_to_u4(acc, a, b);
    partial[2 * t] = a;
    partial[2 * t + 1] = b;
}

__global__ void __launch_bounds__(128) batch_inv_down_kernel(uint4* __restrict__ data, const uint4* __restrict__ partial_inv, size_t n) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = t * BI_CHUNK;
    if (lo >= n) return;
    size_t hi = lo + BI_CHUNK < n ? lo + BI_CHUNK : n;
    // forward prefix products kept in registers/local memory, then the backward walk
    fr_t pre[BI_CHUNK];
    fr_t acc = fp_one<FrP>();
    for (size_t i = lo; i < hi; i++) {
        pre[i - lo] = acc;                                     // product of the non-zero elements before i
        fr_t x = fr_load(data, i);
        if (!fp_is_zero(x)) acc = fp_mul(acc, x);
    }
    fr_t inv = fr_load(partial_inv, t);                        // 1 / (product of the chunk)
    for (size_t i = hi; i-- > lo;) {
        fr_t x = fr_load(data, i);
        if (fp_is_zero(x)) continue;
        fr_t xi = fp_mul(inv, pre[i - lo]);                    // 1 / x
        inv = fp_mul(inv, x);                                  // drop x from the running inverse
        uint4 a, b;
        fr_to_u4(xi, a, b);
        data[2 * i] = a;
        data[2 * i + 1] = b;
    }
}

// n <= BI_SMALL: one block; each thread inverts its strided share with a private Fermat inversion of the product of
// its share (BI_SMALL / 128 = 16 elements per thread -> 3 + 380/16 products per element; only the recursion tail)
__global__ void __launch_bounds__(128) batch_inv_small_kernel(uint4* __restrict__ data, size_t n) {
    const uint32_t t = threadIdx.x;
    fr_t pre[BI_SMALL / 128];
    fr_t acc = fp_one<FrP>();
    uint32_t cnt = 0;
    for (size_t i = t; i < n; i += 128, cnt++) {
        pre[cnt] = acc;
        fr_t x = fr_load(data, i);
        if (!fp_is_zero(x)) acc = fp_mul(acc, x);
    }
    if (cnt == 0) return;
    fr_t inv = fp_inv(acc);
    for (uint32_t c = cnt; c-- > 0;) {
        size_t i = t + (size_t)c * 128;
        fr_t x = fr_load(data, i);
        if (fp_is_zero(x)) continue;
        fr_t xi = fp_mul(inv, pre[c]);
        inv = fp_mul(inv, x);
        uint4 a, b;
        fr_to_u4(xi, a, b);
        data[2 * i] = a;
        data[2 * i + 1] = b;
    }
}

static b200_error_t batch_inverse_rec(void* d_data, size_t n, cudaStream_t s) {
    if (n == 0) return b200_ok();
    if (n <= BI_SMALL) {
        batch_inv_small_kernel<<<1, 128, 0, s>>>(reinterpret_cast<uint4*>(d_data), n);
        KERNEL_CHECK();
        return b200_ok();
    }
    const size_t nt = (n + BI_CHUNK - 1) / BI_CHUNK;
    DevBuf partial;
    CUDA_TRY(partial.alloc(nt * 32, s));
    batch_inv_up_kernel<<<(unsigned)((nt + 127) / 128), 128, 0, s>>>(partial.as<uint4>(), reinterpret_cast<const uint4*>(d_data), n);
    KERNEL_CHECK();
    B200_TRY(batch_inverse_rec(partial.p, nt, s));           // partial products are never zero
    batch_inv_down_kernel<<<(unsigned)((nt + 127) / 128), 128, 0, s>>>(reinterpret_cast<uint4*>(d_data), partial.as<uint4>(), n);
    KERNEL_CHECK();
    return b200_ok();
}

extern "C" b200_error_t b200_fr_batch_inverse_device(void* d_inout, size_t n, void* stream) {
    B200_TRY(b200_require_device());
    if (n && !d_inout) return b200_err(B200_ERR_INVALID_ARG, "fr_batch_inverse: null pointer");
    return batch_inverse_rec(d_inout, n, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// division by the vanishing polynomial of a domain H (size 2^log_h) on the coset g*K of a domain K (size 2^log_k,
// log_k >= log_h):  v_H(g w_K^i) = g^|H| * w_r^i - 1 with r = |K| / |H| : only r distinct values.
//   kernel 1: the r values inverted (r threads, one Fermat inversion each)
//   kernel 2: evals[i] *= inv[i mod r]
// ---------------------------------------------------------------------------------------------
__global__ void vanishing_inv_table_kernel(uint4* __restrict__ out, fr_t g, fr_t root /* 2^47-th root */, uint32_t log_h,
                                           uint32_t log_r, uint32_t r) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= r) return;
    fr_t gh = g;
    for (uint32_t k = 0; k < log_h; k++) gh = fp_sqr(gh);              // g^(2^log_h)
    fr_t wr = root;
    for (uint32_t k = 0; k < FR_TWO_ADICITY - log_r; k++) wr = fp_sqr(wr);   // primitive r-th root of unity
    fr_t v = fp_sub(fp_mul(gh, fp_pow_u64(wr, j)), fp_one<FrP>());
    v = fp_inv(v);
    uint4 a, b;
    fr_to_u4(v, a, b);
    out[2 * j] = a;
    out[2 * j + 1] = b;
}
__global__ void __launch_bounds__(256) mul_periodic_kernel(uint4* __restrict__ data, const uint4* __restrict__ table, size_t n, uint32_t mask) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_t x = fp_mul(fr_load(data, i), fr_load(table, i & mask));
    uint4 a, b;
    fr_to_u4(x, a, b);
    data[2 * i] = a;
    data[2 * i + 1] = b;
}

extern "C" b200_error_t b200_fr_divide_by_vanishing_on_coset_device(void* d_evals, uint32_t log_k, uint32_t log_h, void* stream) {
    B200_TRY(b200_require_device());
    if (log_k > 28 || log_h > log_k || log_k - log_h > 20) return b200_err(B200_ERR_INVALID_ARG, "divide_by_vanishing: bad domain sizes");
    if (!d_evals) return b200_err(B200_ERR_INVALID_ARG, "divide_by_vanishing: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t log_r = log_k - log_h, r = 1u << log_r;
    DevBuf table;
    CUDA_TRY(table.alloc((size_t)r * 32, s));
    fr_t g, root;
    memcpy(g.v, FR_GENERATOR, sizeof(g.v));
    memcpy(root.v, FR_TWO_ADIC_ROOT, sizeof(root.v));
    vanishing_inv_table_kernel<<<(r + 63) / 64, 64, 0, s>>>(table.as<uint4>(), g, root, log_h, log_r, r);
    KERNEL_CHECK();
    const size_t n = (size_t)1 << log_k;
    mul_periodic_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(reinterpret_cast<uint4*>(d_evals), table.as<uint4>(), n, r - 1);
    KERNEL_CHECK();
    return b200_ok();
}
