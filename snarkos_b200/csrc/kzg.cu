// kzg.cu -- the rest of SURVEY.md 8a rows a8 / a9 and 8f rank 4 around the MSM:
//   Projective::batch_normalization + to_affine, compressed G1Affine encoding        (what a commitment looks like inside a
//       Varuna proof: [UPSTREAM curves/src/templates/short_weierstrass_jacobian/{affine,projective}.rs,
//       utilities/src/serialize/flags.rs])
//   linear combination of polynomials, division by (X - z), KZG10::open               [UPSTREAM algorithms/src/polycommit/
//       kzg10/mod.rs: open / open_with_witness_polynomial / compute_witness_polynomial; sonic_pc/mod.rs: open_combinations]
// All device resident; the host-buffer forms copy in and out on the calling thread's stream.
#include <cstring>

#include "common.cuh"
#include "msm_core.cuh"
#include "ntt_core.cuh"

// ---------------------------------------------------------------------------------------------
// Jacobian -> affine, 16 points per thread sharing ONE inversion (Montgomery's trick over Z; Z = 0 skipped)
// ---------------------------------------------------------------------------------------------
#define NORM_CHUNK 16

struct JacPoint { fq_t X, Y, Z; };
__device__ __forceinline__ JacPoint jac_load(const uint4* in, size_t i) {
    const uint4* p = in + 9 * i;
    JacPoint q;
    q.X = fq_from_u4x3(p);
    q.Y = fq_from_u4x3(p + 3);
    q.Z = fq_from_u4x3(p + 6);
    return q;
}

// mode 0: affine images (x, y Montgomery at 0 / 48, infinity byte at 96, rest of the stride zero)
// mode 1: compressed 48-byte encodings (canonical x little-endian; bit 7 of the last byte: y > -y; bit 6: infinity)
__global__ void __launch_bounds__(64) g1_normalize_kernel(uint8_t* __restrict__ out, const uint4* __restrict__ in, size_t n,
                                                          size_t stride, int mode) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t lo = t * NORM_CHUNK;
    if (lo >= n) return;
    const size_t hi = lo + NORM_CHUNK < n ? lo + NORM_CHUNK : n;
    fq_t pre[NORM_CHUNK];
    fq_t acc = fp_one<FqP>();
    for (size_t i = lo; i < hi; i++) {
        pre[i - lo] = acc;
        const fq_t Z = fq_from_u4x3(in + 9 * i + 6);
        if (!fp_is_zero(Z)) acc = fp_mul(acc, Z);
    }
    fq_t inv = fp_inv_gcd(acc);
    for (size_t i = hi; i-- > lo;) {
        const JacPoint p = jac_load(in, i);
        uint8_t* o = out + i * stride;
        const bool inf = fp_is_zero(p.Z);
        fq_t x = fp_zero<FqP>(), y = fp_zero<FqP>();
        if (!inf) {
            const fq_t zi = fp_mul(inv, pre[i - lo]);                  // 1 / Z
            inv = fp_mul(inv, p.Z);
            const fq_t zi2 = fp_sqr(zi);
            x = fp_mul(p.X, zi2);
            y = fp_mul(p.Y, fp_mul(zi2, zi));
        }
        if (mode == 0) {
            uint32_t* o32 = reinterpret_cast<uint32_t*>(o);            // G1Affine is 8-byte aligned, stride % 8 == 0
#pragma unroll
            for (int k = 0; k < 12; k++) { o32[k] = x.v[k]; o32[12 + k] = y.v[k]; }
            for (size_t k = 24; k < stride / 4; k++) o32[k] = 0;
            o[96] = inf ? 1 : 0;
        } else {
            const fq_t xc = fp_from_mont(x), yc = fp_from_mont(y);
            uint32_t w[12];
#pragma unroll
            for (int k = 0; k < 12; k++) w[k] = xc.v[k];
            if (inf) {
                w[11] |= 1u << 30;
            } else {
                // y > p - y  <=>  2y > p (y != 0 here: the curve has no point of order 2 in the prime-order subgroup,
                // and y = 0 compares equal, flag clear)
                const fq_t ny = fp_neg(yc);                            // p - y on canonical limbs (plain modular negation)
                bool greater = false, decided = false;
#pragma unroll
                for (int k = 11; k >= 0; k--) {
                    if (!decided && yc.v[k] != ny.v[k]) { greater = yc.v[k] > ny.v[k]; decided = true; }
                }
                if (greater) w[11] |= 1u << 31;
            }
            uint32_t* o32 = reinterpret_cast<uint32_t*>(o);            // 48-byte records: 4-byte aligned
#pragma unroll
            for (int k = 0; k < 12; k++) o32[k] = w[k];
        }
    }
}

static b200_error_t g1_normalize_device(void* d_out, const void* d_in, size_t n, size_t stride, int mode, cudaStream_t s) {
    if (n == 0) return b200_ok();
    if (!d_out || !d_in) return b200_err(B200_ERR_INVALID_ARG, "g1_normalize: null pointer");
    if (mode == 0 && (stride < 104 || (stride & 7))) return b200_err(B200_ERR_INVALID_ARG, "g1_normalize: affine stride must be >= 104 and 8-byte aligned");
    if (reinterpret_cast<uintptr_t>(d_in) & 15) return b200_err(B200_ERR_INVALID_ARG, "g1_normalize: input must be 16-byte aligned on the device");
    const size_t nt = (n + NORM_CHUNK - 1) / NORM_CHUNK;
    g1_normalize_kernel<<<(unsigned)((nt + 63) / 64), 64, 0, s>>>(reinterpret_cast<uint8_t*>(d_out), reinterpret_cast<const uint4*>(d_in), n, stride, mode);
    KERNEL_CHECK();
    return b200_ok();
}

extern "C" b200_error_t b200_g1_batch_normalize_device(void* d_out_affine, const void* d_in_jacobian, size_t count, size_t affine_stride, void* stream) {
    B200_TRY(b200_require_device());
    return g1_normalize_device(d_out_affine, d_in_jacobian, count, affine_stride, 0, (cudaStream_t)stream);
}
extern "C" b200_error_t b200_g1_compress_device(void* d_out_48B, const void* d_in_jacobian, size_t count, void* stream) {
    B200_TRY(b200_require_device());
    return g1_normalize_device(d_out_48B, d_in_jacobian, count, 48, 1, (cudaStream_t)stream);
}

static b200_error_t g1_normalize_host(void* out, const void* in, size_t n, size_t stride, int mode) {
    B200_TRY(b200_require_device());
    if (n == 0) return b200_ok();
    if (!out || !in) return b200_err(B200_ERR_INVALID_ARG, "g1_normalize: null pointer");
    cudaStream_t s = b200_thread_stream();
    DevBuf d_in, d_out;
    CUDA_TRY(d_in.alloc(n * 144, s));
    CUDA_TRY(d_out.alloc(n * stride, s));
    B200_TRY(b200_h2d(d_in.p, in, n * 144, s));
    B200_TRY(g1_normalize_device(d_out.p, d_in.p, n, stride, mode, s));
    B200_TRY(b200_d2h(out, d_out.p, n * stride, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return b200_ok();
}
extern "C" b200_error_t b200_g1_batch_normalize(void* out_affine, const void* in_jacobian, size_t count, size_t affine_stride) {
    return g1_normalize_host(out_affine, in_jacobian, count, affine_stride, 0);
}
extern "C" b200_error_t b200_g1_compress(void* out_48B, const void* in_jacobian, size_t count) {
    return g1_normalize_host(out_48B, in_jacobian, count, 48, 1);
}

// ---------------------------------------------------------------------------------------------
// out[i] = sum_m c[m] * p_m[i]   (p_m = polys[off[m] .. off[m + 1]), zero beyond its length)
// the linear combination KZG10::open / SonicKZG10::open_combinations form with the opening challenges before ONE
// witness division and ONE MSM
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fr_lincomb_kernel(uint4* __restrict__ out, const uint4* __restrict__ polys,
                                                         const unsigned long long* __restrict__ off, uint32_t k,
                                                         const uint4* __restrict__ coeffs, size_t out_len) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= out_len) return;
    fr_t acc = fp_zero<FrP>();
    for (uint32_t m = 0; m < k; m++) {
        const unsigned long long lo = off[m], len = off[m + 1] - lo;
        if (i < len) acc = fp_add(acc, fp_mul(fr_load(coeffs, m), fr_load(polys, lo + i)));
    }
    uint4 a, b;
    fr_to_u4(acc, a, b);
    out[2 * i] = a;
    out[2 * i + 1] = b;
}

extern "C" b200_error_t b200_fr_linear_combination_device(void* d_out, const void* d_polys, const uint64_t* offsets_host, size_t k,
                                                          const void* d_coeffs_mont, size_t out_len, void* stream) {
    B200_TRY(b200_require_device());
    if (out_len == 0) return b200_ok();
    if (!d_out || !offsets_host || (k && (!d_polys || !d_coeffs_mont))) return b200_err(B200_ERR_INVALID_ARG, "fr_linear_combination: null pointer");
    if (k > 65536) return b200_err(B200_ERR_TOO_LARGE, "fr_linear_combination: more than 65536 polynomials");
    for (size_t m = 0; m < k; m++)
        if (offsets_host[m + 1] < offsets_host[m]) return b200_err(B200_ERR_INVALID_ARG, "fr_linear_combination: offsets must be non-decreasing");
    cudaStream_t s = (cudaStream_t)stream;
    DevBuf d_off;
    CUDA_TRY(d_off.alloc((k + 1) * 8, s));
    CUDA_TRY(cudaMemcpyAsync(d_off.p, offsets_host, (k + 1) * 8, cudaMemcpyHostToDevice, s));
    fr_lincomb_kernel<<<(unsigned)((out_len + 255) / 256), 256, 0, s>>>(reinterpret_cast<uint4*>(d_out), reinterpret_cast<const uint4*>(d_polys),
                                                                        d_off.as<unsigned long long>(), (uint32_t)k,
                                                                        reinterpret_cast<const uint4*>(d_coeffs_mont), out_len);
    KERNEL_CHECK();
    return b200_ok();
}

// ---------------------------------------------------------------------------------------------
// Division by (X - z): with H_j = sum_{i >= j} p_i z^(i - j) (the suffix Horner values) the quotient is q_j = H_{j+1}
// and the remainder H_0 = p(z).  H obeys H_j = p_j + z H_{j+1} -- a linear recurrence, parallelised in chunks of
// DIV_CHUNK: (1) every thread folds its chunk into S_t = sum_{i in chunk} p_i z^(i - lo); (2) the chunk sums are the
// coefficients of the SAME problem at the point z^DIV_CHUNK (recursion, n / 64 elements); (3) every thread walks its
// chunk downwards from the carry of the chunk above.  2 products per coefficient (+ 1/64 of that recursively).
// ---------------------------------------------------------------------------------------------
#define DIV_CHUNK 64

__global__ void __launch_bounds__(128) horner_chunk_sums_kernel(uint4* __restrict__ sums, uint4* __restrict__ z_pow, const uint4* __restrict__ p,
                                                                size_t n, const uint4* __restrict__ zp) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t lo = t * DIV_CHUNK;
    if (lo >= n) return;
    const size_t hi = lo + DIV_CHUNK < n ? lo + DIV_CHUNK : n;
    const fr_t z = fr_load(zp, 0);
    fr_t acc = fp_zero<FrP>();
    for (size_t i = hi; i-- > lo;) acc = fp_add(fp_mul(acc, z), fr_load(p, i));
    uint4 a, b;
    fr_to_u4(acc, a, b);
    sums[2 * t] = a;
    sums[2 * t + 1] = b;
    if (t == 0) {                                                      // z^DIV_CHUNK for the next level
        fr_t w = z;
#pragma unroll 1
        for (int k = 1; k < DIV_CHUNK; k <<= 1) w = fp_sqr(w);
        fr_to_u4(w, a, b);
        z_pow[0] = a;
        z_pow[1] = b;
    }
}

// H_j for the chunk of thread t, from carry = H_{hi} (0 above the top); written to out[j - shift] for j >= shift, and
// H_0 to rem when shift == 1 (quotient form).  out may alias p when shift == 0 is NOT used by callers (in place unsafe).
__global__ void __launch_bounds__(128) horner_walk_kernel(uint4* __restrict__ out, uint4* __restrict__ rem, const uint4* __restrict__ p, size_t n,
                                                          const uint4* __restrict__ zp, const uint4* __restrict__ carry_in, uint32_t shift) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t lo = t * DIV_CHUNK;
    if (lo >= n) return;
    const size_t hi = lo + DIV_CHUNK < n ? lo + DIV_CHUNK : n;
    const fr_t z = fr_load(zp, 0);
    fr_t acc = (carry_in && hi < n) ? fr_load(carry_in, t + 1) : fp_zero<FrP>();
    for (size_t j = hi; j-- > lo;) {
        acc = fp_add(fp_mul(acc, z), fr_load(p, j));
        uint4 a, b;
        fr_to_u4(acc, a, b);
        if (j >= shift) {
            out[2 * (j - shift)] = a;
            out[2 * (j - shift) + 1] = b;
        } else if (rem) {
            rem[0] = a;
            rem[1] = b;
        }
    }
}

// d_H[j] (shift 0) or quotient / remainder (shift 1) of p at the point *d_z
static b200_error_t horner_suffix(uint4* d_out, uint4* d_rem, const uint4* d_p, size_t n, const uint4* d_z, uint32_t shift, cudaStream_t s) {
    const size_t nt = (n + DIV_CHUNK - 1) / DIV_CHUNK;
    if (nt <= 1) {
        horner_walk_kernel<<<1, 32, 0, s>>>(d_out, d_rem, d_p, n, d_z, nullptr, shift);
        KERNEL_CHECK();
        return b200_ok();
    }
    DevBuf sums, carry, zpow;
    CUDA_TRY(sums.alloc(nt * 32, s));
    CUDA_TRY(carry.alloc(nt * 32, s));
    CUDA_TRY(zpow.alloc(32, s));
    horner_chunk_sums_kernel<<<(unsigned)((nt + 127) / 128), 128, 0, s>>>(sums.as<uint4>(), zpow.as<uint4>(), d_p, n, d_z);
    KERNEL_CHECK();
    B200_TRY(horner_suffix(carry.as<uint4>(), nullptr, sums.as<uint4>(), nt, zpow.as<uint4>(), 0, s));
    horner_walk_kernel<<<(unsigned)((nt + 127) / 128), 128, 0, s>>>(d_out, d_rem, d_p, n, d_z, carry.as<uint4>(), shift);
    KERNEL_CHECK();
    return b200_ok();
}

extern "C" b200_error_t b200_fr_divide_by_linear_device(void* d_quotient, const void* d_poly, size_t n, const void* d_point_mont,
                                                        void* d_out_remainder, void* stream) {
    B200_TRY(b200_require_device());
    if (!d_point_mont) return b200_err(B200_ERR_INVALID_ARG, "fr_divide_by_linear: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) {
        if (d_out_remainder) CUDA_TRY(cudaMemsetAsync(d_out_remainder, 0, 32, s));
        return b200_ok();
    }
    if (!d_poly || (n > 1 && !d_quotient)) return b200_err(B200_ERR_INVALID_ARG, "fr_divide_by_linear: null pointer");
    if (d_quotient == d_poly) return b200_err(B200_ERR_INVALID_ARG, "fr_divide_by_linear: quotient must not alias the polynomial");
    return horner_suffix(reinterpret_cast<uint4*>(d_quotient), reinterpret_cast<uint4*>(d_out_remainder),
                         reinterpret_cast<const uint4*>(d_poly), n, reinterpret_cast<const uint4*>(d_point_mont), 1, s);
}

// ---------------------------------------------------------------------------------------------
// KZG10::open: witness polynomial (p(X) - p(z)) / (X - z) -- the quotient of p by (X - z), the remainder being p(z) --
// committed against the resident powers.  The hiding part of snarkVM's open (the same division on the blinding
// polynomial against powers_of_beta_times_gamma_g, and its evaluation random_v) is this call once more.
// ---------------------------------------------------------------------------------------------
extern "C" b200_error_t b200_kzg_open_device(void* d_out, uint64_t handle, const void* d_coeffs_mont, size_t n, const void* d_point_mont,
                                             void* d_out_eval_mont, void* stream) {
    B200_TRY(b200_require_device());
    if (!d_out || !d_point_mont || (n && !d_coeffs_mont)) return b200_err(B200_ERR_INVALID_ARG, "kzg_open: null pointer");
    RegisteredBases rb;
    B200_TRY(b200_lookup_bases(handle, &rb));
    if (n > rb.n + 1) return b200_err(B200_ERR_INVALID_ARG, "kzg_open: witness polynomial longer than the registered powers");
    cudaStream_t s = (cudaStream_t)stream;
    DevBuf q;
    CUDA_TRY(q.alloc((n > 1 ? n - 1 : 1) * 32, s));
    B200_TRY(b200_fr_divide_by_linear_device(q.p, d_coeffs_mont, n, d_point_mont, d_out_eval_mont, s));
    return b200_kzg_commit_device(d_out, handle, q.p, n > 1 ? n - 1 : 0, s);
}

extern "C" b200_error_t b200_kzg_open(void* out, uint64_t handle, const void* coeffs_mont, size_t n, const void* point_mont, void* out_eval_mont) {
    B200_TRY(b200_require_device());
    if (!out || !point_mont || (n && !coeffs_mont)) return b200_err(B200_ERR_INVALID_ARG, "kzg_open: null pointer");
    cudaStream_t s = b200_thread_stream();
    DevBuf d_c, d_z, d_out;
    CUDA_TRY(d_c.alloc(n * 32, s));
    CUDA_TRY(d_z.alloc(64, s));                                         // point | evaluation
    CUDA_TRY(d_out.alloc(144, s));
    if (n) B200_TRY(b200_h2d(d_c.p, coeffs_mont, n * 32, s));
    CUDA_TRY(cudaMemcpyAsync(d_z.p, point_mont, 32, cudaMemcpyHostToDevice, s));
    B200_TRY(b200_kzg_open_device(d_out.p, handle, d_c.p, n, d_z.p, d_z.as<uint8_t>() + 32, s));
    CUDA_TRY(cudaMemcpyAsync(out, d_out.p, 144, cudaMemcpyDeviceToHost, s));
    if (out_eval_mont) CUDA_TRY(cudaMemcpyAsync(out_eval_mont, d_z.as<uint8_t>() + 32, 32, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return b200_ok();
}
