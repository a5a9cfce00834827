/*
 * snarkos_b200.h -- C ABI of the B200-native BLS12-377 MSM / Fr-NTT hot path.
 *
 * Drop-in boundary for the path snarkOS reaches through snarkVM:
 *
 *   snarkvm-algorithms  msm::VariableBase::msm(bases, scalars)            -> b200_msm_g1_bls12_377
 *       [UPSTREAM algorithms/src/msm/variable_base/mod.rs; reached from snarkOS at
 *        node/src/validator/mod.rs:383-391 (prove), node/bft/ledger-service/src/ledger.rs:341,346 (verify)]
 *   snarkvm-algorithms  fft::EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place
 *                                                                         -> b200_ntt_fr_bls12_377
 *       [UPSTREAM algorithms/src/fft/domain.rs: in_order_fft_in_place / in_order_ifft_in_place /
 *        in_order_coset_fft_in_place / in_order_coset_ifft_in_place]
 *
 * The hook precedent is snarkVM's own optional `cuda` feature (snarkvm-algorithms-cuda: extern "C"
 * snarkvm_msm + NTT entry points behind a type-id / size guard); this library replaces that backend
 * without reusing it.  A Rust -sys binding is shown in INTEGRATION.md.
 *
 * Data conventions (identical to snarkVM's in-memory layout, little-endian u64 limbs):
 *   Fr element      32 B, Montgomery form (value * 2^256 mod r), fully reduced
 *   Fq element      48 B, Montgomery form (value * 2^384 mod p), fully reduced
 *   G1Affine        `affine_stride` bytes (104 for rustc's layout): x @0, y @48, infinity flag byte @96
 *   scalar          32 B BigInteger256, canonical (NOT Montgomery), < r (a non-canonical value k >= r is reduced mod r
 *                   on the device: the result is still k * P)
 *   G1Projective    144 B Jacobian (X, Y, Z) Montgomery; Z = 0 <=> infinity (returned as (1, 1, 0))
 *
 * Error model: every call returns b200_error_t; code == 0 is success, code > 0 is a cudaError_t,
 * code < 0 is a library error.  msg points to a static string.  Nothing throws across the ABI.
 * There is NO CPU fallback: without a CUDA device every compute call fails with a non-zero code.
 *
 * Threading: all entry points are re-entrant; concurrent callers (rayon workers, tokio blocking
 * threads) each run on their own CUDA stream.
 */
#ifndef SNARKOS_B200_H
#define SNARKOS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int32_t code;
    const char* msg;
} b200_error_t;

enum {
    B200_OK = 0,
    B200_ERR_INVALID_ARG = -1,
    B200_ERR_NO_DEVICE = -2,
    B200_ERR_TOO_LARGE = -3,
    B200_ERR_BAD_HANDLE = -4,
    B200_ERR_NOT_INITIALIZED = -5
};

enum { B200_NTT_FORWARD = 0, B200_NTT_INVERSE = 1 };

/* ---- lifecycle ------------------------------------------------------------------------------- */
/* Bind the calling process to CUDA device `device` (one process per GPU) and build the static
 * tables.  Idempotent.  b200_shutdown releases every cached table and registered base set. */
b200_error_t b200_init(int device);
void b200_shutdown(void);
/* Hands every cached scratch block (per-stream caches of the calls' temporaries, up to half the device memory in total) back to the
 * driver's stream-ordered memory pool of the device: for a long-running node between proving bursts, or before calls of
 * a very different shape.  The pool keeps the memory mapped (cudaMemPoolTrimTo on the device's default pool returns it
 * to the OS; the next large call then pays for mapping it again).  Blocks cached on the library's own streams go back in
 * their stream's order; blocks cached on a CALLER's stream (the _device entry points) are freed on a stream of the
 * library, because the caller's stream may be gone -- the driver then reuses them across streams, which made a 2^25-point
 * call on such a stream 25 % slower for the following calls (185 vs 147 ms).  Never required for correctness: an
 * allocation that fails does the same before it retries. */
b200_error_t b200_release_scratch(void);
/* ABI version of this header (for the -sys crate's build-time check). */
uint32_t b200_abi_version(void);
/* Tuning knobs.  The B200_* environment variables are read once, on first use; this call changes a knob afterwards
 * (sweep tools, tests, the -sys crate's init).  Keys: msm_window_bits, msm_glv, msm_affine_rounds, msm_slices,
 * msm_chunk, msm_seg_len, msm_reduce_quad_max, msm_host_pipeline, msm_host_first_log, msm_host_chunk_log, msm_stream_two, msm_auto_table, msm_list_budget_bytes,
 * msm_queue_threshold, msm_queue_linger_us, ntt_plan ("a,b,c"), ntt_tile_log, ntt_radix4, ntt_boundary_tables, ntt_host_pipeline, ntt_variant,
 * staged_copies.  Unknown key -> B200_ERR_INVALID_ARG. */
b200_error_t b200_set_option(const char* key, const char* value);
/* Observable fallbacks / activity: kernel_launches, msm_xyzz_fallbacks (calls whose pair-round lists did not fit in HBM
 * and ran the XYZZ-only accumulation), queue_submits, queue_batches, streams_created (per-thread
 * CUDA streams alive: bounded by the number of threads calling at the same time, exited threads hand theirs back). */
b200_error_t b200_get_counter(const char* name, uint64_t* out);

/* ---- VariableBase::msm ------------------------------------------------------------------------
 * out_jacobian_144B = sum_i scalars[i] * points[i].  Host buffers; copies are done internally.    */
b200_error_t b200_msm_g1_bls12_377(void* out_jacobian_144B, const void* points, size_t npoints,
                                   const void* scalars_32B_canonical, size_t affine_stride);
/* Same, all pointers are DEVICE pointers; work is enqueued on `stream` (cudaStream_t, may be NULL)
 * and the call returns without synchronising. */
b200_error_t b200_msm_g1_bls12_377_device(void* d_out_jacobian_144B, const void* d_points, size_t npoints,
                                          const void* d_scalars, size_t affine_stride, void* stream);

/* Batched small MSMs: nmsm independent sums over consecutive ranges of one (points, scalars)
 * pair, MSM m covering points [offsets[m], offsets[m + 1]) (offsets[0] = 0, non-decreasing, nmsm + 1 entries).  One
 * launch set for all of them: the batch verifier's per-transaction linear combinations (KZG10::batch_check
 * [UPSTREAM algorithms/src/polycommit/kzg10/mod.rs]) aggregated across a block instead of hundreds of launch-bound
 * calls.  out receives nmsm Jacobian points (144 B each); an empty range yields infinity. */
b200_error_t b200_msm_batch_g1_bls12_377(void* out_jacobian, const void* points, const void* scalars,
                                         const uint64_t* offsets, size_t nmsm, size_t affine_stride);
b200_error_t b200_msm_batch_g1_bls12_377_device(void* d_out_jacobian, const void* d_points, const void* d_scalars,
                                                const void* d_offsets_u64, size_t nmsm, size_t npoints,
                                                size_t affine_stride, void* stream);

/* Coalescing queue for concurrent SMALL MSMs (BASELINE configs[3]: a validator verifies the transactions of a block on
 * up to 512 threads -- /root/reference/cli/src/commands/start.rs:623-640, node/bft/ledger-service/src/ledger.rs:341-347 --
 * and each verification ends in VariableBase::msm calls over tens of points).  b200_msm_submit enqueues one MSM and
 * returns a ticket at once; a dispatcher thread packs everything pending into ONE segmented launch set; b200_msm_wait
 * blocks until that MSM is done and copies its 144-byte result.  points / scalars must stay valid until the wait
 * returns; a ticket is waited on exactly once.  No lock is held while the GPU works.  With option
 * msm_queue_threshold = T, b200_msm_g1_bls12_377 routes every call of <= T points through this queue by itself, so the
 * Rust call sites need no change. */
b200_error_t b200_msm_submit(const void* points, size_t npoints, const void* scalars_32B_canonical, size_t affine_stride,
                             uint64_t* out_ticket);
b200_error_t b200_msm_wait(uint64_t ticket, void* out_jacobian_144B);

/* Resident bases (the SRS `powers_of_beta_g` is fixed for the process lifetime -- KZG10::commit
 * [UPSTREAM algorithms/src/polycommit/kzg10/mod.rs] calls msm on a prefix of it every time).  Sets of 2^10 .. 2^20
 * points are stored with their window table (see b200_msm_register_bases_tabulated: 16 x the memory, no window fold,
 * half the latency of a small commit); larger sets as packed points only. */
b200_error_t b200_msm_register_bases(const void* points, size_t npoints, size_t affine_stride,
                                     uint64_t* out_handle);
b200_error_t b200_msm_register_bases_device(const void* d_points, size_t npoints, size_t affine_stride,
                                            void* stream, uint64_t* out_handle);
/* Same, and additionally tabulates 2^(c*w) * P_i for every window w in HBM (nwin * npoints * 128 bytes: 27.9 GB for
 * a 2^24-point SRS at the automatic c = 20 -- a B200 has 180 GB): MSMs against the handle then share ONE bucket set across windows,
 * skip the window fold and reduce 2^(c-1) buckets once.  window_bits 0 = automatic (16..24).  Registration costs about
 * 2900 Fq products per point, once. */
b200_error_t b200_msm_register_bases_tabulated(const void* points, size_t npoints, size_t affine_stride,
                                               uint32_t window_bits, uint64_t* out_handle);
b200_error_t b200_msm_register_bases_tabulated_device(const void* d_points, size_t npoints, size_t affine_stride,
                                                      uint32_t window_bits, void* stream, uint64_t* out_handle);
/* MSM over the first `nscalars` registered bases (nscalars <= registered count). */
b200_error_t b200_msm_registered(void* out_jacobian_144B, uint64_t handle, const void* scalars,
                                 size_t nscalars);
b200_error_t b200_msm_registered_device(void* d_out_jacobian_144B, uint64_t handle, const void* d_scalars,
                                        size_t nscalars, void* stream);
b200_error_t b200_msm_release_bases(uint64_t handle);

/* KZG10::commit / commit_lagrange against resident powers [UPSTREAM algorithms/src/polycommit/kzg10/mod.rs]:
 * out = sum_i coeffs[i] * powers[i] for coefficients (or Lagrange evaluations) given as Montgomery Fr elements,
 * exactly what snarkVM holds in a DensePolynomial; the to_bigint conversion runs on the device.  A hiding commitment is
 * this call on powers_of_beta_g plus this call on powers_of_beta_times_gamma_g with the blinding polynomial, added. */
b200_error_t b200_kzg_commit(void* out_jacobian_144B, uint64_t handle, const void* coeffs_32B_mont, size_t ncoeffs);
b200_error_t b200_kzg_commit_device(void* d_out_jacobian_144B, uint64_t handle, const void* d_coeffs_mont,
                                    size_t ncoeffs, void* stream);

/* k commitments against ONE resident set in one launch set (a Varuna round commits its polynomials together:
 * SonicKZG10::commit over a slice of labeled polynomials [UPSTREAM algorithms/src/polycommit/sonic_pc/mod.rs]).
 * Polynomial m = coeffs[offsets[m] .. offsets[m + 1]) (Montgomery Fr, concatenated; offsets[0] = 0, k + 1 entries, a HOST
 * array in both forms) multiplies powers 0 .. len_m - 1.  out receives k Jacobian points. */
b200_error_t b200_kzg_commit_batch(void* out_jacobian, uint64_t handle, const void* coeffs_32B_mont,
                                   const uint64_t* offsets, size_t k);
b200_error_t b200_kzg_commit_batch_device(void* d_out_jacobian, uint64_t handle, const void* d_coeffs_mont,
                                          const uint64_t* offsets_host, size_t k, void* stream);

/* KZG10::open [UPSTREAM algorithms/src/polycommit/kzg10/mod.rs: open -> compute_witness_polynomial ->
 * open_with_witness_polynomial]: out = commitment to the witness polynomial (p(X) - p(z)) / (X - z) against the
 * resident powers, *out_eval = p(z) (may be NULL).  Montgomery Fr coefficients and point.  The hiding part of
 * snarkVM's open is the same call on the blinding polynomial against powers_of_beta_times_gamma_g (its evaluation is
 * random_v), added by the caller.  SonicKZG10::open_combinations first forms ONE polynomial with
 * b200_fr_linear_combination_device. */
b200_error_t b200_kzg_open(void* out_jacobian_144B, uint64_t handle, const void* coeffs_32B_mont, size_t ncoeffs,
                           const void* point_32B_mont, void* out_eval_32B_mont);
b200_error_t b200_kzg_open_device(void* d_out_jacobian_144B, uint64_t handle, const void* d_coeffs_mont, size_t ncoeffs,
                                  const void* d_point_mont, void* d_out_eval_mont, void* stream);

/* Window width the library would pick for `npoints` (0 = library default); B200_MSM_C overrides. */
uint32_t b200_msm_window_bits(size_t npoints);
/* number of batched-affine pair rounds (snarkVM batched::batch_add counterpart) a call of this size runs before the
 * XYZZ finish, for the window width above (diagnostic: bench.py's roofline accounting) */
uint32_t b200_msm_affine_rounds(size_t npoints);
/* what b200_msm_g1_bls12_377 runs with for npoints points: out4 = {window bits, windows, pair rounds, GLV split (0/1)} */
void b200_msm_describe(size_t npoints, uint32_t* out4);

/* Sum of `count` Jacobian points (144 B each): the multi-GPU partial-sum combine.  Device pointers. */
b200_error_t b200_g1_sum_jacobian_device(void* d_out_jacobian_144B, const void* d_in, size_t count,
                                         void* stream);

/* ---- output forms of a commitment ----------------------------------------------
 * Projective::batch_normalization + to_affine [UPSTREAM curves/src/templates/short_weierstrass_jacobian/projective.rs]:
 * `count` Jacobian images (144 B) -> G1Affine images of `affine_stride` bytes (x, y Montgomery, infinity flag byte at
 * 96, padding zero), one shared inversion per 16 points. */
b200_error_t b200_g1_batch_normalize(void* out_affine, const void* in_jacobian, size_t count, size_t affine_stride);
b200_error_t b200_g1_batch_normalize_device(void* d_out_affine, const void* d_in_jacobian, size_t count,
                                            size_t affine_stride, void* stream);
/* Compressed serialisation of the same points, 48 bytes each -- what `ToBytes::write_le` / `serialize_compressed` emit
 * for a G1Affine inside a Varuna proof [UPSTREAM .../short_weierstrass_jacobian/affine.rs, utilities/src/serialize/
 * flags.rs]: canonical x little-endian; bit 7 of byte 47 set iff y > -y; bit 6 = infinity (x = 0). */
b200_error_t b200_g1_compress(void* out_48B, const void* in_jacobian, size_t count);
b200_error_t b200_g1_compress_device(void* d_out_48B, const void* d_in_jacobian, size_t count, void* stream);

/* ---- EvaluationDomain (i)FFT -------------------------------------------------------------------
 * In-place transform of `batch` polynomials of 2^log_n Montgomery Fr elements, natural order in and
 * out; polynomial b starts at element b * batch_stride_elems.  direction: B200_NTT_FORWARD /
 * B200_NTT_INVERSE (inverse includes the n^-1 scaling).  coset != 0 selects coset_fft_in_place
 * (input j scaled by 22^j first) / coset_ifft_in_place (output j scaled by 22^-j last).
 * The caller has already zero-padded to the domain size, exactly like EvaluationDomain does. */
b200_error_t b200_ntt_fr_bls12_377(void* inout_32B_mont, uint32_t log_n, size_t batch,
                                   size_t batch_stride_elems, int direction, int coset);
b200_error_t b200_ntt_fr_bls12_377_device(void* d_inout, uint32_t log_n, size_t batch,
                                          size_t batch_stride_elems, int direction, int coset, void* stream);

/* ---- polynomial glue between (i)FFTs, device resident (Varuna prover rounds) -------------
 * Montgomery Fr vectors.  op 0: out = a*b  1: a+b  2: a-b  3: a*b + c  4: a*b - c; b is a vector or (b_is_scalar) one
 * broadcast element; out may alias any input.  [UPSTREAM algorithms/src/fft/evaluations.rs: Mul/Add/Sub impls] */
b200_error_t b200_fr_vec_op_device(int op, void* d_out, const void* d_a, const void* d_b, const void* d_c, size_t n,
                                   int b_is_scalar, void* stream);
/* In-place batch inversion (Montgomery's trick); zero elements stay zero like snarkVM's batch_inversion
 * [UPSTREAM fields/src/traits/field.rs]. */
b200_error_t b200_fr_batch_inverse_device(void* d_inout, size_t n, void* stream);
/* evals[i] /= v_H(22 * w_K^i) for i < 2^log_k: division by the vanishing polynomial of a size-2^log_h domain on the
 * coset of a size-2^log_k domain [UPSTREAM algorithms/src/fft/domain.rs: divide_by_vanishing_poly_on_coset_in_place]. */
b200_error_t b200_fr_divide_by_vanishing_on_coset_device(void* d_evals, uint32_t log_k, uint32_t log_h, void* stream);

/* out[i] = sum_m coeffs[m] * p_m[i] for i < out_len, p_m = polys[offsets[m] .. offsets[m + 1]) (zero beyond its length):
 * the combination by opening challenges that KZG10::open / SonicKZG10::open_combinations form before one witness
 * division and one MSM.  offsets: HOST array of k + 1 entries; polys, coeffs (k Montgomery Fr), out: device. */
b200_error_t b200_fr_linear_combination_device(void* d_out, const void* d_polys, const uint64_t* offsets_host, size_t k,
                                               const void* d_coeffs_mont, size_t out_len, void* stream);
/* Division of the degree n - 1 polynomial p by (X - z): quotient (n - 1 coefficients, must not alias p) and remainder
 * p(z) (may be NULL) [UPSTREAM algorithms/src/fft/polynomial/dense.rs: Div by a linear divisor / evaluate]. */
b200_error_t b200_fr_divide_by_linear_device(void* d_quotient, const void* d_poly, size_t n, const void* d_point_mont,
                                             void* d_out_remainder, void* stream);

/* Multi-GPU four-step building block: scales a block of a distributed polynomial of total size 2^log_n.
 * kind 0: element (r, c) of the row-major [rows x cols] block times w_N^((row_base + r) * (col_base + c)) -- the
 * twiddle between the two transform axes; kind 1: element i times 22^(row_base + i) (direction 0) or
 * 22^-(row_base + i) (direction 1) -- the coset shift.  w_N is the inverse root for direction 1. */
b200_error_t b200_fr_mul_powers_device(void* d_data, uint32_t log_n, int direction, int kind,
                                       unsigned long long rows, unsigned long long cols,
                                       unsigned long long row_base, unsigned long long col_base, void* stream);

/* Fused exchange of the four-step NTT ("fuse twiddle multiply into the pack kernel", all-to-all over
 * NVLink): the local slab [r_local x c] (row-major, 32 B Montgomery Fr) of a matrix distributed by rows over `world`
 * ranks is transposed, optionally multiplied by w_N^((row_base + r) * col), and every element is written straight to
 * its place in the [c / world x r_local * world] slab of the rank that owns its column.  dst_ptrs: HOST array of `world`
 * device pointers, dst_ptrs[d] = that rank's destination slab (this rank's own buffer for d == rank, peer memory opened
 * with b200_peer_buffer_open otherwise).  The caller separates writers and readers of a buffer with a barrier. */
b200_error_t b200_fr_exchange_transpose_device(const void* d_src, void* const* dst_ptrs, uint32_t world, uint32_t rank,
                                               unsigned long long r_local, unsigned long long c, uint32_t log_n,
                                               int direction, int twiddle, unsigned long long row_base, void* stream);
/* The same for a PART of the slab: rows [row_off, row_off + r_count) of the r_local rows (d_src_slab points at row 0)
 * and, for every destination rank, columns [col_lo, col_lo + col_cnt) of that rank's c / world columns -- a slab can be
 * exchanged in chunks while the transforms of the chunks that have arrived are already running.  cta_limit != 0
 * bounds the grid (the kernel strides over its tiles) so that a concurrently running transform keeps its SM slots. */
b200_error_t b200_fr_exchange_transpose_part_device(const void* d_src_slab, void* const* dst_ptrs, uint32_t world, uint32_t rank,
                                                    unsigned long long r_local, unsigned long long row_off,
                                                    unsigned long long r_count, unsigned long long c,
                                                    unsigned long long col_lo, unsigned long long col_cnt, uint32_t log_n,
                                                    int direction, int twiddle, unsigned long long row_base,
                                                    uint32_t cta_limit, void* stream);
/* The local transform and the exchange in ONE launch set: `rows` rows of 2^log_len elements at d_rows (row-major, left
 * unchanged) are transformed (direction as in b200_ntt_fr_bls12_377; the inverse scales by 2^-log_len) and the LAST pass
 * of the transform stores every output (row, col) -- times w_N^((row_base + row) * col), N = 2^log_n_total, when
 * `twiddle` -- straight to its place in the transposed slab of the rank that owns column col, exactly where
 * b200_fr_exchange_transpose_device would put it: no local slab is written and read again between the butterflies and
 * the peer stores.  Tiles of that pass hold up to 16 adjacent rows, so every column leaves as one run of up to 512 B. */
b200_error_t b200_ntt_rows_exchange_device(const void* d_rows, void* const* dst_ptrs, uint32_t world, uint32_t rank,
                                           unsigned long long rows, uint32_t log_len, uint32_t log_n_total, int direction,
                                           int twiddle, unsigned long long row_base, void* stream);
/* Exchange buffers: cudaMalloc memory with its 64-byte CUDA IPC handle (sent to the other ranks by the caller), the
 * mapping of a peer's buffer into this process, and their release. */
b200_error_t b200_peer_buffer_alloc(size_t bytes, void** d_ptr, void* handle64);
b200_error_t b200_peer_buffer_open(const void* handle64, void** d_ptr);
b200_error_t b200_peer_buffer_close(void* d_ptr);
b200_error_t b200_peer_buffer_free(void* d_ptr);

/* ---- synthetic inputs & diagnostics (used by bench.py / tests; not on the snarkVM call path) ---- */
/* points[i] = k_i * G with k_i = splitmix64(seed, i), written as G1Affine images (Montgomery). */
b200_error_t b200_g1_synthetic_bases_device(void* d_out_points, size_t npoints, size_t affine_stride,
                                            uint64_t seed, void* stream);
/* Element-wise device arithmetic on host buffers, for parity tests of the primitives.
 * op: 0 fr_mul 1 fr_add 2 fr_sub 3 fq_mul 4 fq_add 5 fq_sub 6 fq_inv 7 fr_inv */
b200_error_t b200_debug_field_op(int op, void* out, const void* a, const void* b, size_t n);
/* Curve primitives: op 0: out[i] = a[i] + b[i] (affine inputs, XYZZ madd), 1: out[i] = 2 a[i],
 * 2: out[i] = k[i] * a[i] with 64-bit k in b.  out = Jacobian 144 B each. */
b200_error_t b200_debug_g1_op(int op, void* out_jacobian, const void* a_affine, const void* b,
                              size_t n, size_t affine_stride);
/* Throughput microbenchmarks (roofline denominators of the integer-multiply pipe).  kind: 0 IMAD, 1 IMAD.WIDE with
 * addend (ptxas splits it), 2 IMAD.HI, 6 IADD3, 7 IMAD.WIDE without addend, 8 IMAD.WIDE.X carry chains, 11 DFMA,
 * 14 IMAD.WIDE with addend and no chain, 3/4 Fr/Fq Montgomery product, 5 XYZZ mixed add.  Runs `iters`
 * dependent operations per thread on a full-chip grid; reports elapsed milliseconds and operations executed. */
b200_error_t b200_debug_microbench(int kind, uint32_t iters, float* out_ms, double* out_ops);
/* Per-stage device timing of the calls the CALLING THREAD makes between begin and end (CUDA events on the
 * launching stream).  b200_profile_end synchronises the device and writes "stage=ms;stage=ms;..." into buf. */
void b200_profile_begin(void);
b200_error_t b200_profile_end(char* buf, size_t buflen);
/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
uint64_t b200_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SNARKOS_B200_H */
