//! Rust binding of include/snarkos_b200.h (ABI version 2).  Generic over the snarkVM types only through their memory
//! layout, like upstream's snarkvm-algorithms-cuda: the caller passes slices, this crate passes pointers and
//! `size_of::<T>()`.  Every `extern "C"` function of the header is declared in [`sys`]; the safe wrappers below are the
//! ones the `cfg(feature = "b200")` hooks in snarkvm-algorithms call.
//!
//! NOT compiled in this repository's image (no cargo / rustc); `tests/test_abi.py::test_rust_binding_covers_the_header`
//! checks the declarations below against the header by name and arity.
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

pub mod fixture;

pub mod sys {
    use super::*;

    #[repr(C)]
    #[derive(Copy, Clone)]
    pub struct b200_error_t {
        pub code: i32,
        pub msg: *const c_char,
    }

    extern "C" {
        // lifecycle
        pub fn b200_init(device: c_int) -> b200_error_t;
        pub fn b200_shutdown();
        pub fn b200_abi_version() -> u32;
        pub fn b200_release_scratch() -> b200_error_t;
        pub fn b200_set_option(key: *const c_char, value: *const c_char) -> b200_error_t;
        pub fn b200_get_counter(name: *const c_char, out: *mut u64) -> b200_error_t;
        // VariableBase::msm
        pub fn b200_msm_g1_bls12_377(out_jacobian_144b: *mut c_void, points: *const c_void, npoints: usize, scalars: *const c_void, affine_stride: usize) -> b200_error_t;
        pub fn b200_msm_g1_bls12_377_device(d_out: *mut c_void, d_points: *const c_void, npoints: usize, d_scalars: *const c_void, affine_stride: usize, stream: *mut c_void) -> b200_error_t;
        pub fn b200_msm_batch_g1_bls12_377(out_jacobian: *mut c_void, points: *const c_void, scalars: *const c_void, offsets: *const u64, nmsm: usize, affine_stride: usize) -> b200_error_t;
        pub fn b200_msm_batch_g1_bls12_377_device(d_out: *mut c_void, d_points: *const c_void, d_scalars: *const c_void, d_offsets_u64: *const c_void, nmsm: usize, npoints: usize, affine_stride: usize, stream: *mut c_void) -> b200_error_t;
        pub fn b200_msm_submit(points: *const c_void, npoints: usize, scalars: *const c_void, affine_stride: usize, out_ticket: *mut u64) -> b200_error_t;
        pub fn b200_msm_wait(ticket: u64, out_jacobian_144b: *mut c_void) -> b200_error_t;
        // resident bases, KZG10
        pub fn b200_msm_register_bases(points: *const c_void, npoints: usize, affine_stride: usize, out_handle: *mut u64) -> b200_error_t;
        pub fn b200_msm_register_bases_device(d_points: *const c_void, npoints: usize, affine_stride: usize, stream: *mut c_void, out_handle: *mut u64) -> b200_error_t;
        pub fn b200_msm_register_bases_tabulated(points: *const c_void, npoints: usize, affine_stride: usize, window_bits: u32, out_handle: *mut u64) -> b200_error_t;
        pub fn b200_msm_register_bases_tabulated_device(d_points: *const c_void, npoints: usize, affine_stride: usize, window_bits: u32, stream: *mut c_void, out_handle: *mut u64) -> b200_error_t;
        pub fn b200_msm_registered(out_jacobian_144b: *mut c_void, handle: u64, scalars: *const c_void, nscalars: usize) -> b200_error_t;
        pub fn b200_msm_registered_device(d_out: *mut c_void, handle: u64, d_scalars: *const c_void, nscalars: usize, stream: *mut c_void) -> b200_error_t;
        pub fn b200_msm_release_bases(handle: u64) -> b200_error_t;
        pub fn b200_kzg_commit(out_jacobian_144b: *mut c_void, handle: u64, coeffs_mont: *const c_void, ncoeffs: usize) -> b200_error_t;
        pub fn b200_kzg_commit_device(d_out: *mut c_void, handle: u64, d_coeffs_mont: *const c_void, ncoeffs: usize, stream: *mut c_void) -> b200_error_t;
        pub fn b200_kzg_commit_batch(out_jacobian: *mut c_void, handle: u64, coeffs_mont: *const c_void, offsets: *const u64, k: usize) -> b200_error_t;
        pub fn b200_kzg_commit_batch_device(d_out: *mut c_void, handle: u64, d_coeffs_mont: *const c_void, offsets_host: *const u64, k: usize, stream: *mut c_void) -> b200_error_t;
        pub fn b200_kzg_open(out_jacobian_144b: *mut c_void, handle: u64, coeffs_mont: *const c_void, ncoeffs: usize, point_mont: *const c_void, out_eval_mont: *mut c_void) -> b200_error_t;
        pub fn b200_kzg_open_device(d_out: *mut c_void, handle: u64, d_coeffs_mont: *const c_void, ncoeffs: usize, d_point_mont: *const c_void, d_out_eval_mont: *mut c_void, stream: *mut c_void) -> b200_error_t;
        pub fn b200_fr_linear_combination_device(d_out: *mut c_void, d_polys: *const c_void, offsets_host: *const u64, k: usize, d_coeffs_mont: *const c_void, out_len: usize, stream: *mut c_void) -> b200_error_t;
        pub fn b200_fr_divide_by_linear_device(d_quotient: *mut c_void, d_poly: *const c_void, n: usize, d_point_mont: *const c_void, d_out_remainder: *mut c_void, stream: *mut c_void) -> b200_error_t;
        pub fn b200_msm_window_bits(npoints: usize) -> u32;
        pub fn b200_msm_affine_rounds(npoints: usize) -> u32;
        pub fn b200_msm_describe(npoints: usize, out4: *mut u32);
        pub fn b200_g1_sum_jacobian_device(d_out: *mut c_void, d_in: *const c_void, count: usize, stream: *mut c_void) -> b200_error_t;
        // curve output forms (Projective::to_affine / batch_normalization, CanonicalSerialize compressed)
        pub fn b200_g1_batch_normalize(out_affine: *mut c_void, in_jacobian: *const c_void, count: usize, affine_stride: usize) -> b200_error_t;
        pub fn b200_g1_batch_normalize_device(d_out_affine: *mut c_void, d_in_jacobian: *const c_void, count: usize, affine_stride: usize, stream: *mut c_void) -> b200_error_t;
        pub fn b200_g1_compress(out_48b: *mut c_void, in_jacobian: *const c_void, count: usize) -> b200_error_t;
        pub fn b200_g1_compress_device(d_out_48b: *mut c_void, d_in_jacobian: *const c_void, count: usize, stream: *mut c_void) -> b200_error_t;
        // EvaluationDomain
        pub fn b200_ntt_fr_bls12_377(inout: *mut c_void, log_n: u32, batch: usize, batch_stride_elems: usize, direction: c_int, coset: c_int) -> b200_error_t;
        pub fn b200_ntt_fr_bls12_377_device(d_inout: *mut c_void, log_n: u32, batch: usize, batch_stride_elems: usize, direction: c_int, coset: c_int, stream: *mut c_void) -> b200_error_t;
        // polynomial glue
        pub fn b200_fr_vec_op_device(op: c_int, d_out: *mut c_void, d_a: *const c_void, d_b: *const c_void, d_c: *const c_void, n: usize, b_is_scalar: c_int, stream: *mut c_void) -> b200_error_t;
        pub fn b200_fr_batch_inverse_device(d_inout: *mut c_void, n: usize, stream: *mut c_void) -> b200_error_t;
        pub fn b200_fr_divide_by_vanishing_on_coset_device(d_evals: *mut c_void, log_k: u32, log_h: u32, stream: *mut c_void) -> b200_error_t;
        // multi-GPU building blocks
        pub fn b200_fr_mul_powers_device(d_data: *mut c_void, log_n: u32, direction: c_int, kind: c_int, rows: u64, cols: u64, row_base: u64, col_base: u64, stream: *mut c_void) -> b200_error_t;
        pub fn b200_fr_exchange_transpose_device(d_src: *const c_void, dst_ptrs: *const *mut c_void, world: u32, rank: u32, r_local: u64, c: u64, log_n: u32, direction: c_int, twiddle: c_int, row_base: u64, stream: *mut c_void) -> b200_error_t;
        pub fn b200_fr_exchange_transpose_part_device(d_src_slab: *const c_void, dst_ptrs: *const *mut c_void, world: u32, rank: u32, r_local: u64, row_off: u64, r_count: u64, c: u64, col_lo: u64, col_cnt: u64, log_n: u32, direction: c_int, twiddle: c_int, row_base: u64, cta_limit: u32, stream: *mut c_void) -> b200_error_t;
        pub fn b200_ntt_rows_exchange_device(d_rows: *const c_void, dst_ptrs: *const *mut c_void, world: u32, rank: u32, rows: u64, log_len: u32, log_n_total: u32, direction: c_int, twiddle: c_int, row_base: u64, stream: *mut c_void) -> b200_error_t;
        pub fn b200_peer_buffer_alloc(bytes: usize, d_ptr: *mut *mut c_void, handle64: *mut c_void) -> b200_error_t;
        pub fn b200_peer_buffer_open(handle64: *const c_void, d_ptr: *mut *mut c_void) -> b200_error_t;
        pub fn b200_peer_buffer_close(d_ptr: *mut c_void) -> b200_error_t;
        pub fn b200_peer_buffer_free(d_ptr: *mut c_void) -> b200_error_t;
        // synthetic inputs and diagnostics
        pub fn b200_g1_synthetic_bases_device(d_out_points: *mut c_void, npoints: usize, affine_stride: usize, seed: u64, stream: *mut c_void) -> b200_error_t;
        pub fn b200_debug_field_op(op: c_int, out: *mut c_void, a: *const c_void, b: *const c_void, n: usize) -> b200_error_t;
        pub fn b200_debug_g1_op(op: c_int, out_jacobian: *mut c_void, a_affine: *const c_void, b: *const c_void, n: usize, affine_stride: usize) -> b200_error_t;
        pub fn b200_debug_microbench(kind: c_int, iters: u32, out_ms: *mut f32, out_ops: *mut f64) -> b200_error_t;
        pub fn b200_profile_begin();
        pub fn b200_profile_end(buf: *mut c_char, buflen: usize) -> b200_error_t;
        pub fn b200_kernel_launch_count() -> u64;
    }
}

pub const ABI_VERSION: u32 = 2;

#[derive(Debug)]
pub struct Error {
    pub code: i32,
    pub msg: String,
}

impl core::fmt::Display for Error {
    fn fmt(&self, f: &mut core::fmt::Formatter<'_>) -> core::fmt::Result {
        write!(f, "snarkos_b200 error {}: {}", self.code, self.msg)
    }
}
impl std::error::Error for Error {}

fn check(e: sys::b200_error_t) -> Result<(), Error> {
    if e.code == 0 {
        return Ok(());
    }
    let msg = if e.msg.is_null() { String::new() } else { unsafe { std::ffi::CStr::from_ptr(e.msg) }.to_string_lossy().into_owned() };
    Err(Error { code: e.code, msg })
}

#[derive(Copy, Clone, Debug, PartialEq, Eq)]
pub enum Direction {
    Forward = 0,
    Inverse = 1,
}
#[derive(Copy, Clone, Debug, PartialEq, Eq)]
pub enum Kind {
    Standard = 0,
    Coset = 1,
}

/// Binds the process to CUDA device `device` (-1: the current one) and checks the ABI version of the loaded library.
pub fn init(device: i32) -> Result<(), Error> {
    let v = unsafe { sys::b200_abi_version() };
    if v != ABI_VERSION {
        return Err(Error { code: -1, msg: format!("libsnarkos_b200 ABI version {v}, binding expects {ABI_VERSION}") });
    }
    check(unsafe { sys::b200_init(device as c_int) })
}

pub fn set_option(key: &str, value: &str) -> Result<(), Error> {
    let k = std::ffi::CString::new(key).unwrap();
    let v = std::ffi::CString::new(value).unwrap();
    check(unsafe { sys::b200_set_option(k.as_ptr(), v.as_ptr()) })
}

/// `VariableBase::msm` for BLS12-377 G1.  `A` = G1Affine (x, y Montgomery, infinity flag at byte 96), `S` = BigInteger256
/// (canonical), `P` = G1Projective (Jacobian X, Y, Z Montgomery; 144 bytes).  The result is the same group element
/// snarkVM computes; its Jacobian representative may differ (compare after `to_affine`).
pub fn msm<A, P: Default, S>(points: &[A], scalars: &[S]) -> Result<P, Error> {
    assert_eq!(core::mem::size_of::<S>(), 32);
    assert_eq!(core::mem::size_of::<P>(), 144);
    let n = points.len().min(scalars.len());
    let mut out = P::default();
    check(unsafe {
        sys::b200_msm_g1_bls12_377(&mut out as *mut P as *mut c_void, points.as_ptr() as *const c_void, n, scalars.as_ptr() as *const c_void, core::mem::size_of::<A>())
    })?;
    Ok(out)
}

/// `EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place` on a slice already resized to the domain.
pub fn ntt<T>(log_n: u32, data: &mut [T], dir: Direction, kind: Kind) -> Result<(), Error> {
    assert_eq!(core::mem::size_of::<T>(), 32);
    assert_eq!(data.len(), 1usize << log_n);
    check(unsafe { sys::b200_ntt_fr_bls12_377(data.as_mut_ptr() as *mut c_void, log_n, 1, data.len(), dir as c_int, kind as c_int) })
}

/// `batch` polynomials of 2^log_n elements stored back to back (Varuna transforms the polynomials of a round together).
pub fn ntt_batch<T>(log_n: u32, data: &mut [T], batch: usize, dir: Direction, kind: Kind) -> Result<(), Error> {
    assert_eq!(core::mem::size_of::<T>(), 32);
    assert_eq!(data.len(), batch << log_n);
    check(unsafe { sys::b200_ntt_fr_bls12_377(data.as_mut_ptr() as *mut c_void, log_n, batch, 1usize << log_n, dir as c_int, kind as c_int) })
}

/// Device-resident base set: `powers_of_beta_g` of the SRS, registered once per process.
pub struct ResidentBases {
    handle: u64,
    len: usize,
}

impl ResidentBases {
    pub fn new<A>(points: &[A]) -> Result<Self, Error> {
        let mut handle = 0u64;
        check(unsafe { sys::b200_msm_register_bases(points.as_ptr() as *const c_void, points.len(), core::mem::size_of::<A>(), &mut handle) })?;
        Ok(Self { handle, len: points.len() })
    }

    pub fn len(&self) -> usize {
        self.len
    }

    pub fn is_empty(&self) -> bool {
        self.len == 0
    }

    /// `VariableBase::msm(&powers[..scalars.len()], scalars)`
    pub fn msm<P: Default, S>(&self, scalars: &[S]) -> Result<P, Error> {
        assert_eq!(core::mem::size_of::<S>(), 32);
        assert_eq!(core::mem::size_of::<P>(), 144);
        let mut out = P::default();
        check(unsafe { sys::b200_msm_registered(&mut out as *mut P as *mut c_void, self.handle, scalars.as_ptr() as *const c_void, scalars.len()) })?;
        Ok(out)
    }

    /// `KZG10::commit`'s MSM: coefficients as held in a `DensePolynomial` (Montgomery `Fr`), converted on the device.
    pub fn commit<P: Default, F>(&self, coeffs: &[F]) -> Result<P, Error> {
        assert_eq!(core::mem::size_of::<F>(), 32);
        assert_eq!(core::mem::size_of::<P>(), 144);
        let mut out = P::default();
        check(unsafe { sys::b200_kzg_commit(&mut out as *mut P as *mut c_void, self.handle, coeffs.as_ptr() as *const c_void, coeffs.len()) })?;
        Ok(out)
    }

    /// All commitments of a round in one launch set.
    pub fn commit_batch<P: Default + Clone, F: Copy>(&self, polys: &[&[F]]) -> Result<Vec<P>, Error> {
        assert_eq!(core::mem::size_of::<F>(), 32);
        assert_eq!(core::mem::size_of::<P>(), 144);
        let mut flat: Vec<F> = Vec::with_capacity(polys.iter().map(|p| p.len()).sum());
        let mut offsets = vec![0u64];
        for p in polys {
            flat.extend_from_slice(p);
            offsets.push(flat.len() as u64);
        }
        let mut out = vec![P::default(); polys.len()];
        check(unsafe { sys::b200_kzg_commit_batch(out.as_mut_ptr() as *mut c_void, self.handle, flat.as_ptr() as *const c_void, offsets.as_ptr(), polys.len()) })?;
        Ok(out)
    }

    /// `KZG10::open`'s group arithmetic: the commitment to (p(X) - p(z)) / (X - z), and p(z).
    pub fn open<P: Default, F: Default>(&self, coeffs: &[F], point: &F) -> Result<(P, F), Error> {
        assert_eq!(core::mem::size_of::<F>(), 32);
        assert_eq!(core::mem::size_of::<P>(), 144);
        let mut out = P::default();
        let mut eval = F::default();
        check(unsafe {
            sys::b200_kzg_open(&mut out as *mut P as *mut c_void, self.handle, coeffs.as_ptr() as *const c_void, coeffs.len(), point as *const F as *const c_void, &mut eval as *mut F as *mut c_void)
        })?;
        Ok((out, eval))
    }
}

impl Drop for ResidentBases {
    fn drop(&mut self) {
        unsafe { sys::b200_msm_release_bases(self.handle) };
    }
}

/// Ticket of a small MSM handed to the coalescing queue (the verifier's linear combinations, issued from many threads).
pub struct MsmTicket(u64);

/// SAFETY: `points` and `scalars` must stay alive and unchanged until `wait` returns.
pub unsafe fn msm_submit<A, S>(points: &[A], scalars: &[S]) -> Result<MsmTicket, Error> {
    assert_eq!(core::mem::size_of::<S>(), 32);
    let n = points.len().min(scalars.len());
    let mut t = 0u64;
    check(sys::b200_msm_submit(points.as_ptr() as *const c_void, n, scalars.as_ptr() as *const c_void, core::mem::size_of::<A>(), &mut t))?;
    Ok(MsmTicket(t))
}

impl MsmTicket {
    pub fn wait<P: Default>(self) -> Result<P, Error> {
        assert_eq!(core::mem::size_of::<P>(), 144);
        let mut out = P::default();
        check(unsafe { sys::b200_msm_wait(self.0, &mut out as *mut P as *mut c_void) })?;
        Ok(out)
    }
}

/// `Projective::batch_normalization` + `to_affine`: `count` Jacobian images -> affine images with stride `size_of::<A>()`.
pub fn batch_normalize<A: Default + Clone, P>(points: &[P]) -> Result<Vec<A>, Error> {
    assert_eq!(core::mem::size_of::<P>(), 144);
    let mut out = vec![A::default(); points.len()];
    check(unsafe { sys::b200_g1_batch_normalize(out.as_mut_ptr() as *mut c_void, points.as_ptr() as *const c_void, points.len(), core::mem::size_of::<A>()) })?;
    Ok(out)
}

/// Compressed encoding (48 bytes each) of `count` Jacobian images: what `CanonicalSerialize::serialize_compressed` /
/// `ToBytes::write_le` emit for a `G1Affine` inside a Varuna proof.
pub fn compress<P>(points: &[P]) -> Result<Vec<[u8; 48]>, Error> {
    assert_eq!(core::mem::size_of::<P>(), 144);
    let mut out = vec![[0u8; 48]; points.len()];
    check(unsafe { sys::b200_g1_compress(out.as_mut_ptr() as *mut c_void, points.as_ptr() as *const c_void, points.len()) })?;
    Ok(out)
}
