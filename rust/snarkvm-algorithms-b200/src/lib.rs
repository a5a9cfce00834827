//! Rust binding of include/snarkos_b200.h.  Generic over the snarkVM types only through their memory layout, like
//! upstream's snarkvm-algorithms-cuda: the caller passes slices, this crate passes pointers + size_of::<T>().
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

#[repr(C)]
pub struct b200_error_t {
    pub code: i32,
    pub msg: *const c_char,
}

extern "C" {
    pub fn b200_init(device: c_int) -> b200_error_t;
    pub fn b200_shutdown();
    pub fn b200_abi_version() -> u32;
    pub fn b200_msm_g1_bls12_377(out_jacobian_144b: *mut c_void, points: *const c_void, npoints: usize, scalars: *const c_void, affine_stride: usize) -> b200_error_t;
    pub fn b200_msm_register_bases(points: *const c_void, npoints: usize, affine_stride: usize, out_handle: *mut u64) -> b200_error_t;
    pub fn b200_msm_registered(out_jacobian_144b: *mut c_void, handle: u64, scalars: *const c_void, nscalars: usize) -> b200_error_t;
    pub fn b200_msm_release_bases(handle: u64) -> b200_error_t;
    pub fn b200_ntt_fr_bls12_377(inout: *mut c_void, log_n: u32, batch: usize, batch_stride_elems: usize, direction: c_int, coset: c_int) -> b200_error_t;
}

#[derive(Debug)]
pub struct Error {
    pub code: i32,
    pub msg: String,
}

fn check(e: b200_error_t) -> Result<(), Error> {
    if e.code == 0 {
        return Ok(());
    }
    let msg = unsafe { std::ffi::CStr::from_ptr(e.msg) }.to_string_lossy().into_owned();
    Err(Error { code: e.code, msg })
}

#[derive(Copy, Clone)]
pub enum Direction { Forward = 0, Inverse = 1 }
#[derive(Copy, Clone)]
pub enum Kind { Standard = 0, Coset = 1 }

/// `VariableBase::msm` for BLS12-377 G1.  `A` = G1Affine (x, y Montgomery, infinity flag at byte 96), `S` = BigInteger256
/// (canonical), `P` = G1Projective (Jacobian X, Y, Z Montgomery; 144 bytes).
pub fn msm<A, P: Default, S>(points: &[A], scalars: &[S]) -> Result<P, Error> {
    assert_eq!(core::mem::size_of::<S>(), 32);
    assert_eq!(core::mem::size_of::<P>(), 144);
    let n = points.len().min(scalars.len());
    let mut out = P::default();
    check(unsafe {
        b200_msm_g1_bls12_377(&mut out as *mut P as *mut c_void, points.as_ptr() as *const c_void, n, scalars.as_ptr() as *const c_void, core::mem::size_of::<A>())
    })?;
    Ok(out)
}

/// `EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place` on a slice already resized to the domain.
pub fn ntt<T>(log_n: u32, data: &mut [T], dir: Direction, kind: Kind) -> Result<(), Error> {
    assert_eq!(core::mem::size_of::<T>(), 32);
    assert_eq!(data.len(), 1usize << log_n);
    check(unsafe { b200_ntt_fr_bls12_377(data.as_mut_ptr() as *mut c_void, log_n, 1, data.len(), dir as c_int, kind as c_int) })
}
