//! Fixture files that pin this repository's oracle and CUDA path to snarkVM itself (tests/golden/FIXTURES.md).
//!
//! All integers little-endian.  Field elements and points are RAW MEMORY IMAGES of the snarkVM types (that is what
//! crosses the C ABI), so the files also pin rustc's layout of `G1Affine`.
//!
//! MSM file  `snarkvm_msm_<tag>.bin`
//!   0   8  magic  "B2MSM001"
//!   8   8  n                       number of (base, scalar) pairs
//!   16  8  stride                  size_of::<G1Affine>()
//!   24  8  off_x, 32 8 off_y, 40 8 off_infinity   byte offsets of the fields inside a G1Affine
//!   48  8  proj_bytes              size_of::<G1Projective>() (144)
//!   56  n * stride                 bases (raw)
//!   ..  n * 32                     scalars (BigInteger256, canonical)
//!   ..  proj_bytes                 VariableBase::msm result (raw G1Projective: X, Y, Z Montgomery)
//!   ..  stride                     the same result through to_affine() (raw G1Affine)
//!   ..  48                         the same result, compressed serialisation (to_bytes_le of the G1Affine)
//!
//! NTT file  `snarkvm_ntt_<tag>.bin`
//!   0   8  magic  "B2NTT001"
//!   8   8  log_n
//!   16  8  n_in                    number of input coefficients (<= 2^log_n; snarkVM zero-pads)
//!   24  n_in * 32                  input (raw Fr, Montgomery)
//!   ..  2^log_n * 32  x 4          fft_in_place, ifft_in_place, coset_fft_in_place, coset_ifft_in_place of the input
use std::io::{self, Write};

pub const MSM_MAGIC: &[u8; 8] = b"B2MSM001";
pub const NTT_MAGIC: &[u8; 8] = b"B2NTT001";

/// raw bytes of a slice of plain-old-data values
pub fn raw<T>(v: &[T]) -> &[u8] {
    unsafe { core::slice::from_raw_parts(v.as_ptr() as *const u8, core::mem::size_of_val(v)) }
}

pub fn raw_one<T>(v: &T) -> &[u8] {
    unsafe { core::slice::from_raw_parts(v as *const T as *const u8, core::mem::size_of::<T>()) }
}

pub fn put_u64<W: Write>(w: &mut W, v: u64) -> io::Result<()> {
    w.write_all(&v.to_le_bytes())
}

/// byte offset of a field inside its struct
pub fn offset_of<S, F>(s: &S, f: &F) -> u64 {
    (f as *const F as usize - s as *const S as usize) as u64
}
