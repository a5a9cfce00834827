//! Writes tests/golden/snarkvm_{msm,ntt}_*.bin from snarkVM (rev dea322b, the one snarkOS pins) on seeded inputs.
//! Needs cargo + network (to fetch snarkVM), NO GPU.  Format: src/fixture.rs / tests/golden/FIXTURES.md.
//!
//!   SNARKOS_B200_DIR=<checkout> cargo run --release --features harness --bin make_fixtures
//!
//! The committed pytest `tests/test_snarkvm_fixtures.py` consumes the files: without a GPU it checks the C oracle
//! against them, with a GPU the CUDA path through the C ABI.
use std::{fs::File, io::BufWriter, io::Write, path::PathBuf};

use snarkvm_algorithms::{fft::EvaluationDomain, msm::VariableBase};
use snarkvm_algorithms_b200::fixture::*;
use snarkvm_curves::{
    bls12_377::{Fr, G1Affine, G1Projective},
    AffineCurve, ProjectiveCurve,
};
use snarkvm_fields::{PrimeField, Zero};
use snarkvm_utilities::{rand::TestRng, ToBytes, Uniform};

fn out_dir() -> PathBuf {
    let root = PathBuf::from(std::env::var("SNARKOS_B200_DIR").expect("set SNARKOS_B200_DIR to the snarkos_b200 checkout"));
    root.join("tests").join("golden")
}

fn write_msm(tag: &str, bases: &[G1Affine], scalars: &[<Fr as PrimeField>::BigInteger]) -> std::io::Result<()> {
    let result: G1Projective = VariableBase::msm(bases, scalars);
    let affine = result.to_affine();
    let mut w = BufWriter::new(File::create(out_dir().join(format!("snarkvm_msm_{tag}.bin")))?);
    let probe = G1Affine::prime_subgroup_generator();
    w.write_all(MSM_MAGIC)?;
    put_u64(&mut w, bases.len() as u64)?;
    put_u64(&mut w, core::mem::size_of::<G1Affine>() as u64)?;
    put_u64(&mut w, offset_of(&probe, &probe.x))?;
    put_u64(&mut w, offset_of(&probe, &probe.y))?;
    put_u64(&mut w, offset_of(&probe, &probe.infinity))?;
    put_u64(&mut w, core::mem::size_of::<G1Projective>() as u64)?;
    w.write_all(raw(bases))?;
    w.write_all(raw(scalars))?;
    w.write_all(raw_one(&result))?;
    w.write_all(raw_one(&affine))?;
    let compressed = affine.to_bytes_le().expect("to_bytes_le");
    assert_eq!(compressed.len(), 48);
    w.write_all(&compressed)?;
    w.flush()
}

fn write_ntt(tag: &str, log_n: u32, input: &[Fr]) -> std::io::Result<()> {
    let domain = EvaluationDomain::<Fr>::new(1usize << log_n).expect("domain");
    assert_eq!(domain.size(), 1usize << log_n);
    let mut w = BufWriter::new(File::create(out_dir().join(format!("snarkvm_ntt_{tag}.bin")))?);
    w.write_all(NTT_MAGIC)?;
    put_u64(&mut w, log_n as u64)?;
    put_u64(&mut w, input.len() as u64)?;
    w.write_all(raw(input))?;
    let mut v = input.to_vec();
    domain.fft_in_place(&mut v);
    w.write_all(raw(&v))?;
    let mut v = input.to_vec();
    domain.ifft_in_place(&mut v);
    w.write_all(raw(&v))?;
    let mut v = input.to_vec();
    domain.coset_fft_in_place(&mut v);
    w.write_all(raw(&v))?;
    let mut v = input.to_vec();
    domain.coset_ifft_in_place(&mut v);
    w.write_all(raw(&v))?;
    w.flush()
}

fn main() -> std::io::Result<()> {
    std::fs::create_dir_all(out_dir())?;
    // the seed the reference's own tests use (/root/reference/node/bft/tests/common/utils.rs:98)
    let mut rng = TestRng::fixed(123456789);

    // MSM: the sizes of snarkVM's variable_base::tests plus the dispatch edges (n < 15 bit-serial, c = 1 below 32)
    for &n in &[1usize, 2, 14, 15, 31, 32, 100, 1000, 1 << 12, 1 << 16] {
        let bases: Vec<G1Affine> = (0..n).map(|_| G1Projective::rand(&mut rng).to_affine()).collect();
        let scalars: Vec<_> = (0..n).map(|_| Fr::rand(&mut rng).to_bigint()).collect();
        write_msm(&format!("n{n}"), &bases, &scalars)?;
    }
    // degenerate inputs: infinity among the bases, zero / one / r - 1 scalars, repeated points
    {
        let g = G1Affine::prime_subgroup_generator();
        let mut bases: Vec<G1Affine> = (0..64).map(|_| G1Projective::rand(&mut rng).to_affine()).collect();
        let mut scalars: Vec<_> = (0..64).map(|_| Fr::rand(&mut rng).to_bigint()).collect();
        bases[3] = G1Affine::zero();
        bases[10] = g;
        bases[11] = g;
        bases[12] = -g;
        scalars[5] = Fr::zero().to_bigint();
        scalars[6] = Fr::from(1u64).to_bigint();
        scalars[7] = (-Fr::from(1u64)).to_bigint();
        scalars[10] = scalars[11];
        scalars[12] = scalars[11];
        write_msm("edge", &bases, &scalars)?;
    }
    // NTT: one size per pass structure of the CUDA plan (1, 2 and 3 passes) and a zero-padded input
    for &log_n in &[0u32, 1, 2, 3, 8, 12, 16, 20] {
        let input: Vec<Fr> = (0..1usize << log_n).map(|_| Fr::rand(&mut rng)).collect();
        write_ntt(&format!("log{log_n}"), log_n, &input)?;
    }
    let input: Vec<Fr> = (0..700).map(|_| Fr::rand(&mut rng)).collect();
    write_ntt("log10_padded", 10, &input)?;
    println!("fixtures written to {}", out_dir().display());
    Ok(())
}
