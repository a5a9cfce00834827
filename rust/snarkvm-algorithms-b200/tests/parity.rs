//! Parity of the B200 backend against snarkVM (rev dea322b) on a box with a B200 and a Rust toolchain:
//!   SNARKOS_B200_DIR=<checkout> cargo test --release --features harness -- --nocapture
//! Mirrors snarkVM's own tests for this path (variable_base::tests::{test_msm, test_msm_cuda}: seeded random inputs,
//! CPU result vs accelerated result; fft domain tests: round trips) and compares at the level the C ABI returns:
//! MSM results after to_affine(), NTT outputs limb for limb.
use snarkvm_algorithms::{fft::EvaluationDomain, msm::VariableBase};
use snarkvm_algorithms_b200 as b200;
use snarkvm_curves::{
    bls12_377::{Fr, G1Affine, G1Projective},
    AffineCurve, ProjectiveCurve,
};
use snarkvm_fields::{PrimeField, Zero};
use snarkvm_utilities::{rand::TestRng, ToBytes, Uniform};

fn sample(n: usize, rng: &mut TestRng) -> (Vec<G1Affine>, Vec<<Fr as PrimeField>::BigInteger>) {
    let bases = (0..n).map(|_| G1Projective::rand(rng).to_affine()).collect();
    let scalars = (0..n).map(|_| Fr::rand(rng).to_bigint()).collect();
    (bases, scalars)
}

#[test]
fn layout_is_what_the_c_abi_assumes() {
    let g = G1Affine::prime_subgroup_generator();
    assert_eq!(core::mem::size_of::<Fr>(), 32);
    assert_eq!(core::mem::size_of::<G1Projective>(), 144);
    assert_eq!(b200::fixture::offset_of(&g, &g.x), 0);
    assert_eq!(b200::fixture::offset_of(&g, &g.y), 48);
    assert_eq!(b200::fixture::offset_of(&g, &g.infinity), 96);
    assert!(core::mem::size_of::<G1Affine>() >= 97 && core::mem::size_of::<G1Affine>() % 8 == 0);
}

#[test]
fn msm_matches_variable_base() {
    b200::init(-1).expect("b200_init");
    let mut rng = TestRng::fixed(123456789);
    for n in [0usize, 1, 2, 14, 15, 31, 32, 33, 100, 1000, 1 << 10, 1 << 12, 1 << 16, 1 << 20] {
        let (bases, scalars) = sample(n, &mut rng);
        let want: G1Projective = VariableBase::msm(&bases, &scalars);
        let got: G1Projective = b200::msm(&bases, &scalars).expect("b200 msm");
        assert_eq!(got.to_affine(), want.to_affine(), "n = {n}");
    }
}

#[test]
fn msm_degenerate_inputs() {
    b200::init(-1).expect("b200_init");
    let mut rng = TestRng::fixed(987654321);
    let (mut bases, mut scalars) = sample(256, &mut rng);
    let g = G1Affine::prime_subgroup_generator();
    bases[0] = G1Affine::zero();
    bases[1] = g;
    bases[2] = g;
    bases[3] = -g;
    scalars[2] = scalars[1];
    scalars[3] = scalars[1];
    scalars[4] = Fr::zero().to_bigint();
    scalars[5] = (-Fr::from(1u64)).to_bigint();
    let want: G1Projective = VariableBase::msm(&bases, &scalars);
    let got: G1Projective = b200::msm(&bases, &scalars).expect("b200 msm");
    assert_eq!(got.to_affine(), want.to_affine());
    // all-equal scalars and all-equal points
    let same_s = vec![scalars[7]; 256];
    let same_p = vec![bases[7]; 256];
    for (b, s) in [(&bases, &same_s), (&same_p, &scalars), (&same_p, &same_s)] {
        let want: G1Projective = VariableBase::msm(b, s);
        let got: G1Projective = b200::msm(b, s).expect("b200 msm");
        assert_eq!(got.to_affine(), want.to_affine());
    }
}

#[test]
fn ntt_matches_evaluation_domain() {
    b200::init(-1).expect("b200_init");
    let mut rng = TestRng::fixed(123456789);
    for log_n in [0u32, 1, 2, 3, 5, 8, 10, 12, 14, 16, 20, 22] {
        let n = 1usize << log_n;
        let domain = EvaluationDomain::<Fr>::new(n).unwrap();
        let input: Vec<Fr> = (0..n).map(|_| Fr::rand(&mut rng)).collect();
        let cases: [(b200::Direction, b200::Kind, fn(&EvaluationDomain<Fr>, &mut Vec<Fr>)); 4] = [
            (b200::Direction::Forward, b200::Kind::Standard, |d, v| d.fft_in_place(v)),
            (b200::Direction::Inverse, b200::Kind::Standard, |d, v| d.ifft_in_place(v)),
            (b200::Direction::Forward, b200::Kind::Coset, |d, v| d.coset_fft_in_place(v)),
            (b200::Direction::Inverse, b200::Kind::Coset, |d, v| d.coset_ifft_in_place(v)),
        ];
        for (dir, kind, cpu) in cases {
            let mut want = input.clone();
            cpu(&domain, &mut want);
            let mut got = input.clone();
            b200::ntt(log_n, &mut got, dir, kind).expect("b200 ntt");
            assert!(got == want, "log_n = {log_n}, {dir:?} {kind:?}");
        }
    }
}

#[test]
fn compressed_encoding_matches_to_bytes_le() {
    b200::init(-1).expect("b200_init");
    let mut rng = TestRng::fixed(55);
    let mut pts: Vec<G1Projective> = (0..100).map(|_| G1Projective::rand(&mut rng)).collect();
    pts[17] = G1Projective::zero();
    let got = b200::compress(&pts).expect("b200 compress");
    for (p, g) in pts.iter().zip(got.iter()) {
        assert_eq!(&p.to_affine().to_bytes_le().unwrap()[..], &g[..]);
    }
    let aff: Vec<G1Affine> = b200::batch_normalize(&pts).expect("b200 normalize");
    for (p, a) in pts.iter().zip(aff.iter()) {
        assert_eq!(p.to_affine(), *a);
    }
}
