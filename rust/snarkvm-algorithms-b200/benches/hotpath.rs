//! The CPU baseline the bench contract asks for, as the reference itself: snarkVM's VariableBase::msm and
//! EvaluationDomain on the box's host cores next to the B200 backend on identical inputs.
//!   SNARKOS_B200_DIR=<checkout> cargo bench --features harness [-- <log2 n, default 20>]
//! Prints one JSON line per measurement (std::time::Instant; no criterion so that it builds offline once vendored).
use std::time::Instant;

use snarkvm_algorithms::{fft::EvaluationDomain, msm::VariableBase};
use snarkvm_algorithms_b200 as b200;
use snarkvm_curves::{
    bls12_377::{Fr, G1Affine, G1Projective},
    ProjectiveCurve,
};
use snarkvm_fields::PrimeField;
use snarkvm_utilities::{rand::TestRng, Uniform};

fn time<F: FnMut()>(reps: usize, mut f: F) -> f64 {
    f();
    let t = Instant::now();
    for _ in 0..reps {
        f();
    }
    t.elapsed().as_secs_f64() / reps as f64
}

fn main() {
    let log_n: u32 = std::env::args().skip(1).find_map(|a| a.parse().ok()).unwrap_or(20);
    let n = 1usize << log_n;
    let cores = std::thread::available_parallelism().map(|c| c.get()).unwrap_or(1);
    b200::init(-1).expect("b200_init");
    let mut rng = TestRng::fixed(123456789);
    // distinct bases k_i * G by a running sum of random points (n scalar multiplications would dominate set-up)
    let mut acc = G1Projective::rand(&mut rng);
    let step = G1Projective::rand(&mut rng);
    let proj: Vec<G1Projective> = (0..n).map(|_| { acc += step; acc }).collect();
    let mut proj_n = proj.clone();
    G1Projective::batch_normalization(&mut proj_n);
    let bases: Vec<G1Affine> = proj_n.iter().map(|p| p.to_affine()).collect();
    let scalars: Vec<_> = (0..n).map(|_| Fr::rand(&mut rng).to_bigint()).collect();

    let mut cpu_out = G1Projective::default();
    let cpu = time(2, || cpu_out = VariableBase::msm(&bases, &scalars));
    let mut gpu_out = G1Projective::default();
    let gpu = time(5, || gpu_out = b200::msm(&bases, &scalars).expect("b200 msm"));
    assert_eq!(cpu_out.to_affine(), gpu_out.to_affine());
    println!("{{\"op\": \"msm\", \"log_n\": {log_n}, \"cores\": {cores}, \"snarkvm_cpu_s\": {cpu:.6}, \"b200_e2e_s\": {gpu:.6}, \"cpu_mpoints_s\": {:.3}, \"b200_mpoints_s\": {:.3}}}",
             n as f64 / cpu / 1e6, n as f64 / gpu / 1e6);

    let domain = EvaluationDomain::<Fr>::new(n).unwrap();
    let input: Vec<Fr> = (0..n).map(|_| Fr::rand(&mut rng)).collect();
    let mut a = input.clone();
    let cpu = time(3, || { a.copy_from_slice(&input); domain.fft_in_place(&mut a) });
    let mut b = input.clone();
    let gpu = time(5, || { b.copy_from_slice(&input); b200::ntt(log_n, &mut b, b200::Direction::Forward, b200::Kind::Standard).expect("b200 ntt") });
    assert!(a == b);
    println!("{{\"op\": \"fft_in_place\", \"log_n\": {log_n}, \"cores\": {cores}, \"snarkvm_cpu_s\": {cpu:.6}, \"b200_e2e_s\": {gpu:.6}, \"cpu_gelem_s\": {:.4}, \"b200_gelem_s\": {:.4}}}",
             n as f64 / cpu / 1e9, n as f64 / gpu / 1e9);
}
