// Builds (or locates) libsnarkos_b200.so and links it.  SNARKOS_B200_DIR points at a checkout of this repository.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("SNARKOS_B200_DIR").expect("set SNARKOS_B200_DIR to the snarkos_b200 checkout"));
    let status = Command::new("make").arg("-C").arg(root.join("snarkos_b200/csrc")).arg("-j8").status().expect("make");
    assert!(status.success(), "nvcc build of libsnarkos_b200.so failed (needs CUDA >= 12.8, sm_100a)");
    println!("cargo:rustc-link-search=native={}", root.join("snarkos_b200").display());
    println!("cargo:rustc-link-lib=dylib=snarkos_b200");
    println!("cargo:rerun-if-changed={}", root.join("include/snarkos_b200.h").display());
}
