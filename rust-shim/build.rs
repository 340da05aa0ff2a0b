// rust-shim/build.rs -- NOT COMPILED in this repository's environment (no Rust toolchain); see INTEGRATION.md.
// Builds libnbody_b200.so from the CUDA sources with the system nvcc (sm_100a only) and tells cargo
// where it is, so that the run-time loader in nbody.rs (or a direct -lnbody_b200 link) finds it.
use std::process::Command;

fn main() {
    let repo = std::env::var("NBODY_B200_REPO").expect("set NBODY_B200_REPO to the nbody_b200 checkout");
    let status = Command::new("make")
        .args(&["-C", &format!("{}/rust_exp_b200/csrc", repo), "-j8", "../libnbody_b200.so"])
        .status()
        .expect("failed to run make (needs nvcc 12.9+ for sm_100a)");
    assert!(status.success());
    println!("cargo:rustc-link-search=native={}/rust_exp_b200", repo);
    println!("cargo:rustc-env=NBODY_B200_LIB={}/rust_exp_b200/libnbody_b200.so", repo);
    println!("cargo:rerun-if-changed={}/include/nbody_b200.h", repo);
}
