// rust-shim/build.rs -- NOT COMPILED in this repository's environment (no Rust toolchain); see INTEGRATION.md.
// Builds libnbody_b200.a from the CUDA sources with the system nvcc (sm_100a only) and links it, plus what its
// objects need, into the rust_exp static archive.  `[package] build = "build.rs"` in Cargo.toml enables it.
use std::process::Command;

fn main() {
    let repo = std::env::var("NBODY_B200_REPO").expect("set NBODY_B200_REPO to the nbody_b200 checkout");
    let cuda = std::env::var("CUDA_HOME").unwrap_or("/usr/local/cuda".to_string());
    let status = Command::new("make")
        .args(&["-C", &format!("{}/rust_exp_b200/csrc", repo), "-j8", "../libnbody_b200.a"])
        .status()
        .expect("failed to run make (needs nvcc 12.9+ for sm_100a)");
    assert!(status.success());
    println!("cargo:rustc-link-search=native={}/rust_exp_b200", repo);
    println!("cargo:rustc-link-search=native={}/lib64", cuda);
    println!("cargo:rustc-link-lib=static=nbody_b200");   // bundled into librust_exp.a (staticlib output)
    // libnbody_b200.a's objects are nvcc/g++ output: they need the static CUDA runtime and the C++ runtime at the
    // FINAL link (the Haskell package's, rust-exp.cabal:46-47 -- see INTEGRATION.md for the extra-libraries line)
    println!("cargo:rustc-link-lib=static=cudart_static");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=dl");
    println!("cargo:rustc-link-lib=dylib=rt");
    println!("cargo:rustc-link-lib=dylib=pthread");
    println!("cargo:rerun-if-changed={}/include/nbody_b200.h", repo);
}
