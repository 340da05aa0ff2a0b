// rust-shim/nbody.rs -- replacement for rs-src/nbody.rs of blitzcode/rust-exp (option A).
//
// Six forwarders, no logic, no state: the particle set lives in GPU memory inside libnbody_b200
// (include/nbody_b200.h).  The crate stays a staticlib (Cargo.toml:5-7) and keeps exporting the unmangled nb_*
// symbols that hs-src/RustNBodyExperiment.hs:101-106 imports; each forwards to the b200_-prefixed implementation in
// libnbody_b200.a, which build.rs links into the archive.  The archive keeps its own unprefixed nb_* in a separate
// object (nb_alias.o) that the linker never pulls in here, so nothing is defined twice.
//
// NOT COMPILED in the build environment of this repository (no Rust toolchain there).  Its exact C restatement,
// tests/c_host/shim_host.c, IS compiled, statically linked against libnbody_b200.a and run on the GPU by
// tests/test_c_host.py; this file is kept free of logic so that the difference is syntax only.

use std::os::raw::{c_float, c_int};

extern "C" {
    fn b200_nb_num_particles() -> c_int;
    fn b200_nb_random_disk(num_particles: c_int);
    fn b200_nb_stable_orbits(num_particles: c_int, rmin: c_float, rmax: c_float);
    fn b200_nb_step_brute_force(dt: c_float);
    fn b200_nb_step_barnes_hut(theta: c_float, dt: c_float, nthreads: c_int);
    fn b200_nb_draw(w: c_int, h: c_int, fb: *mut u32);
}

#[no_mangle]
pub extern fn nb_num_particles() -> i32 { unsafe { b200_nb_num_particles() } }

#[no_mangle]
pub extern fn nb_random_disk(num_particles: i32) -> () { unsafe { b200_nb_random_disk(num_particles) } }

#[no_mangle]
pub extern fn nb_stable_orbits(num_particles: i32, rmin: f32, rmax: f32) -> () {
    unsafe { b200_nb_stable_orbits(num_particles, rmin, rmax) }
}

#[no_mangle]
pub extern fn nb_step_brute_force(dt: f32) -> () { unsafe { b200_nb_step_brute_force(dt) } }

// NOTE argument order: theta first (rs-src/nbody.rs:187).
#[no_mangle]
pub extern fn nb_step_barnes_hut(theta: f32, dt: f32, nthreads: i32) -> () {
    unsafe { b200_nb_step_barnes_hut(theta, dt, nthreads) }
}

#[no_mangle]
pub extern fn nb_draw(w: i32, h: i32, fb: *mut u32) -> () { unsafe { b200_nb_draw(w, h, fb) } }
