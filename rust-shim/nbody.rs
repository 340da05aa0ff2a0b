// rust-shim/nbody.rs -- replacement for rs-src/nbody.rs of blitzcode/rust-exp.
//
// Six forwarders, no logic, no state: the particle set lives in GPU memory inside libnbody_b200.so
// (include/nbody_b200.h).  NOT COMPILED OR TESTED in the build environment of this repository (no Rust
// toolchain there); it is kept this thin so that not compiling it is low-risk.  See INTEGRATION.md.
//
// The library exports the same unmangled names this module must export, so it is bound at run time
// (dlopen/dlsym through the `libloading` crate) instead of at link time, which would define each
// symbol twice.  With INTEGRATION.md "Option B" this file is not needed at all.

use libloading::{Library, Symbol};
use std::os::raw::{c_float, c_int};

lazy_static! {
    static ref LIB: Library = unsafe {
        Library::new(std::env::var("NBODY_B200_LIB").unwrap_or("libnbody_b200.so".to_string()))
            .expect("libnbody_b200.so not found (no CPU fallback)")
    };
}

macro_rules! fwd {
    ($name:expr, $ty:ty) => {{
        let f: Symbol<$ty> = unsafe { LIB.get($name).expect("symbol missing in libnbody_b200.so") };
        f
    }};
}

#[no_mangle]
pub extern fn nb_num_particles() -> i32 {
    unsafe { fwd!(b"nb_num_particles\0", unsafe extern "C" fn() -> c_int)() }
}

#[no_mangle]
pub extern fn nb_random_disk(num_particles: i32) -> () {
    unsafe { fwd!(b"nb_random_disk\0", unsafe extern "C" fn(c_int))(num_particles) }
}

#[no_mangle]
pub extern fn nb_stable_orbits(num_particles: i32, rmin: f32, rmax: f32) -> () {
    unsafe { fwd!(b"nb_stable_orbits\0", unsafe extern "C" fn(c_int, c_float, c_float))(num_particles, rmin, rmax) }
}

#[no_mangle]
pub extern fn nb_step_brute_force(dt: f32) -> () {
    unsafe { fwd!(b"nb_step_brute_force\0", unsafe extern "C" fn(c_float))(dt) }
}

// NOTE argument order: theta first (rs-src/nbody.rs:187).
#[no_mangle]
pub extern fn nb_step_barnes_hut(theta: f32, dt: f32, nthreads: i32) -> () {
    unsafe { fwd!(b"nb_step_barnes_hut\0", unsafe extern "C" fn(c_float, c_float, c_int))(theta, dt, nthreads) }
}

#[no_mangle]
pub extern fn nb_draw(w: i32, h: i32, fb: *mut u32) -> () {
    unsafe { fwd!(b"nb_draw\0", unsafe extern "C" fn(c_int, c_int, *mut u32))(w, h, fb) }
}
