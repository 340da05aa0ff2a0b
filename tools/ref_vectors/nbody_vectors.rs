// tools/ref_vectors/nbody_vectors.rs -- reference-vector dumper for blitzcode/rust-exp rs-src/nbody.rs.
//
// PURPOSE.  The reference ships no tests or golden vectors for the N-body path and cannot be built where
// libnbody_b200 is developed (no Rust toolchain there), so the CPU oracle of that repository is "parity unpinned".
// A maintainer with cargo closes that gap in two minutes, WITHOUT modifying the reference:
//
//   cd tools/ref_vectors && python make_inputs.py                    # writes inputs/*.bin (seeded, deterministic)
//   cd crate && RUST_EXP_DIR=/abs/path/to/rust-exp NB_VECTORS_DIR=/abs/path/to/nbody_b200/tools/ref_vectors \
//       cargo test --release -- --nocapture                         # writes ../outputs/*.bin
//   cp ../outputs/*.bin ../../../tests/golden/reference_vectors/ && git add ... && git commit
//
// crate/src/lib.rs include!s the reference's rs-src/nbody.rs byte for byte and this file as a child module of it (a
// child module sees its private items: PARTICLES, Particle, force()).  tests/test_reference_vectors.py then checks the
// oracle (CPU) and the EXACT GPU mode against the vectors bit for bit.  Nothing here changes the reference's
// arithmetic: it only sets the particle vector, calls the reference's own nb_step_* entry points and writes the
// resulting particle vector.  Old-rustc friendly (no to_bits/from_bits, no glob import of private items).
mod b200_vectors {
    use super::{force, nb_step_barnes_hut, nb_step_brute_force, Particle, PARTICLES};
    use std::fs::File;
    use std::io::{Read, Write};

    fn dir() -> String { std::env::var("NB_VECTORS_DIR").expect("set NB_VECTORS_DIR to .../tools/ref_vectors") }

    fn load(name: &str) -> Vec<Particle> {
        let mut bytes = Vec::new();
        File::open(format!("{}/inputs/{}.bin", dir(), name)).expect("run make_inputs.py first").read_to_end(&mut bytes).unwrap();
        assert!(bytes.len() % 20 == 0);
        let mut out = Vec::new();
        for rec in bytes.chunks(20) {
            let f = |k: usize| -> f32 {
                let b = [rec[4 * k], rec[4 * k + 1], rec[4 * k + 2], rec[4 * k + 3]];
                let u = (b[0] as u32) | ((b[1] as u32) << 8) | ((b[2] as u32) << 16) | ((b[3] as u32) << 24);
                unsafe { std::mem::transmute::<u32, f32>(u) }
            };
            out.push(Particle { px: f(0), py: f(1), vx: f(2), vy: f(3), m: f(4) });
        }
        out
    }

    fn put(w: &mut File, v: f32) {
        let u = unsafe { std::mem::transmute::<f32, u32>(v) };
        w.write_all(&[u as u8, (u >> 8) as u8, (u >> 16) as u8, (u >> 24) as u8]).unwrap();
    }

    fn store(name: &str) {
        std::fs::create_dir_all(format!("{}/outputs", dir())).unwrap();
        let mut w = File::create(format!("{}/outputs/{}.bin", dir(), name)).unwrap();
        let particles = PARTICLES.lock().unwrap();
        for p in particles.iter() { put(&mut w, p.px); put(&mut w, p.py); put(&mut w, p.vx); put(&mut w, p.vy); put(&mut w, p.m); }
    }

    fn set(ps: Vec<Particle>) {
        let mut particles = PARTICLES.lock().unwrap();
        particles.clear();
        for p in ps { particles.push(p); }
    }

    // (input file, output name, theta (< 0: nb_step_brute_force), dt, steps, nthreads) -- keep in sync with make_inputs.py
    const CASES: &'static [(&'static str, &'static str, f32, f32, u32, i32)] = &[
        ("kat1_two_bodies", "kat1_two_bodies_brute_dt001_k1", -1.0, 0.01, 1, 1),
        ("disk_1024", "disk_1024_brute_dt001_k100", -1.0, 0.01, 100, 1),
        ("orbits_1024", "orbits_1024_brute_dt001_k100", -1.0, 0.01, 100, 1),
        ("plummer_8192", "plummer_8192_brute_dt001_k10", -1.0, 0.01, 10, 1),
        ("disk_1024", "disk_1024_bh_t05_dt001_k20", 0.5, 0.01, 20, 1),
        ("disk_1024", "disk_1024_bh_t0_dt001_k3", 0.0, 0.01, 3, 3),
        ("orbits_10000", "orbits_10000_bh_t085_dt001_k10", 0.85, 0.01, 10, 4),
        ("disk_16384", "disk_16384_bh_t05_dt001_k10", 0.5, 0.01, 10, 7),
        ("plummer_16384", "plummer_16384_bh_t075_dt001_k5", 0.75, 0.01, 5, 2),
        ("merge_500", "merge_500_bh_t05_dt001_k4", 0.5, 0.01, 4, 1),
        ("kill_4", "kill_4_bh_t05_dt0001_k1", 0.5, 0.001, 1, 1),
        ("disk_65536", "disk_65536_bh_t05_dt001_k2", 0.5, 0.01, 2, 8),
    ];

    #[test]
    fn b200_vectors() {
        for &(input, output, theta, dt, steps, nthreads) in CASES {
            set(load(input));
            for _ in 0..steps {
                if theta < 0.0 { nb_step_brute_force(dt); } else { nb_step_barnes_hut(theta, dt, nthreads); }
            }
            store(output);
            println!("b200_vectors: wrote {}", output);
        }
        // the pair law itself (rs-src/nbody.rs:164-184): a few hundred evaluations on awkward operands
        let mut w = File::create(format!("{}/outputs/force_kat.bin", dir())).unwrap();
        let ps = load("force_kat_pairs");
        let mut i = 0;
        while i + 1 < ps.len() {
            let (a, b) = (&ps[i], &ps[i + 1]);
            let (fx, fy) = force(a.px, a.py, a.m, b.px, b.py, b.m);
            put(&mut w, fx); put(&mut w, fy);
            i += 2;
        }
    }
}
