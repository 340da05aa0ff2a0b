// 2015 edition on purpose (no `edition` key in Cargo.toml): rs-src/nbody.rs says `use rand;`.
#[macro_use]
extern crate lazy_static;
extern crate rand;

pub mod nbody {
    // the reference's file, untouched
    include!(concat!(env!("RUST_EXP_DIR"), "/rs-src/nbody.rs"));
    // the dumper is a child module of it, so it sees PARTICLES, Particle and force()
    #[cfg(test)]
    include!("../../nbody_vectors.rs");
}
