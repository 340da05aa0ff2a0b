// nb_alias.cu -- the reference's unprefixed symbol names (rs-src/nbody.rs:34-35,39-40,73-74,106-107,186-187,482-483
// as imported by hs-src/RustNBodyExperiment.hs:101-106), forwarding to the b200_-prefixed implementations of nb_abi.cu.
// Kept in its own object file on purpose: inside libnbody_b200.a it is only pulled in when nobody else defines nb_*
// (direct drop-in, INTEGRATION.md option B); with the Rust shim of option A, which defines nb_* itself, it stays out.
#include "../../include/nbody_b200.h"

extern "C" {
int32_t nb_num_particles(void) { return b200_nb_num_particles(); }
void nb_random_disk(int32_t n) { b200_nb_random_disk(n); }
void nb_stable_orbits(int32_t n, float rmin, float rmax) { b200_nb_stable_orbits(n, rmin, rmax); }
void nb_step_brute_force(float dt) { b200_nb_step_brute_force(dt); }
void nb_step_barnes_hut(float theta, float dt, int32_t nthreads) { b200_nb_step_barnes_hut(theta, dt, nthreads); }
void nb_draw(int32_t w, int32_t h, uint32_t* fb) { b200_nb_draw(w, h, fb); }
void nb_set_particles(const float* aos5, int32_t n) { b200_nb_set_particles(aos5, n); }
void nb_get_particles(float* aos5_out, int32_t n) { b200_nb_get_particles(aos5_out, n); }
}
