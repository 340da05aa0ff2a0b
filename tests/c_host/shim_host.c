/* The Rust shim of INTEGRATION.md option A, restated in C so that it can be linked and run here (no Rust toolchain
 * in this image): a host-side module that ITSELF defines the six reference symbols nb_* (as rust-shim/nbody.rs does
 * with #[no_mangle]) and forwards each to the b200_-prefixed implementation in the STATIC archive libnbody_b200.a --
 * the reference's link contract (one static archive in the Haskell binary, Cargo.toml:5-7, rust-exp.cabal:46-47).
 * Linking this must not produce duplicate definitions of nb_*: nb_alias.o stays in the archive.
 * main() then plays the Haskell host (hs-src/RustNBodyExperiment.hs:42-62) through the shim's symbols. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "nbody_b200.h" /* only for the b200_nb_* prototypes; the nb_* below are OUR definitions */

int32_t nb_num_particles(void) { return b200_nb_num_particles(); }
void nb_random_disk(int32_t n) { b200_nb_random_disk(n); }
void nb_stable_orbits(int32_t n, float rmin, float rmax) { b200_nb_stable_orbits(n, rmin, rmax); }
void nb_step_brute_force(float dt) { b200_nb_step_brute_force(dt); }
void nb_step_barnes_hut(float theta, float dt, int32_t nthreads) { b200_nb_step_barnes_hut(theta, dt, nthreads); }
void nb_draw(int32_t w, int32_t h, uint32_t *fb) { b200_nb_draw(w, h, fb); }

static double now_ms(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec * 1e3 + t.tv_nsec * 1e-6;
}

int main(void)
{
    const int w = 320, h = 240;
    uint32_t *fb = (uint32_t *)malloc(sizeof(uint32_t) * w * h);
    nb_stable_orbits(10000, 0.5f, 30.0f);
    double step_ms = 0.0;
    for (int f = 0; f < 5; f++) {
        const double t0 = now_ms();               /* the reference host's timeIt around the step call */
        nb_step_barnes_hut(0.85f, 0.01f, 1);
        step_ms = now_ms() - t0;
        nb_draw(w, h, fb);
    }
    unsigned long long lit = 0;
    for (int i = 0; i < w * h; i++) lit += fb[i] != 0;
    const int cross_ok = fb[(h / 2) * w + w / 2] == 0x00FF00FFu;
    printf("{\"particles\": %d, \"lit_pixels\": %llu, \"cross_ok\": %d, \"last_step_wall_ms\": %.4f}\n", nb_num_particles(), lit,
           cross_ok, step_ms);
    free(fb);
    return (cross_ok && lit > 1000 && nb_num_particles() == 10000) ? 0 : 2;
}
