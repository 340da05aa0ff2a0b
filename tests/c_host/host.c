/* A plain-C host that drives libnbody_b200.so exactly the way the reference's Haskell host drives
 * rs-src/nbody.rs through its FFI (hs-src/RustNBodyExperiment.hs:42-62): nb_stable_orbits(10000, 0.5, 30),
 * then per frame nb_step_barnes_hut(theta, dt, nthreads) followed by nb_draw(w, h, fb).
 * Only the six reference symbols are used -- no nbx_*, no Python, no torch: this is the drop-in boundary.
 * Build: gcc -O2 -I include tests/c_host/host.c -L rust_exp_b200 -lnbody_b200 -Wl,-rpath,$PWD/rust_exp_b200 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "nbody_b200.h"

int main(int argc, char **argv)
{
    const int frames = argc > 1 ? atoi(argv[1]) : 5;
    const int w = 320, h = 240;
    uint32_t *fb = (uint32_t *)malloc(sizeof(uint32_t) * w * h);
    nb_stable_orbits(10000, 0.5f, 30.0f);             /* withExperiment */
    if (nb_num_particles() != 10000) { fprintf(stderr, "unexpected particle count\n"); return 1; }
    for (int f = 0; f < frames; f++) {
        nb_step_barnes_hut(0.85f, 0.01f, 1);           /* theta FIRST */
        nb_draw(w, h, fb);
    }
    nb_step_barnes_hut(0.0f, 0.01f, 1);               /* theta == 0 -> brute force path */
    nb_step_brute_force(0.01f);
    nb_draw(w, h, fb);
    unsigned long long lit = 0, sum = 0;
    for (int i = 0; i < w * h; i++) { lit += fb[i] != 0; sum += fb[i]; }
    const int cross_ok = fb[(h / 2) * w + w / 2] == 0x00FF00FFu && fb[(h / 2) * w + w / 2 + 1] == 0x00FF00FFu;
    nb_random_disk(2000);
    printf("{\"particles\": %d, \"lit_pixels\": %llu, \"checksum\": %llu, \"cross_ok\": %d, \"after_random_disk\": %d}\n",
           10000, lit, sum, cross_ok, nb_num_particles());
    free(fb);
    return (cross_ok && lit > 1000 && nb_num_particles() == 2000) ? 0 : 2;
}
