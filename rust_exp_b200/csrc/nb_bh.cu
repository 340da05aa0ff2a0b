// nb_bh.cu -- Barnes-Hut step (rs-src/nbody.rs:186-480).  Placeholder until the device tree lands.
#include "nb_engine.h"
namespace nb {
void bh_step(Engine&, float, float) { fatal("nb_step_barnes_hut: device tree not built into this library yet", __FILE__, __LINE__); }
void bh_accelerations(Engine&, float, float2*) { fatal("bh_accelerations: not built yet", __FILE__, __LINE__); }
void bh_shutdown(Engine&) {}
}  // namespace nb
