// oracle/ref_noinject.cc -- TEST INFRASTRUCTURE ONLY.
// Stubs for the plain build (libtptref.so): the reference's own drand_r (src/utils.cc:28-32,
// thread_local mt19937) is used untouched; there is no deterministic stream.
#include <cstdint>
extern "C" {
void ref_rng_begin_sample(uint32_t, uint32_t, uint32_t, uint32_t) {}
void ref_rng_next_stage(void) {}
void ref_rng_end(void) {}
int ref_rng_is_injected(void) { return 0; }
uint64_t ref_rng_draws(void) { return 0; }
}
