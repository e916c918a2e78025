// oracle/ref_inject.cc -- TEST INFRASTRUCTURE ONLY.
//
// Strong definition of the reference's `double drand_r(double,double)` (declared
// headers/utils.h:17, defined src/utils.cc:28-32) for libtptref_det.so. oracle/Makefile weakens
// the original definition inside utils.o with `objcopy --weaken-symbol`, so every call site in
// the unmodified reference objects binds to this one.
//
//  * no sample context active  -> identical behaviour to the original: a thread_local
//    default-seeded std::mt19937 through a fresh uniform_real_distribution<double>(min,max)
//    (so scene construction -- random_scene(), bvh_node axis choice -- is unchanged);
//  * sample context active     -> the k-th draw of stage s of (pixel, sample) is
//        u = (philox4x32_10(ctr = {pixel, sample, stage, k/4}, key = {seed_lo, seed_hi})[k%4] >> 8) * 2^-24
//    a float-exact 24-bit uniform in [0,1), so host and device consume identical values.
//    stage 0 = camera (pixel jitter + lens + shutter, main.cpp:121-124); stage d+1 = the draws
//    made by color() at recursion depth d (advanced by the world->hit decorator in ref_harness.cc).
#include <cstdint>
#include <random>

namespace {
struct rng_ctx {
  bool active = false;
  uint32_t key[2] = {0, 0};
  uint32_t pixel = 0, sample = 0, stage = 0, ndraw = 0;
  uint32_t block_id = 0xffffffffu;
  uint32_t block[4];
  uint64_t total = 0;
};
thread_local rng_ctx g;

inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
} // namespace

double drand_r(double min, double max) {
  if (!g.active) {
    static thread_local std::mt19937 generator; // as src/utils.cc:29
    std::uniform_real_distribution<double> dis(min, max);
    return dis(generator);
  }
  uint32_t blk = g.ndraw >> 2;
  if (blk != g.block_id) {
    uint32_t ctr[4] = {g.pixel, g.sample, g.stage, blk};
    philox4x32_10(ctr, g.key, g.block);
    g.block_id = blk;
  }
  uint32_t x = g.block[g.ndraw & 3];
  g.ndraw++;
  g.total++;
  double u = (double)(x >> 8) * (1.0 / 16777216.0);
  return min + (max - min) * u; // render-time callers all use (0,1): exact
}

extern "C" {
void ref_rng_begin_sample(uint32_t seed_lo, uint32_t seed_hi, uint32_t pixel, uint32_t sample) {
  g.active = true;
  g.key[0] = seed_lo;
  g.key[1] = seed_hi;
  g.pixel = pixel;
  g.sample = sample;
  g.stage = 0;
  g.ndraw = 0;
  g.block_id = 0xffffffffu;
}
void ref_rng_next_stage(void) {
  if (!g.active) return;
  g.stage++;
  g.ndraw = 0;
  g.block_id = 0xffffffffu;
}
void ref_rng_end(void) { g.active = false; }
int ref_rng_is_injected(void) { return 1; }
uint64_t ref_rng_draws(void) { return g.total; }
void ref_philox4x32_10(const uint32_t *ctr, const uint32_t *key, uint32_t *out) {
  philox4x32_10(ctr, key, out);
}
}
