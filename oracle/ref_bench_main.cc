// oracle/ref_bench_main.cc -- TEST / BENCH INFRASTRUCTURE ONLY: the timing executable behind bench.py's CPU legs.
//
// oracle/_ref/ref_bench = the unmodified reference sources compiled like the reference's own Release executable
// (g++ -std=c++14 -O3 -DNDEBUG, NO -fPIC) + ref_harness.cc (the sample loop main.cpp:115-134 restated around the
// reference's own color(), a thread pool over rows) + this main(). bench.py loads the same harness as a PIC shared
// library for everything else; timed from there the reference runs ~10 % slower (thread-local drand_r state through
// __tls_get_addr, PLT calls) than the stock executable's steady state -- so the numbers reported as the CPU baseline
// come from this non-PIC build (r02 on the GPU box's 16 cores: 29.8 Mpaths/s here, 26.6 through the library, 32 for
// the stock program's slope).
//
// usage: ref_bench <scene> <nx> <ny> <spp> <depth> <fov> <cornell|book> <threads> <repeats> [image.jpg]
// prints one line per repeat: "seconds <s> paths <n> threads <t>"
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

struct light_desc {
  int kind;
  float p[5];
};
struct ref_render_args { // mirrors oracle/ref_harness.cc
  float lookfrom[3], lookat[3], vup[3];
  float vfov, aspect, aperture, focus_dist, t0, t1;
  int32_t nx, ny, ns, max_depth;
  int32_t slices;
  int32_t n_lights;
  light_desc lights[8];
  int32_t deterministic;
  uint32_t seed_lo, seed_hi;
  int32_t threads;
  int32_t count_rays;
  int32_t x0, y0, x1, y1;
};
struct ref_render_stats {
  uint64_t paths, rays, draws;
  double seconds;
  int32_t threads;
};
extern "C" {
void *ref_scene_create(const char *name, const unsigned char *img, int iw, int ih);
int ref_render(void *s, const ref_render_args *a, float *out_sum, float *out_samples, ref_render_stats *st);
unsigned char *ref_load_image(const char *path, int *w, int *h, int *ch);
}

int main(int argc, char **argv) {
  if (argc < 10) {
    std::fprintf(stderr, "usage: ref_bench scene nx ny spp depth fov cornell|book threads repeats [image]\n");
    return 2;
  }
  const std::string scene = argv[1], camera = argv[7];
  const int nx = std::atoi(argv[2]), ny = std::atoi(argv[3]), spp = std::atoi(argv[4]), depth = std::atoi(argv[5]);
  const float fov = (float)std::atof(argv[6]);
  const int threads = std::atoi(argv[8]), repeats = std::atoi(argv[9]);
  unsigned char *img = nullptr;
  int iw = 0, ih = 0, ch = 0;
  if (argc > 10) {
    img = ref_load_image(argv[10], &iw, &ih, &ch);
    if (!img || ch != 3) {
      std::fprintf(stderr, "cannot load %s as a 3-channel image\n", argv[10]);
      return 2;
    }
  }
  void *s = ref_scene_create(scene.c_str(), img, iw, ih);
  if (!s) {
    std::fprintf(stderr, "unknown scene %s\n", scene.c_str());
    return 2;
  }
  ref_render_args a;
  std::memset(&a, 0, sizeof(a));
  a.vup[1] = 1.f;
  if (camera == "cornell") { // main.cpp:87-91
    a.lookfrom[2] = 800.f;
    a.focus_dist = 10.f;
  } else { // the commented alternative main.cpp:82-84
    a.lookfrom[0] = 13.f;
    a.lookfrom[1] = 2.f;
    a.lookfrom[2] = 3.f;
    a.focus_dist = std::sqrt(182.0f);
  }
  a.vfov = fov;
  a.aspect = (float)nx / (float)ny;
  a.aperture = 0.1f;
  a.nx = nx;
  a.ny = ny;
  a.ns = spp;
  a.max_depth = depth;
  a.slices = 1;
  a.n_lights = 2; // main.cpp:99-106
  a.lights[0] = light_desc{0, {-100.f, 100.f, -150.f, -50.f, 298.f}};
  a.lights[1] = light_desc{1, {120.f, -50.f, 40.f, 120.f, 0.f}};
  a.threads = threads;
  a.x1 = nx;
  a.y1 = ny;
  std::vector<float> out((size_t)nx * ny * 3);
  for (int r = 0; r < repeats; r++) {
    ref_render_stats st;
    if (ref_render(s, &a, out.data(), nullptr, &st) != 0) return 3;
    std::printf("seconds %.6f paths %llu threads %d\n", st.seconds, (unsigned long long)st.paths, st.threads);
    std::fflush(stdout);
  }
  return 0;
}
