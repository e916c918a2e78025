// tpt_scene_programs.h -- a TEST scene family: random "programs" over the scene classes.
//
// "program:<seed>" (and "programm:<seed>", the same plus participating media; "programL:<seed>" / "programLm:<seed>",
// several times as many objects; "programp:<seed>", with Perlin marble textures) builds a hitable tree out of everything the class API offers -- spheres, moving spheres, the three
// rects with and without flip_normal, boxes, translate / rotate_y wrappers around primitives AND around groups,
// hitable_lists and bvh_nodes nested in each other (also bvh_nodes of one element, and bvh_nodes under lists), all
// four surface materials, checker textures -- inside a Cornell-sized room with the lamp where the reference's
// hard-coded light list expects it (main.cpp:99-106).
//
// The header names the classes unqualified and is compiled twice: against this front end's classes
// (host/tpt_scene.h, scene "program:<seed>" of tpt_host_build_scene) and, by the test harness, against the
// reference's own headers (ref_harness.cc in the checkers' directory; the dependency points
// from there to here, never the other way). Both then build the same tree from a seed, and the parity checks
// can leave the handful of fixed scenes: reference vs restatement on the CPU, CUDA vs restatement on the GPU.
// It draws from its own generator; bvh_node's constructor keeps taking its split axes from the thread's
// default-seeded drand_r stream on both sides, in the same order.
#pragma once
#include <cstdint>

namespace scene_programs {

struct rng {
  uint32_t s;
  explicit rng(uint32_t seed) : s(seed * 2654435761u + 12345u) {}
  uint32_t next() {
    s = s * 1664525u + 1013904223u;
    return s >> 8;
  }
  float u() { return (float)next() * (1.0f / 16777216.0f); }
  int below(int n) {
    int k = (int)(u() * (float)n);
    return k < n ? k : n - 1;
  }
  float range(float a, float b) { return a + (b - a) * u(); }
};

// "programp:<seed>": every third texture is the marble of src/texture.cc:18-25 (Perlin turbulence at world coordinates of
// several hundred, on moved and rotated objects). The noise tables are static and re-randomised by every perlin_noise
// constructor from the wall clock: the test reads them back from the reference-side build and forces them on the other.
static bool g_with_perlin = false;

inline texture *any_texture(rng &g) {
  vec3 c(g.range(0.05f, 0.95f), g.range(0.05f, 0.95f), g.range(0.05f, 0.95f));
  if (g_with_perlin && g.below(3) == 0) return new perlin_noise_texture(g.range(0.01f, 0.1f));
  if (g.below(5) == 0) return new checker_texture(new constant_texture(c), new constant_texture(vec3(0.9f, 0.9f, 0.9f)));
  return new constant_texture(c);
}

inline material *any_material(rng &g) {
  switch (g.below(6)) {
  case 0: return new metal(vec3(g.range(0.4f, 1.0f), g.range(0.4f, 1.0f), g.range(0.4f, 1.0f)), g.below(2) ? 0.0f : g.range(0.0f, 0.6f));
  case 1: return new dielectric(g.range(1.2f, 1.8f));
  default: return new lambertian(any_texture(g));
  }
}

inline vec3 any_point(rng &g) { return vec3(g.range(60.f, 495.f), g.range(30.f, 420.f), g.range(60.f, 495.f)); }

inline hitable *any_primitive(rng &g) {
  const vec3 c = any_point(g);
  const float r = g.range(15.f, 70.f);
  switch (g.below(9)) {
  case 0:
  case 1: return new sphere(c, r, any_material(g));
  case 2: return new moving_sphere(c, c + vec3(g.range(-40.f, 40.f), g.range(-40.f, 40.f), g.range(-40.f, 40.f)), 0.0f, 1.0f, r, any_material(g));
  case 3: {
    hitable *q = new xy_rect(c.x() - r, c.x() + r, c.y() - r, c.y() + 0.5f * r, c.z(), any_material(g));
    return g.below(2) ? (hitable *)new flip_normal(q) : q;
  }
  case 4: {
    hitable *q = new xz_rect(c.x() - r, c.x() + r, c.z() - 0.7f * r, c.z() + r, c.y(), any_material(g));
    return g.below(2) ? (hitable *)new flip_normal(q) : q;
  }
  case 5: {
    hitable *q = new yz_rect(c.y() - r, c.y() + r, c.z() - r, c.z() + r, c.x(), any_material(g));
    return g.below(2) ? (hitable *)new flip_normal(q) : q;
  }
  case 6: return new box(c - vec3(r, 0.6f * r, 0.8f * r), c + vec3(r, 0.6f * r, 0.8f * r), any_material(g));
  case 7: // the Cornell idiom: a box at the origin, rotated, then moved into place (src/utils.cc:305-312)
    return new translate(new rotate_y(new box(vec3(0, 0, 0), vec3(2 * r, 3 * r, 2 * r), any_material(g)), g.range(-40.f, 40.f)), c - vec3(r, 0, r));
  default: return new translate(new sphere(vec3(0, 0, 0), r, any_material(g)), c);
  }
}

inline hitable *any_group(rng &g, int depth, bool under_list);

inline hitable *any_member(rng &g, int depth, bool under_list) {
  if (depth <= 0 || g.below(3) != 0) return any_primitive(g);
  return any_group(g, depth - 1, under_list);
}

inline hitable *any_group(rng &g, int depth, bool under_list) {
  const int kind = g.below(under_list ? 5 : 4);
  if (kind <= 1) { // bvh_node of 1..7 members (one member: left_ == right_)
    const int n = 1 + g.below(7);
    hitable **l = new hitable *[n];
    for (int i = 0; i < n; i++) l[i] = any_member(g, depth, under_list);
    return new bvh_node(l, n, 0.0f, 1.0f);
  }
  if (kind == 2) { // hitable_list of 1..5 members
    const int n = 1 + g.below(5);
    hitable **l = new hitable *[n];
    for (int i = 0; i < n; i++) l[i] = any_member(g, depth, true);
    return new hitable_list(l, n);
  }
  if (kind == 3) { // a whole group under a transform, centred first so that the rotation keeps it in the room
    hitable *inner = any_group(g, depth > 0 ? depth - 1 : 0, under_list);
    return new translate(new rotate_y(new translate(inner, vec3(-278.f, 0.f, -278.f)), g.range(-30.f, 30.f)), vec3(278.f, g.range(-20.f, 20.f), 278.f));
  }
  return any_primitive(g);
}

// the room, the reference's lamp, and a few top-level groups; root = hitable_list or bvh_node
inline hitable *build(uint32_t seed, bool with_media = false, bool large = false, bool with_perlin = false) {
  rng g(seed);
  g_with_perlin = with_perlin;
  hitable **l = new hitable *[48];
  int n = 0;
  l[n++] = new flip_normal(new xz_rect(213, 343, 227, 332, 554, new diffuse_light(new constant_texture(vec3(15, 15, 15)))));
  { // the Cornell room (src/utils.cc:287-303): floor, ceiling and back wall always -- an open room stays nearly black under
    // the reference's black background --, the coloured side walls in half of the programs each
    material *white = new lambertian(new constant_texture(vec3(0.73f, 0.73f, 0.73f)));
    l[n++] = new xz_rect(0, 555, 0, 555, 0, white);
    l[n++] = new flip_normal(new xz_rect(0, 555, 0, 555, 555, white));
    l[n++] = new flip_normal(new xy_rect(0, 555, 0, 555, 555, white));
    if (g.below(2)) l[n++] = new flip_normal(new yz_rect(0, 555, 0, 555, 555, new lambertian(any_texture(g))));
    // (a constant colour on the wall in the plane x = 0: a checker_texture there is sin(10 x) sin(10 y) sin(10 z) AT a zero of
    // its first factor, i.e. the sign of the rounding residue of o.x + t d.x -- exactly 0 in 89 % of the reference's own
    // evaluations, never with a fused multiply-add. FAST mode cannot and need not reproduce that)
    if (g.below(2)) l[n++] = new yz_rect(0, 555, 0, 555, 0, new lambertian(new constant_texture(vec3(g.range(0.05f, 0.95f), g.range(0.05f, 0.95f), g.range(0.05f, 0.95f)))));
  }
  // "programL:<seed>": enough primitives (60-400) for the large-scene code paths (SAH BVH, skip-pointer / replay walks)
  const int groups = large ? 8 + g.below(24) : 1 + g.below(4);
  for (int i = 0; i < groups; i++) l[n++] = large ? any_group(g, 3, false) : any_member(g, 3, false);
  if (with_media) { // "programm:<seed>": one or two participating media (src/hitable.cc:92-128) with a sphere, a box or a moved box as boundary
    const int media = 1 + g.below(2);
    for (int i = 0; i < media; i++) {
      const vec3 c = any_point(g);
      const float r = g.range(40.f, 110.f);
      hitable *boundary;
      switch (g.below(3)) {
      case 0: boundary = new sphere(c, r, new dielectric(1.5f)); break;
      case 1: boundary = new box(c - vec3(r, r, r), c + vec3(r, r, r), new dielectric(1.5f)); break;
      default: boundary = new translate(new rotate_y(new box(vec3(0, 0, 0), vec3(2 * r, 2 * r, 2 * r), new dielectric(1.5f)), g.range(-30.f, 30.f)), c - vec3(r, r, r)); break;
      }
      l[n++] = new constant_medium(boundary, g.range(0.002f, 0.02f), new constant_texture(vec3(g.range(0.1f, 1.f), g.range(0.1f, 1.f), g.range(0.1f, 1.f))));
    }
  }
  if (g.below(2)) return new bvh_node(l, n, 0.0f, 1.0f);
  return new hitable_list(l, n);
}

} // namespace scene_programs
