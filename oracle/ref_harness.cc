// oracle/ref_harness.cc -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A thin extern "C" harness compiled TOGETHER WITH THE UNMODIFIED REFERENCE SOURCES
// (/root/reference/src/*.cc, headers/*.h) into oracle/_ref/libtptref.so by oracle/Makefile.
// It adds no arithmetic of its own on the hot path: every hit record, every radiance sample
// comes out of the reference's own `hitable::hit` / `color()` / `camera::get_ray`.
//
// What it restates (because main.cpp is monolithic and cannot be linked):
//   * the per-pixel sample loop  main.cpp:115-134  (ref_render)
//   * the hard-coded light list  main.cpp:99-106   (passed in as data)
// What it decorates (public pointers only, reference code untouched):
//   * every geometric leaf is wrapped in `tagged_leaf` so a hit reports the DFS leaf index
//     (hit_record carries no id: headers/hitable.h:14-21)
//   * `world` is wrapped in `stage_counter` so the deterministic RNG stream can be keyed on
//     (pixel, sample, stage) where stage = number of world->hit calls so far in the sample
//     (color() calls world->hit exactly once per invocation: src/utils.cc:61).
//
// Deterministic stream: in libtptref_det.so the reference's `drand_r` (src/utils.cc:28-32) is
// replaced at link time (objcopy --weaken-symbol + strong definition in ref_inject.cc) by a
// Philox4x32-10 counter-based generator; see ref_inject.cc.
#include "camera.h"
#include "hitable.h"
#include "hitable_list.h"
#include "material.h"
#include "perlin_noise.h"
#include "rect_box.h"
#include "sphere.h"
#include "texture.h"
#include "utils.h"

#include "../tiny-path-tracer_b200/host/tpt_scene_programs.h" // test scene family, compiled here against the reference's classes

#include <atomic>
#include <chrono>
#include <memory>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <thread>
#include <vector>

// ---- deterministic-stream context (defined in ref_inject.cc for the _det build, stubbed in the
// plain build by ref_noinject.cc) --------------------------------------------------------------
extern "C" {
void ref_rng_begin_sample(uint32_t seed_lo, uint32_t seed_hi, uint32_t pixel, uint32_t sample);
void ref_rng_next_stage(void);
void ref_rng_end(void);
int ref_rng_is_injected(void);
uint64_t ref_rng_draws(void);
}

namespace {

struct tag_info {
  int prim_id;
  material *real_mat;
};

// Decorator around one geometric leaf (sphere / moving_sphere / x?_rect). On a successful hit it
// overwrites rec.mat_ptr with a pointer to its tag so the id travels with the record through
// bvh_node / hitable_list / flip_normal / translate / rotate_y untouched.
class tagged_leaf : public hitable {
public:
  tagged_leaf(hitable *inner, int id) : inner_(inner) {
    tag_.prim_id = id;
    tag_.real_mat = nullptr;
  }
  bool hit(const ray &r, float t_min, float t_max, hit_record &rec) const override {
    if (inner_->hit(r, t_min, t_max, rec)) {
      tag_.real_mat = rec.mat_ptr;
      rec.mat_ptr = reinterpret_cast<material *>(&tag_);
      return true;
    }
    return false;
  }
  bool bounding_box(float t0, float t1, AABB &b) const override {
    return inner_->bounding_box(t0, t1, b);
  }
  hitable *inner_;
  mutable tag_info tag_;
};

thread_local uint64_t g_rays = 0;

// Decorator around `world`: counts rays and advances the RNG stage.
class stage_counter : public hitable {
public:
  explicit stage_counter(hitable *w) : w_(w) {}
  bool hit(const ray &r, float t_min, float t_max, hit_record &rec) const override {
    ++g_rays;
    ref_rng_next_stage();
    return w_->hit(r, t_min, t_max, rec);
  }
  bool bounding_box(float t0, float t1, AABB &b) const override {
    return w_->bounding_box(t0, t1, b);
  }
  hitable *w_;
};

struct ref_scene {
  hitable *world = nullptr;        // untouched tree (for color())
  hitable *world_tagged = nullptr; // second instance with tagged leaves (for hit batches)
  int n_leaves = 0;
  std::vector<material *> mats; // material id = index of first DFS encounter
  unsigned char *image = nullptr;
};

bool is_leaf(hitable *h) {
  return dynamic_cast<sphere *>(h) || dynamic_cast<moving_sphere *>(h) ||
         dynamic_cast<xy_rect *>(h) || dynamic_cast<xz_rect *>(h) ||
         dynamic_cast<yz_rect *>(h);
}

material *leaf_material(hitable *h) {
  if (auto *s = dynamic_cast<sphere *>(h)) return s->mat_ptr_;
  if (auto *s = dynamic_cast<moving_sphere *>(h)) return s->mat_ptr_;
  if (auto *s = dynamic_cast<xy_rect *>(h)) return s->mat_ptr_;
  if (auto *s = dynamic_cast<xz_rect *>(h)) return s->mat_ptr_;
  if (auto *s = dynamic_cast<yz_rect *>(h)) return s->mat_ptr_;
  return nullptr;
}

// Pre-order DFS over the public pointers of the reference classes; wraps leaves in place.
// Same enumeration order as the product's flattener (left before right, list order, box faces
// in list order src/rect_box.cc:100-112); a leaf reachable twice (bvh n==1: left_==right_,
// src/hitable.cc:41-42) is wrapped once.
struct tagger {
  ref_scene *sc;
  std::map<hitable *, hitable *> done;
  hitable *visit(hitable *h) {
    auto it = done.find(h);
    if (it != done.end()) return it->second;
    hitable *out = h;
    if (is_leaf(h)) {
      material *m = leaf_material(h);
      bool seen = false;
      for (auto *x : sc->mats) seen = seen || (x == m);
      if (!seen) sc->mats.push_back(m);
      out = new tagged_leaf(h, sc->n_leaves++);
    } else if (auto *b = dynamic_cast<bvh_node *>(h)) {
      hitable *l = b->left_, *r = b->right_;
      b->left_ = visit(l);
      b->right_ = (r == l) ? b->left_ : visit(r);
    } else if (auto *l = dynamic_cast<hitable_list *>(h)) {
      for (int i = 0; i < l->list_size_; i++) l->list_[i] = visit(l->list_[i]);
    } else if (auto *bx = dynamic_cast<box *>(h)) {
      bx->list_ptr_ = visit(bx->list_ptr_);
    } else if (auto *t = dynamic_cast<translate *>(h)) {
      t->ptr_ = visit(t->ptr_);
    } else if (auto *ro = dynamic_cast<rotate_y *>(h)) {
      ro->ptr_ = visit(ro->ptr_);
    } else if (auto *f = dynamic_cast<flip_normal *>(h)) {
      f->ptr_ = visit(f->ptr_);
    }
    done[h] = out;
    return out;
  }
};

hitable *build_named(const std::string &name, ref_scene *sc, const unsigned char *img, int iw,
                     int ih) {
  if (name == "cornell_box") return cornell_box();               // src/utils.cc:287
  if (name == "sphere_cornell_box") return sphere_cornell_box(); // src/utils.cc:257
  if (name == "cornell_box_smoke") return cornell_box_smoke();   // src/utils.cc:321 (constant_medium)
  if (name == "oneweek_final") return oneweek_final();           // src/utils.cc:359 (stbi_load("earthmap.jpg") from the CWD)
  if (name == "random_scene") return random_scene();             // src/utils.cc:96 (bvh)
  if (name == "random_scene_list") {
    // the "no BVH" alternative is the commented line src/utils.cc:139: same leaves, flat list.
    // Collect the leaves of the bvh in DFS order? No: the list order of the commented line is the
    // construction order, which the bvh ctor has since permuted in place (std::sort on `list`,
    // src/hitable.cc:35-39). The array itself is still reachable only through the tree, so we
    // re-collect leaves left-to-right; closest-hit results do not depend on list order except
    // on exact ties.
    hitable *root = random_scene();
    std::vector<hitable *> leaves;
    std::vector<hitable *> st{root};
    std::map<hitable *, bool> seen;
    // explicit pre-order
    struct rec_ {
      static void go(hitable *h, std::vector<hitable *> &out, std::map<hitable *, bool> &seen) {
        if (auto *b = dynamic_cast<bvh_node *>(h)) {
          go(b->left_, out, seen);
          if (b->right_ != b->left_) go(b->right_, out, seen);
        } else if (!seen[h]) {
          seen[h] = true;
          out.push_back(h);
        }
      }
    };
    rec_::go(root, leaves, seen);
    hitable **arr = new hitable *[leaves.size()];
    for (size_t i = 0; i < leaves.size(); i++) arr[i] = leaves[i];
    return new hitable_list(arr, (int)leaves.size());
  }
  if (name == "two_perlin_spheres") return two_perlin_spheres(); // src/utils.cc:227
  if (name == "light_spheres") return light_spheres();           // src/utils.cc:242
  if (name == "textured_lit") {
    // TEST-ONLY scene (not in the reference): the reference's texture classes under real light,
    // because at HEAD every textured reference scene is black (no emitter, sky commented out).
    // image-textured sphere (main.cpp:78-81) on a checker ground (src/utils.cc:99-103) lit by a
    // diffuse_light sphere and a down-facing lamp rect; perlin marble sphere (src/utils.cc:228).
    unsigned char *copy = new unsigned char[(size_t)iw * ih * 3];
    std::memcpy(copy, img, (size_t)iw * ih * 3);
    sc->image = copy;
    hitable **l = new hitable *[6];
    texture *checker = new checker_texture(new constant_texture({0.1, 0.1, 0.1}), new constant_texture({0.9, 0.9, 0.9}));
    l[0] = new sphere(vec3(0, 0, 0), 3, new lambertian(new image_texture(copy, iw, ih)));
    l[1] = new sphere(vec3(0, -1003, 0), 1000, new lambertian(checker));
    l[2] = new sphere(vec3(-3, 6, 4), 2, new diffuse_light(new constant_texture(vec3(8, 8, 8))));
    l[3] = new flip_normal(new xz_rect(-2, 2, -2, 2, 7, new diffuse_light(new constant_texture(vec3(4, 4, 4)))));
    l[4] = new sphere(vec3(5, -1, -2), 2, new lambertian(new perlin_noise_texture(2.0f)));
    return new hitable_list(l, 5);
  }
  // TEST-ONLY scene family: random programs over the reference's own classes; the generator is the header the
  // product's front end compiles against ITS classes, so both sides build the same tree from a seed
  if (name.rfind("programp:", 0) == 0) return scene_programs::build((uint32_t)std::strtoul(name.c_str() + 9, nullptr, 10), false, false, true);
  if (name.rfind("programLm:", 0) == 0) return scene_programs::build((uint32_t)std::strtoul(name.c_str() + 10, nullptr, 10), true, true);
  if (name.rfind("programL:", 0) == 0) return scene_programs::build((uint32_t)std::strtoul(name.c_str() + 9, nullptr, 10), false, true);
  if (name.rfind("programm:", 0) == 0) return scene_programs::build((uint32_t)std::strtoul(name.c_str() + 9, nullptr, 10), true);
  if (name.rfind("program:", 0) == 0) return scene_programs::build((uint32_t)std::strtoul(name.c_str() + 8, nullptr, 10));
  if (name == "earth") {
    // main.cpp:78-81 (commented alternative): sphere r=3 with image_texture(earthmap.jpg)
    unsigned char *copy = new unsigned char[(size_t)iw * ih * 3];
    std::memcpy(copy, img, (size_t)iw * ih * 3);
    sc->image = copy;
    return new sphere(vec3(0, 0, 0), 3, new lambertian(new image_texture(copy, iw, ih)));
  }
  return nullptr;
}

struct light_desc {
  int kind; // 0 = xz_rect(x0,x1,z0,z1,k)   1 = sphere(cx,cy,cz,r)   (main.cpp:99-102)
  float p[5];
};

} // namespace

extern "C" {

struct ref_hit_out {
  int32_t hit, prim, mat;
  float t, u, v, p[3], n[3];
};

struct ref_camera_out { // public fields of camera_with_blur, headers/camera.h:14-20
  float origin[3], lower_left[3], vertical[3], horizontal[3], u[3], v[3], w[3];
  float lens_radius, time0, time1;
};

// Build a named scene twice (plain + tagged), each in a FRESH std::thread so that the
// thread_local default-seeded mt19937 inside drand_r (src/utils.cc:29) starts at seed 5489 exactly
// as in the reference's main thread (main.cpp:74 is its first consumer).
void *ref_scene_create(const char *name, const unsigned char *img, int iw, int ih) {
  ref_scene *sc = new ref_scene();
  std::string nm(name);
  bool ok = true;
  std::thread t1([&] {
    sc->world = build_named(nm, sc, img, iw, ih);
    if (!sc->world) ok = false;
  });
  t1.join();
  if (!ok) {
    delete sc;
    return nullptr;
  }
  // Perlin tables are static and re-randomised by every perlin_noise ctor with a wall-clock seed
  // (src/perlin_noise.cc:3-21): snapshot them after the first build and restore after the second
  // so both instances (and the product, via ref_get_perlin) see the same tables.
  auto rv = perlin_noise::random_vec3_;
  auto px = perlin_noise::permute_x_, py = perlin_noise::permute_y_, pz = perlin_noise::permute_z_;
  std::thread t2([&] {
    ref_scene tmp;
    hitable *w = build_named(nm, &tmp, img, iw, ih);
    tagger tg;
    tg.sc = sc;
    sc->world_tagged = tg.visit(w);
  });
  t2.join();
  perlin_noise::random_vec3_ = rv;
  perlin_noise::permute_x_ = px;
  perlin_noise::permute_y_ = py;
  perlin_noise::permute_z_ = pz;
  return sc;
}

int ref_scene_num_leaves(void *s) { return static_cast<ref_scene *>(s)->n_leaves; }
int ref_scene_num_materials(void *s) { return (int)static_cast<ref_scene *>(s)->mats.size(); }

void ref_get_perlin(float *vec3s, int32_t *px, int32_t *py, int32_t *pz) {
  for (int i = 0; i < 256; i++) {
    for (int c = 0; c < 3; c++) vec3s[3 * i + c] = perlin_noise::random_vec3_[i][c];
    px[i] = perlin_noise::permute_x_[i];
    py[i] = perlin_noise::permute_y_[i];
    pz[i] = perlin_noise::permute_z_[i];
  }
}

void ref_set_perlin(const float *vec3s, const int32_t *px, const int32_t *py, const int32_t *pz) {
  for (int i = 0; i < 256; i++) {
    perlin_noise::random_vec3_[i] = vec3(vec3s[3 * i], vec3s[3 * i + 1], vec3s[3 * i + 2]);
    perlin_noise::permute_x_[i] = px[i];
    perlin_noise::permute_y_[i] = py[i];
    perlin_noise::permute_z_[i] = pz[i];
  }
}

// world->hit(r, tmin, tmax, rec) on a batch (headers/hitable.h:32-33). rays = n x 7 floats
// (origin, direction, time: headers/ray.h:15-17).
void ref_hit_batch(void *s, const float *rays, int n, float tmin, float tmax, ref_hit_out *out) {
  ref_scene *sc = static_cast<ref_scene *>(s);
  for (int i = 0; i < n; i++) {
    const float *q = rays + 7 * (size_t)i;
    ray r(vec3(q[0], q[1], q[2]), vec3(q[3], q[4], q[5]), q[6]);
    hit_record rec;
    ref_hit_out &o = out[i];
    std::memset(&o, 0, sizeof(o));
    o.prim = -1;
    o.mat = -1;
    if (sc->world_tagged->hit(r, tmin, tmax, rec)) {
      tag_info *tg = reinterpret_cast<tag_info *>(rec.mat_ptr);
      o.hit = 1;
      o.prim = tg->prim_id;
      for (size_t m = 0; m < sc->mats.size(); m++)
        if (sc->mats[m] == tg->real_mat) o.mat = (int)m;
      o.t = rec.t;
      o.u = rec.u;
      o.v = rec.v;
      for (int c = 0; c < 3; c++) {
        o.p[c] = rec.point[c];
        o.n[c] = rec.normal[c];
      }
    }
  }
}

// camera ctor src/camera.cc:2-21; returns its public fields.
void ref_camera_make(const float *lookfrom, const float *lookat, const float *vup, float vfov,
                     float aspect, float aperture, float focus_dist, float t0, float t1,
                     ref_camera_out *o) {
  camera cam(vec3(lookfrom[0], lookfrom[1], lookfrom[2]), vec3(lookat[0], lookat[1], lookat[2]),
             vec3(vup[0], vup[1], vup[2]), vfov, aspect, aperture, focus_dist, t0, t1);
  for (int c = 0; c < 3; c++) {
    o->origin[c] = cam.origin_[c];
    o->lower_left[c] = cam.lower_left_corner_[c];
    o->vertical[c] = cam.vertical_[c];
    o->horizontal[c] = cam.horizontal_[c];
    o->u[c] = cam.u_[c];
    o->v[c] = cam.v_[c];
    o->w[c] = cam.w_[c];
  }
  o->lens_radius = cam.lens_radius_;
  o->time0 = cam.time0;
  o->time1 = cam.time1;
}

struct ref_render_args {
  // camera ctor arguments (main.cpp:87-91)
  float lookfrom[3], lookat[3], vup[3];
  float vfov, aspect, aperture, focus_dist, t0, t1;
  int32_t nx, ny, ns, max_depth; // main.cpp:31-37
  int32_t slices;                // bonus_pic if allow_bonus_pic else 1 (main.cpp:111-114)
  int32_t n_lights;              // main.cpp:99-106
  light_desc lights[8];
  int32_t deterministic; // 1: inject Philox stream keyed (pixel,sample,stage) (needs _det build)
  uint32_t seed_lo, seed_hi;
  int32_t threads; // 0 = hardware_concurrency
  int32_t count_rays;
  // window of pixels to render (x0<=i<x1, y0<=j<y1); others left untouched
  int32_t x0, y0, x1, y1;
};

struct ref_render_stats {
  uint64_t paths, rays, draws;
  double seconds;
  int32_t threads;
};

// Restatement of main.cpp:115-134: for each pixel, ns jittered samples, col += de_nan(color()).
// out_sum[slice][j][i][3] = running radiance sum after (slice+1)*ns/slices samples
// (what pixel_sample_cols holds, main.cpp:127-133); row j=0 is the BOTTOM row (v=(j+xi)/ny).
// out_samples (optional, may be NULL): [j][i][k][3] individual de_nan'ed samples.
int ref_render(void *s, const ref_render_args *a, float *out_sum, float *out_samples,
               ref_render_stats *st) {
  ref_scene *sc = static_cast<ref_scene *>(s);
  if (a->deterministic && !ref_rng_is_injected()) return -1;
  const int nx = a->nx, ny = a->ny, ns = a->ns;
  const int slices = a->slices > 0 ? a->slices : 1;
  const int per_slice = ns / slices;
  if (per_slice <= 0) return -2; // main.cpp:127 would SIGFPE on count % 0
  hitable *arr[8]; // leaked like every scene object in the reference (src/utils.cc:288-318)
  for (int i = 0; i < a->n_lights; i++) {
    const light_desc &l = a->lights[i];
    if (l.kind == 0)
      arr[i] = new xz_rect(l.p[0], l.p[1], l.p[2], l.p[3], l.p[4], nullptr);
    else
      arr[i] = new sphere(vec3(l.p[0], l.p[1], l.p[2]), l.p[3], nullptr);
  }
  hitable_list hlist(arr, a->n_lights);
  camera cam(vec3(a->lookfrom[0], a->lookfrom[1], a->lookfrom[2]),
             vec3(a->lookat[0], a->lookat[1], a->lookat[2]), vec3(a->vup[0], a->vup[1], a->vup[2]),
             a->vfov, a->aspect, a->aperture, a->focus_dist, a->t0, a->t1);
  const bool wrap = a->deterministic || a->count_rays;
  stage_counter counted(sc->world);
  hitable *world = wrap ? static_cast<hitable *>(&counted) : sc->world;

  int nthreads = a->threads > 0 ? a->threads : (int)std::thread::hardware_concurrency();
  if (nthreads < 1) nthreads = 1;
  std::atomic<int> next_row(a->y0);
  std::atomic<uint64_t> total_rays(0), total_draws(0);
  auto t_begin = std::chrono::high_resolution_clock::now();
  auto worker = [&]() {
    g_rays = 0;
    uint64_t draws0 = ref_rng_draws();
    for (;;) {
      int j = next_row.fetch_add(1);
      if (j >= a->y1) break;
      for (int i = a->x0; i < a->x1; i++) {
        vec3 col(0.0, 0.0, 0.0);
        int count = 0;
        for (int k = 0; k < ns; k++) {
          ++count;
          if (a->deterministic)
            ref_rng_begin_sample(a->seed_lo, a->seed_hi, (uint32_t)(j * nx + i), (uint32_t)k);
          float u = ((float)i + drand_r()) / (float)nx; // main.cpp:121
          float v = ((float)j + drand_r()) / (float)ny; // main.cpp:122
          ray r = cam.get_ray(u, v);                    // main.cpp:124
          vec3 tmp = color(r, world, &hlist, 0, a->max_depth); // main.cpp:125
          vec3 d = de_nan(tmp);
          col += d; // main.cpp:126
          if (out_samples) {
            float *o = out_samples + (((size_t)j * nx + i) * ns + k) * 3;
            o[0] = d[0];
            o[1] = d[1];
            o[2] = d[2];
          }
          if (count % per_slice == 0) { // main.cpp:127-133
            int sl = count / per_slice - 1;
            if (sl < slices) {
              float *o = out_sum + (((size_t)sl * ny + j) * nx + i) * 3;
              o[0] = col[0];
              o[1] = col[1];
              o[2] = col[2];
            }
          }
        }
      }
    }
    if (a->deterministic) ref_rng_end();
    total_rays += g_rays;
    total_draws += ref_rng_draws() - draws0;
  };
  std::vector<std::thread> pool;
  for (int t = 0; t < nthreads; t++) pool.emplace_back(worker);
  for (auto &t : pool) t.join();
  auto t_end = std::chrono::high_resolution_clock::now();
  if (st) {
    st->paths = (uint64_t)(a->x1 - a->x0) * (uint64_t)(a->y1 - a->y0) * (uint64_t)ns;
    st->rays = total_rays.load();
    st->draws = total_draws.load();
    st->seconds = std::chrono::duration<double>(t_end - t_begin).count();
    st->threads = nthreads;
  }
  return 0;
}

// A single color() call on a given ray under the deterministic stream (debug aid for tests).
int ref_color_one(void *s, const ref_render_args *a, const float *ray7, uint32_t pixel,
                  uint32_t sample, float *out3) {
  ref_scene *sc = static_cast<ref_scene *>(s);
  hitable *arr[8]; // leaked like every scene object in the reference (src/utils.cc:288-318)
  for (int i = 0; i < a->n_lights; i++) {
    const light_desc &l = a->lights[i];
    if (l.kind == 0)
      arr[i] = new xz_rect(l.p[0], l.p[1], l.p[2], l.p[3], l.p[4], nullptr);
    else
      arr[i] = new sphere(vec3(l.p[0], l.p[1], l.p[2]), l.p[3], nullptr);
  }
  hitable_list hlist(arr, a->n_lights);
  stage_counter counted(sc->world);
  if (a->deterministic) ref_rng_begin_sample(a->seed_lo, a->seed_hi, pixel, sample);
  ray r(vec3(ray7[0], ray7[1], ray7[2]), vec3(ray7[3], ray7[4], ray7[5]), ray7[6]);
  vec3 c = color(r, &counted, &hlist, 0, a->max_depth);
  if (a->deterministic) ref_rng_end();
  out3[0] = c[0];
  out3[1] = c[1];
  out3[2] = c[2];
  return 0;
}

// texture::value / perlin probes for unit-level known-answer tests (headers/texture.h:9-12)
void ref_perlin_turb(const float *pts, int n, float scale, float *out_rgb) {
  perlin_noise_texture *t = nullptr;
  {
    auto rv = perlin_noise::random_vec3_;
    auto px = perlin_noise::permute_x_, py = perlin_noise::permute_y_,
         pz = perlin_noise::permute_z_;
    t = new perlin_noise_texture(scale); // ctor re-randomises the static tables: restore
    perlin_noise::random_vec3_ = rv;
    perlin_noise::permute_x_ = px;
    perlin_noise::permute_y_ = py;
    perlin_noise::permute_z_ = pz;
  }
  for (int i = 0; i < n; i++) {
    vec3 c = t->value(0, 0, vec3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]));
    out_rgb[3 * i] = c[0];
    out_rgb[3 * i + 1] = c[1];
    out_rgb[3 * i + 2] = c[2];
  }
}

void ref_image_value(const unsigned char *img, int w, int h, const float *uv, int n,
                     float *out_rgb) {
  image_texture t(const_cast<unsigned char *>(img), w, h);
  for (int i = 0; i < n; i++) {
    vec3 c = t.value(uv[2 * i], uv[2 * i + 1], vec3(0, 0, 0));
    out_rgb[3 * i] = c[0];
    out_rgb[3 * i + 1] = c[1];
    out_rgb[3 * i + 2] = c[2];
  }
}

void ref_checker_value(const float *pts, int n, float *out_rgb) {
  checker_texture t(new constant_texture({0.1, 0.1, 0.1}), new constant_texture({0.9, 0.9, 0.9}));
  for (int i = 0; i < n; i++) {
    vec3 c = t.value(0, 0, vec3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]));
    out_rgb[3 * i] = c[0];
    out_rgb[3 * i + 1] = c[1];
    out_rgb[3 * i + 2] = c[2];
  }
}

// stbi_load through the reference's own wrapper (src/utils.cc:236-240). Returns malloc'ed RGB.
unsigned char *ref_load_image(const char *path, int *w, int *h, int *ch) {
  return load_image_texture(path, *w, *h, *ch);
}

// Leaves of the reference tree in DFS order with their geometry and material parameters, read
// from the public fields (headers/sphere.h:18-20,36-39, headers/rect_box.h:13-14,28-29,40-41,
// headers/material.h:36,49-50,55,71). Lets the tests prove that the product's scene builders +
// flattener describe exactly the scene the reference builds.
struct ref_leaf_out {
  int32_t kind;     // 0 sphere, 1 moving_sphere, 2 xy_rect, 3 xz_rect, 4 yz_rect
  int32_t mat_kind; // 0 lambertian, 1 metal, 2 dielectric, 3 diffuse_light, 4 other
  int32_t mat_id;
  int32_t tex_kind; // 0 constant, 1 checker, 2 perlin, 3 image, -1 none
  float p[12];
  float mat[5];     // metal: albedo rgb, fuzz ; dielectric: ref_idx ; constant texture: rgb ; perlin: scale
};

int ref_dump_leaves(void *s, ref_leaf_out *out, int cap) {
  ref_scene *sc = static_cast<ref_scene *>(s);
  std::vector<tagged_leaf *> leaves;
  std::map<hitable *, bool> seen;
  struct walk {
    static void go(hitable *h, std::vector<tagged_leaf *> &out, std::map<hitable *, bool> &seen) {
      if (seen[h]) return;
      seen[h] = true;
      if (auto *t = dynamic_cast<tagged_leaf *>(h)) out.push_back(t);
      else if (auto *b = dynamic_cast<bvh_node *>(h)) { go(b->left_, out, seen); go(b->right_, out, seen); }
      else if (auto *l = dynamic_cast<hitable_list *>(h)) { for (int i = 0; i < l->list_size_; i++) go(l->list_[i], out, seen); }
      else if (auto *bx = dynamic_cast<box *>(h)) go(bx->list_ptr_, out, seen);
      else if (auto *t2 = dynamic_cast<translate *>(h)) go(t2->ptr_, out, seen);
      else if (auto *r = dynamic_cast<rotate_y *>(h)) go(r->ptr_, out, seen);
      else if (auto *f = dynamic_cast<flip_normal *>(h)) go(f->ptr_, out, seen);
    }
  };
  walk::go(sc->world_tagged, leaves, seen);
  int n = 0;
  for (tagged_leaf *t : leaves) {
    if (n >= cap) break;
    ref_leaf_out &o = out[n++];
    std::memset(&o, 0, sizeof(o));
    hitable *h = t->inner_;
    material *m = leaf_material(h);
    if (auto *q = dynamic_cast<sphere *>(h)) {
      o.kind = 0; o.p[0] = q->center_[0]; o.p[1] = q->center_[1]; o.p[2] = q->center_[2]; o.p[3] = q->radius_;
    } else if (auto *q = dynamic_cast<moving_sphere *>(h)) {
      o.kind = 1;
      for (int c = 0; c < 3; c++) { o.p[c] = q->center0_[c]; o.p[4 + c] = q->center1_[c]; }
      o.p[3] = q->radius_; o.p[7] = q->time0_; o.p[8] = q->time1_;
    } else if (auto *q = dynamic_cast<xy_rect *>(h)) {
      o.kind = 2; o.p[0] = q->x0_; o.p[1] = q->x1_; o.p[2] = q->y0_; o.p[3] = q->y1_; o.p[4] = q->k_;
    } else if (auto *q = dynamic_cast<xz_rect *>(h)) {
      o.kind = 3; o.p[0] = q->x0_; o.p[1] = q->x1_; o.p[2] = q->z0_; o.p[3] = q->z1_; o.p[4] = q->k_;
    } else if (auto *q = dynamic_cast<yz_rect *>(h)) {
      o.kind = 4; o.p[0] = q->y0_; o.p[1] = q->y1_; o.p[2] = q->z0_; o.p[3] = q->z1_; o.p[4] = q->k_;
    }
    o.mat_id = -1;
    for (size_t i = 0; i < sc->mats.size(); i++) if (sc->mats[i] == m) o.mat_id = (int)i;
    texture *tex = nullptr;
    o.tex_kind = -1;
    if (auto *q = dynamic_cast<lambertian *>(m)) { o.mat_kind = 0; tex = q->albedo_; }
    else if (auto *q = dynamic_cast<metal *>(m)) { o.mat_kind = 1; o.mat[0] = q->albedo_[0]; o.mat[1] = q->albedo_[1]; o.mat[2] = q->albedo_[2]; o.mat[3] = q->fuzz_; }
    else if (auto *q = dynamic_cast<dielectric *>(m)) { o.mat_kind = 2; o.mat[0] = q->ref_idx_; }
    else if (auto *q = dynamic_cast<diffuse_light *>(m)) { o.mat_kind = 3; tex = q->emit_; }
    else o.mat_kind = 4;
    if (tex) {
      if (auto *q = dynamic_cast<constant_texture *>(tex)) { o.tex_kind = 0; o.mat[0] = q->color_[0]; o.mat[1] = q->color_[1]; o.mat[2] = q->color_[2]; }
      else if (dynamic_cast<checker_texture *>(tex)) o.tex_kind = 1;
      else if (auto *q = dynamic_cast<perlin_noise_texture *>(tex)) { o.tex_kind = 2; o.mat[0] = q->scale_; }
      else if (dynamic_cast<image_texture *>(tex)) o.tex_kind = 3;
    }
  }
  return n;
}

int ref_hardware_concurrency(void) { return (int)std::thread::hardware_concurrency(); }

} // extern "C"
