// host/tpt_host_capi.cc -- extern "C" access to the C++ front end for the Python tests and
// bench.py (ctypes): build a named scene with the host classes, flatten it, hand out the
// tpt_scene_desc; build a camera. No rendering here -- that is libtpt.so.
#include "tpt_flatten.h"
#include "tpt_image_io.h"
#include "tpt_scene.h"
#include "tpt_scene_programs.h"

#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {
thread_local std::string g_error;

struct host_scene {
  tpt::FlatScene flat;
  tpt_scene_desc desc;
};

// leaves of a bvh in left-to-right order, each once
void collect_leaves(hitable *h, std::vector<hitable *> &out) {
  if (auto *b = dynamic_cast<bvh_node *>(h)) {
    collect_leaves(b->left_, out);
    if (b->right_ != b->left_) collect_leaves(b->right_, out);
  } else {
    out.push_back(h);
  }
}

hitable *build_named(const std::string &name, const unsigned char *img, int iw, int ih) {
  if (name == "cornell_box") return cornell_box();
  if (name == "sphere_cornell_box") return sphere_cornell_box();
  if (name == "random_scene") return random_scene();
  if (name == "random_scene_list") {
    // BASELINE config 1 ("no BVH"): the same leaves in a flat hitable_list (the commented
    // alternative src/utils.cc:139), enumerated left-to-right from the built tree
    std::vector<hitable *> leaves;
    collect_leaves(random_scene(), leaves);
    hitable **arr = new hitable *[leaves.size()];
    for (size_t i = 0; i < leaves.size(); i++) arr[i] = leaves[i];
    return new hitable_list(arr, (int)leaves.size());
  }
  if (name == "two_perlin_spheres") return two_perlin_spheres();
  if (name == "two_checker_spheres") return two_checker_spheres();
  if (name == "light_spheres") return light_spheres();
  if (name == "cornell_box_smoke") return cornell_box_smoke();
  if (name == "oneweek_final") { // src/utils.cc:359-414; the earth texture is passed in by the caller
    if (!img) return nullptr;
    unsigned char *copy = new unsigned char[(size_t)iw * ih * 3];
    std::memcpy(copy, img, (size_t)iw * ih * 3);
    return oneweek_final_with(copy, iw, ih);
  }
  if (name == "textured_lit") {
    // test scene: every texture class under real light (HEAD's own textured scenes are black)
    if (!img) return nullptr;
    unsigned char *copy = new unsigned char[(size_t)iw * ih * 3];
    std::memcpy(copy, img, (size_t)iw * ih * 3);
    hitable **l = new hitable *[6];
    texture *checker = new checker_texture(new constant_texture({0.1, 0.1, 0.1}), new constant_texture({0.9, 0.9, 0.9}));
    l[0] = new sphere(vec3(0, 0, 0), 3, new lambertian(new image_texture(copy, iw, ih)));
    l[1] = new sphere(vec3(0, -1003, 0), 1000, new lambertian(checker));
    l[2] = new sphere(vec3(-3, 6, 4), 2, new diffuse_light(new constant_texture(vec3(8, 8, 8))));
    l[3] = new flip_normal(new xz_rect(-2, 2, -2, 2, 7, new diffuse_light(new constant_texture(vec3(4, 4, 4)))));
    l[4] = new sphere(vec3(5, -1, -2), 2, new lambertian(new perlin_noise_texture(2.0f)));
    return new hitable_list(l, 5);
  }
  if (name == "moving_list_test") {
    // test scene: a hitable_list ROOT that holds a moving_sphere which travels far outside its t = 0
    // position, a lamp and a floor -- list boxes must cover the whole motion (nothing in the reference
    // box-tests a hitable_list; the FAST walk and the scene-bounds tests do)
    hitable **l = new hitable *[3];
    l[0] = new moving_sphere(vec3(-6, 0, 0), vec3(6, 3, 0), 0.0f, 1.0f, 1.0f, new lambertian(new constant_texture({0.7, 0.3, 0.3})));
    l[1] = new flip_normal(new xz_rect(-2, 2, -2, 2, 6, new diffuse_light(new constant_texture(vec3(6, 6, 6)))));
    l[2] = new xz_rect(-8, 8, -4, 4, -1.5f, new lambertian(new constant_texture({0.6, 0.6, 0.6})));
    return new hitable_list(l, 3);
  }
  if (name.rfind("programp:", 0) == 0) return scene_programs::build((uint32_t)std::strtoul(name.c_str() + 9, nullptr, 10), false, false, true);
  if (name.rfind("programLm:", 0) == 0) return scene_programs::build((uint32_t)std::strtoul(name.c_str() + 10, nullptr, 10), true, true);
  if (name.rfind("programL:", 0) == 0) return scene_programs::build((uint32_t)std::strtoul(name.c_str() + 9, nullptr, 10), false, true);
  if (name.rfind("programm:", 0) == 0) return scene_programs::build((uint32_t)std::strtoul(name.c_str() + 9, nullptr, 10), true);
  if (name.rfind("program:", 0) == 0) return scene_programs::build((uint32_t)std::strtoul(name.c_str() + 8, nullptr, 10)); // tpt_scene_programs.h
  if (name == "earth") { // main.cpp:78-81
    if (!img) return nullptr;
    unsigned char *copy = new unsigned char[(size_t)iw * ih * 3];
    std::memcpy(copy, img, (size_t)iw * ih * 3);
    return new sphere(vec3(0, 0, 0), 3, new lambertian(new image_texture(copy, iw, ih)));
  }
  return nullptr;
}
} // namespace

extern "C" {

const char *tpt_host_last_error(void) { return g_error.c_str(); }

// lights == NULL -> the reference's hard-coded list (main.cpp:99-106)
void *tpt_host_build_scene(const char *name, const unsigned char *img, int iw, int ih,
                           const tpt_perlin_tables *force_perlin, const tpt_light *lights,
                           int n_lights, int background) {
  host_scene *hs = new host_scene();
  std::string err;
  bool ok = false;
  std::string nm(name ? name : "");
  // fresh thread => fresh default-seeded thread_local mt19937, i.e. the stream the reference's
  // main thread sees when it builds the scene first thing (main.cpp:74)
  std::thread builder([&] {
    hitable *world = build_named(nm, img, iw, ih);
    if (!world) {
      err = "unknown scene '" + nm + "' (or missing image)";
      return;
    }
    if (force_perlin) {
      for (int i = 0; i < 256; i++) {
        perlin_noise::random_vec3_[i] = vec3(force_perlin->ranvec[i][0], force_perlin->ranvec[i][1],
                                             force_perlin->ranvec[i][2]);
        perlin_noise::permute_x_[i] = force_perlin->perm_x[i];
        perlin_noise::permute_y_[i] = force_perlin->perm_y[i];
        perlin_noise::permute_z_[i] = force_perlin->perm_z[i];
      }
    }
    hitable *shapes[16];
    int n = 0;
    if (!lights) {
      shapes[n++] = new xz_rect(-100, 100, -150, -50, 298, nullptr);
      shapes[n++] = new sphere(vec3(120, -50, 40), 120, nullptr);
    } else {
      for (int i = 0; i < n_lights && i < 16; i++) {
        const tpt_light &L = lights[i];
        if (L.kind == TPT_LIGHT_XZ_RECT)
          shapes[n++] = new xz_rect(L.p[0], L.p[1], L.p[2], L.p[3], L.p[4], nullptr);
        else if (L.kind == TPT_LIGHT_SPHERE)
          shapes[n++] = new sphere(vec3(L.p[0], L.p[1], L.p[2]), L.p[3], nullptr);
        else
          shapes[n++] = new xy_rect(0, 0, 0, 0, 0, nullptr); // no pdf_value/random override
      }
    }
    hitable_list hlist(shapes, n);
    ok = tpt::flatten_scene(world, &hlist, background, hs->flat, err);
  });
  builder.join();
  if (!ok) {
    g_error = err;
    delete hs;
    return nullptr;
  }
  hs->desc = hs->flat.desc();
  return hs;
}

const tpt_scene_desc *tpt_host_scene_desc(void *h) { return &static_cast<host_scene *>(h)->desc; }
int tpt_host_scene_max_depth(void *h) { return static_cast<host_scene *>(h)->flat.max_depth; }
void tpt_host_scene_free(void *h) { delete static_cast<host_scene *>(h); }

void tpt_host_make_camera(const float *lookfrom, const float *lookat, const float *vup, float vfov,
                          float aspect, float aperture, float focus_dist, float t0, float t1,
                          tpt_camera *out) {
  camera cam(vec3(lookfrom[0], lookfrom[1], lookfrom[2]), vec3(lookat[0], lookat[1], lookat[2]),
             vec3(vup[0], vup[1], vup[2]), vfov, aspect, aperture, focus_dist, t0, t1);
  *out = tpt::make_camera_desc(cam);
}

// load_image_texture (the reference's wrapper around stbi_load, src/utils.cc:236-240): returns
// malloc'ed RGB bytes, caller frees with tpt_host_free
unsigned char *tpt_host_load_image(const char *path, int *w, int *h, int *ch) {
  return load_image_texture(path, *w, *h, *ch);
}
void tpt_host_free(void *p) { std::free(p); }

// picture writers in the reference's formats (main.cpp:69,183-189 and :197-211)
int tpt_host_write_ppm(const char *path, const unsigned char *rgb8, int nx, int ny, int bonus_format) {
  bool ok = bonus_format ? tpt::write_ppm_bonus(path, rgb8, nx, ny) : tpt::write_ppm_main(path, rgb8, nx, ny);
  return ok ? 0 : -1;
}

// binary PPM and the ImageMagick-free contact sheet: `n` pictures of nx x ny in the
// library's bottom-up rgb8 layout, side by side, as one baseline JPEG
int tpt_host_write_ppm_binary(const char *path, const unsigned char *rgb8, int nx, int ny) {
  return tpt::write_ppm_binary(path, rgb8, nx, ny) ? 0 : -1;
}
int tpt_host_write_contact_sheet(const char *path, const unsigned char *const *pictures, int n, int nx, int ny,
                                 int quality) {
  std::vector<const uint8_t *> pics(pictures, pictures + n);
  return tpt::write_contact_sheet(path, pics, nx, ny, quality) ? 0 : -1;
}
// emissive primitives of a built scene that the light-sampling list can represent
int tpt_host_derive_lights(void *h, tpt_light *out, int capacity) {
  std::vector<tpt_light> lights = tpt::derive_light_list(static_cast<host_scene *>(h)->flat);
  for (int i = 0; i < (int)lights.size() && i < capacity; i++) out[i] = lights[i];
  return (int)lights.size();
}

} // extern "C"
